"""Opt-in fused inference forwards for the reference's FlowNetC / FlowNet2C / FlowNet2.

The drop-in modules (``Correlation``, ``Resample2d``, ``ChannelNorm``) keep the reference's module graph: each layer
reads and writes whole tensors, and ``models.py`` glues them with ``nn.Upsample``, a subtraction, a division and
``torch.cat``.  The functions here take an UNMODIFIED, already constructed reference network (its sub-modules, weights
and hyper-parameters are used as they are) and run the same data flow with the glue folded into our kernels:

* ``flownetc_forward``  -- FlowNetC.py:70-126: ``LeakyReLU(0.1)(corr(a, b))`` is written by the correlation kernel's
  epilogue straight into channels 32..472 of the 473-channel input of ``conv3_1`` (8f-2): no activation pass over the
  cost volume, no ``torch.cat`` copy of it.
* ``flownet2_forward``  -- models.py:120-185: every ``upsample -> Resample2d -> diff -> ChannelNorm -> cat`` group is
  ONE kernel (``functional.warp_concat_forward``, 8f-1) that reads the quarter-resolution flow directly (8f-3) and writes
  the concat buffer the next sub-network consumes.

The model-level forwards are inference only (``torch.no_grad`` is entered there); for training graphs there is
``WarpConcat`` / ``WarpConcatFunction`` (full-resolution flow, fused forward AND fused backward) next to the
differentiable drop-in modules.  Results
agree with the unfused graph to rounding (tests/test_gpu_parity.py::test_fused_forwards_match_unfused_models).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import functional as F2


class WarpConcatFunction(Function):
    """concat1 = cat(x, Resample2d()(x[:, 3:], flow), flow / div_flow, ChannelNorm()(x[:, :3] - warped)) as one
    differentiable op (models.py:133-138, :145-150); flow is the full-resolution flow (keep nn.Upsample in front of it
    when training: its backward is torch's).  forward and backward are one kernel each."""

    @staticmethod
    def forward(ctx, x, flow, div_flow=20.0):
        ctx.save_for_backward(x, flow)
        ctx.div_flow = float(div_flow)
        return F2.warp_concat_forward(x, flow, C=x.size(1) // 2, flow_div=ctx.div_flow)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cat):
        x, flow = ctx.saved_tensors
        gx, gf = F2.warp_concat_backward(x, flow, grad_cat, C=x.size(1) // 2, flow_div=ctx.div_flow)
        if x.size(1) > gx.size(1):           # odd trailing channels of x take no part
            gx = torch.cat((gx, gx.new_zeros(gx.size(0), x.size(1) - gx.size(1), gx.size(2), gx.size(3))), 1)
        return (gx if ctx.needs_input_grad[0] else None), (gf if ctx.needs_input_grad[1] else None), None


class WarpConcat(torch.nn.Module):
    """nn.Module form of WarpConcatFunction: WarpConcat(div_flow)(x, flow) -> the (3C + 3)-channel concat."""

    def __init__(self, div_flow=20.0):
        super(WarpConcat, self).__init__()
        self.div_flow = div_flow

    def forward(self, x, flow):
        return WarpConcatFunction.apply(x, flow, self.div_flow)


def _corr_module(net_c):
    """FlowNetC.corr is the Correlation module itself, or nn.Sequential(tofp32, Correlation, tofp16) in --fp16 mode
    (FlowNetC.py:25-31)."""
    corr = net_c.corr
    if isinstance(corr, torch.nn.Sequential):
        corr = [m for m in corr if hasattr(m, "max_displacement")][0]
    return corr


def flownetc_forward(net_c, x):
    """networks/FlowNetC.py:70-126 in eval mode -> (flow2,).  x: [B,6,H,W]."""
    return _flownetc_pair(net_c, x[:, 0:3], x[:, 3:])


def _flownetc_pair(net_c, x1, x2):
    """FlowNetC on the two frames given separately (FlowNet2C feeds x[:,:,0] and x[:,:,1], models.py:192-194)."""
    with torch.no_grad():
        out_conv1a = net_c.conv1(x1)
        out_conv2a = net_c.conv2(out_conv1a)
        out_conv3a = net_c.conv3(out_conv2a)
        out_conv3b = net_c.conv3(net_c.conv2(net_c.conv1(x2)))
        out_conv_redir = net_c.conv_redir(out_conv3a)
        corr = _corr_module(net_c)
        B, C, H, W = out_conv3a.shape
        D, oH, oW = F2.correlation_out_shape(C, H, W, corr.pad_size, corr.kernel_size, corr.max_displacement, corr.stride1,
                                             corr.stride2)
        nredir = out_conv_redir.size(1)
        in_conv3_1 = torch.empty((B, nredir + D, oH, oW), dtype=torch.float32, device=x1.device)
        in_conv3_1[:, :nredir].copy_(out_conv_redir)
        F2.correlation_forward_cat(out_conv3a, out_conv3b, in_conv3_1, nredir, net_c.corr_activation.negative_slope,
                                   corr.pad_size, corr.kernel_size, corr.max_displacement, corr.stride1, corr.stride2,
                                   corr.corr_multiply)
        out_conv3_1 = net_c.conv3_1(in_conv3_1.to(out_conv_redir.dtype))
        out_conv4 = net_c.conv4_1(net_c.conv4(out_conv3_1))
        out_conv5 = net_c.conv5_1(net_c.conv5(out_conv4))
        out_conv6 = net_c.conv6_1(net_c.conv6(out_conv5))
        flow6 = net_c.predict_flow6(out_conv6)
        concat5 = torch.cat((out_conv5, net_c.deconv5(out_conv6), net_c.upsampled_flow6_to_5(flow6)), 1)
        flow5 = net_c.predict_flow5(concat5)
        concat4 = torch.cat((out_conv4, net_c.deconv4(concat5), net_c.upsampled_flow5_to_4(flow5)), 1)
        flow4 = net_c.predict_flow4(concat4)
        concat3 = torch.cat((out_conv3_1, net_c.deconv3(concat4), net_c.upsampled_flow4_to_3(flow4)), 1)
        flow3 = net_c.predict_flow3(concat3)
        concat2 = torch.cat((out_conv2a, net_c.deconv2(concat3), net_c.upsampled_flow3_to_2(flow3)), 1)
        return (net_c.predict_flow2(concat2),)


def _normalise(inputs, rgb_max):
    """models.py:121-126: subtract the per-sample, per-channel mean over both frames, divide by rgb_max."""
    rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1,))
    return (inputs - rgb_mean) / rgb_max


def flownet2c_forward(net, inputs):
    """models.FlowNet2C.forward (models.py:189-243) in eval mode: upsample1(flow2 * div_flow)."""
    with torch.no_grad():
        x = _normalise(inputs, net.rgb_max)
        flow2 = _flownetc_pair(net, x[:, :, 0], x[:, :, 1])[0]
        return net.upsample1(flow2 * net.div_flow)


def flownet2_forward(net, inputs):
    """models.FlowNet2.forward (models.py:120-185) with the warp groups fused.  net: the reference's FlowNet2."""
    with torch.no_grad():
        div = float(net.div_flow)
        x = _normalise(inputs, net.rgb_max)
        x = torch.cat((x[:, :, 0], x[:, :, 1]), dim=1).float()
        B, _, H, W = x.shape
        # flownetc -> concat1 = (x | img1 warped | flow / div_flow | |img0 - warped|)            models.py:129-138
        flow2 = flownetc_forward(net.flownetc, x)[0]
        concat1 = F2.warp_concat_forward(x, flow2, upsample="bilinear", flow_mul=div, flow_div=div)
        # flownets1 -> concat2                                                                    models.py:141-150
        flow2 = net.flownets_1(concat1)[0]
        concat2 = F2.warp_concat_forward(x, flow2, upsample="bilinear", flow_mul=div, flow_div=div)
        del concat1
        # flownets2 and flownetsd -> concat3 = (img0 | sd flow | s2 flow | their norms | the two diff norms)   :153-174
        s2_flow2 = net.flownets_2(concat2)[0]
        del concat2
        sd_flow2 = net.flownets_d(x)[0]
        concat3 = torch.empty((B, 11, H, W), dtype=torch.float32, device=x.device)
        F2.warp_concat_forward(x, sd_flow2, upsample="nearest", flow_mul=1.0 / div, flow_div=1.0, out=concat3, ch_x=0, n_x=3,
                               ch_warped=-1, ch_flow=3, ch_flow_norm=7, ch_diff_norm=9)
        F2.warp_concat_forward(x, s2_flow2, upsample="nearest", flow_mul=div, flow_div=1.0, out=concat3, ch_x=-1, n_x=0,
                               ch_warped=-1, ch_flow=5, ch_flow_norm=8, ch_diff_norm=10)
        return net.flownetfusion(concat3)


def fused_forward(net, inputs):
    """Dispatch on the reference class name: FlowNet2, FlowNet2C or a bare FlowNetC."""
    name = type(net).__name__
    if name == "FlowNet2":
        return flownet2_forward(net, inputs)
    if name == "FlowNet2C":
        return flownet2c_forward(net, inputs)
    if name == "FlowNetC":
        return flownetc_forward(net, inputs)
    raise TypeError("fused_forward: no fused data flow for %s (FlowNet2, FlowNet2C, FlowNetC)" % name)
