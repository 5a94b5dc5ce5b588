"""ChannelNorm layer -- drop-in for networks/channelnorm_package/channelnorm.py (:5-38).

``norm_deg`` is stored and passed through but, as in the reference kernels, always L2.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.module import Module

from . import functional as F2


class ChannelNormFunction(Function):

    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        out = F2.channelnorm_forward(input1, norm_deg)
        if input1.dtype != torch.float32:
            out = out.to(input1.dtype)
        ctx.save_for_backward(input1, out)
        ctx.norm_deg = norm_deg
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        g = F2.channelnorm_backward(input1, output, grad_output, ctx.norm_deg)
        if g.dtype != input1.dtype:
            g = g.to(input1.dtype)
        return g, None


class ChannelNorm(Module):

    def __init__(self, norm_deg=2):
        super(ChannelNorm, self).__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)

    def extra_repr(self):
        return "norm_deg=%d" % self.norm_deg
