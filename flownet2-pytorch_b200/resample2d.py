"""Resample2d (flow warp) layer -- drop-in for networks/resample2d_package/resample2d.py.

Differences that are strict supersets of the reference (resample2d.py:7-49): the image may be a
non-contiguous view (no ``.contiguous()`` copy, :48), outputs are not zero-filled before being
overwritten (:18), and the image-gradient buffer is zeroed inside the C-ABI call on the same stream.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.module import Module

from . import functional as F2


class Resample2dFunction(Function):

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        ctx.bilinear = bilinear
        out = F2.resample2d_forward(input1, input2, kernel_size, bilinear)
        return out if input1.dtype == torch.float32 else out.to(input1.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g1, g2 = F2.resample2d_backward(input1, input2, grad_output, ctx.kernel_size, ctx.bilinear,
                                        need1=need1, need2=need2)
        if g1 is not None and g1.dtype != input1.dtype:
            g1 = g1.to(input1.dtype)
        if g2 is not None and g2.dtype != input2.dtype:
            g2 = g2.to(input2.dtype)
        return g1, g2, None, None


class Resample2d(Module):
    """reference: resample2d.py:40-49 -- Resample2d(kernel_size=1, bilinear=True)."""

    def __init__(self, kernel_size=1, bilinear=True):
        super(Resample2d, self).__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        return Resample2dFunction.apply(input1, input2, self.kernel_size, self.bilinear)

    def extra_repr(self):
        return "kernel_size=%d, bilinear=%s" % (self.kernel_size, self.bilinear)
