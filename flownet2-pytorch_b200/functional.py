"""Tensor-level entry points: torch tensors in, C-ABI calls out (device memory + stream plumbing).

These are the functions the autograd wrappers and the ``*_cuda`` shim modules share.  PyTorch is
used for allocation, the current device and the current stream only; all arithmetic happens in
libfn2b200's hand-written sm_100a kernels.
"""
import ctypes

import torch

from ._lib import LIB, check

_I64x4 = ctypes.c_int64 * 4


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "flownet2-pytorch_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU "
                "fallback -- the reference has none either (correlation.py:4)" % t.device.type)


def _f32c(t):
    """fp32 + contiguous (superset of the reference, which assumes both: SURVEY C-2)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    D, oH, oW = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(LIB.fn2b200_correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1,
                                            stride2, ctypes.byref(D), ctypes.byref(oH), ctypes.byref(oW)),
          "correlation_out_shape")
    return D.value, oH.value, oW.value


class CorrWorkspace(object):
    """Scratch of the tensor-core correlation path (bf16 hi/lo copies of input1/input2).  Returned by
    ``correlation_forward(..., return_workspace=True)`` and accepted by ``correlation_backward`` so the
    backward can skip the split pass when it runs on the SAME inputs (what autograd does)."""
    __slots__ = ("buf", "key")

    def __init__(self, buf, key):
        self.buf, self.key = buf, key


def _ws_key(a, b, prm):
    return (a.data_ptr(), b.data_ptr(), a._version, b._version, tuple(a.shape)) + tuple(prm)


def correlation_forward(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2,
                        corr_multiply=1, out=None, return_workspace=False):
    _require_cuda(input1, input2)
    if input1.dim() != 4 or input1.shape != input2.shape:
        raise ValueError("correlation: input1/input2 must be 4-D with equal shapes, got %s and %s"
                         % (tuple(input1.shape), tuple(input2.shape)))
    a, b = _f32c(input1), _f32c(input2)
    B, C, H, W = a.shape
    D, oH, oW = correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    if oH <= 0 or oW <= 0:
        raise ValueError("correlation: empty output (%d x %d) for input %dx%d, pad_size=%d, "
                         "max_displacement=%d, kernel_size=%d" % (oH, oW, H, W, pad_size, max_displacement, kernel_size))
    with torch.cuda.device_of(a):
        if out is None:
            out = torch.empty((B, D, oH, oW), dtype=torch.float32, device=a.device)
        ws_bytes = int(LIB.fn2b200_correlation_forward_workspace(B, C, H, W, pad_size, kernel_size, max_displacement,
                                                                 stride1, stride2))
        if ws_bytes:
            # scratch for the tensor-core path's bf16 hi/lo operands (the analogue of the reference's
            # rbot1/rbot2, correlation.py:20-21); the caching allocator keeps it stream-ordered
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
            check(LIB.fn2b200_correlation_forward_ws(_ptr(a), _ptr(b), _ptr(out), B, C, H, W, pad_size, kernel_size,
                                                     max_displacement, stride1, stride2, int(corr_multiply),
                                                     _ptr(ws), ws_bytes, _stream(a)),
                  "correlation_forward")
            wsobj = CorrWorkspace(ws, _ws_key(a, b, (pad_size, kernel_size, max_displacement, stride1, stride2)))
        else:
            check(LIB.fn2b200_correlation_forward(_ptr(a), _ptr(b), _ptr(out), B, C, H, W, pad_size, kernel_size,
                                                  max_displacement, stride1, stride2, int(corr_multiply), _stream(a)),
                  "correlation_forward")
            wsobj = None
    return (out, wsobj) if return_workspace else out


def correlation_backward(input1, input2, grad_output, pad_size, kernel_size, max_displacement, stride1,
                         stride2, corr_multiply=1, need1=True, need2=True, out1=None, out2=None, workspace=None):
    _require_cuda(input1, input2, grad_output)
    a, b, g = _f32c(input1), _f32c(input2), _f32c(grad_output)
    B, C, H, W = a.shape
    D, oH, oW = correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    if tuple(g.shape) != (B, D, oH, oW):
        raise ValueError("correlation_backward: grad_output shape %s != expected %s"
                         % (tuple(g.shape), (B, D, oH, oW)))
    with torch.cuda.device_of(a):
        g1 = (out1 if out1 is not None else torch.empty_like(a)) if need1 else None
        g2 = (out2 if out2 is not None else torch.empty_like(b)) if need2 else None
        ws_bytes = int(LIB.fn2b200_correlation_backward_workspace(B, C, H, W, pad_size, kernel_size,
                                                                  max_displacement, stride1, stride2))
        if ws_bytes:
            have = int(workspace is not None and workspace.buf.numel() >= ws_bytes and workspace.buf.device == a.device
                       and workspace.key == _ws_key(a, b, (pad_size, kernel_size, max_displacement, stride1, stride2)))
            ws = workspace.buf if have else torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
            check(LIB.fn2b200_correlation_backward_ws(_ptr(a), _ptr(b), _ptr(g), _ptr(g1), _ptr(g2), B, C, H, W,
                                                      pad_size, kernel_size, max_displacement, stride1, stride2,
                                                      int(corr_multiply), _ptr(ws), ws_bytes, have, _stream(a)),
                  "correlation_backward")
        else:
            check(LIB.fn2b200_correlation_backward(_ptr(a), _ptr(b), _ptr(g), _ptr(g1), _ptr(g2), B, C, H, W,
                                                   pad_size, kernel_size, max_displacement, stride1, stride2,
                                                   int(corr_multiply), _stream(a)),
                  "correlation_backward")
    return g1, g2


def _img_strides(img):
    return _I64x4(*img.stride())


def resample2d_forward(input1, input2, kernel_size=1, bilinear=True, out=None):
    _require_cuda(input1, input2)
    if input1.dim() != 4 or input2.dim() != 4 or input2.size(1) != 2:
        raise ValueError("resample2d: expected image [B,C,H,W] and flow [B,2,H,W], got %s and %s"
                         % (tuple(input1.shape), tuple(input2.shape)))
    img = input1 if input1.dtype == torch.float32 else input1.float()   # strided views accepted as-is
    flow = _f32c(input2)
    _, C, iH, iW = img.shape
    B, _, H, W = flow.shape
    if img.size(0) < B:
        raise ValueError("resample2d: image batch %d smaller than flow batch %d" % (img.size(0), B))
    with torch.cuda.device_of(flow):
        if out is None:
            out = torch.empty((B, C, H, W), dtype=torch.float32, device=flow.device)
        check(LIB.fn2b200_resample2d_forward(_ptr(img), _img_strides(img), _ptr(flow), _ptr(out), B, C, iH, iW,
                                             H, W, int(kernel_size), int(bool(bilinear)), _stream(flow)),
              "resample2d_forward")
    return out


def resample2d_backward(input1, input2, grad_output, kernel_size=1, bilinear=True, need1=True, need2=True,
                        out1=None, out2=None, zero_out1=True):
    _require_cuda(input1, input2, grad_output)
    img = input1 if input1.dtype == torch.float32 else input1.float()
    flow, g = _f32c(input2), _f32c(grad_output)
    _, C, iH, iW = img.shape
    B, _, H, W = flow.shape
    if tuple(g.shape) != (B, C, H, W):
        raise ValueError("resample2d_backward: grad_output shape %s != expected %s" % (tuple(g.shape), (B, C, H, W)))
    with torch.cuda.device_of(flow):
        g1 = None
        if need1:
            g1 = out1 if out1 is not None else torch.empty(tuple(img.shape), dtype=torch.float32, device=flow.device)
            if img.size(0) != B:      # rows the flow batch does not touch stay zero
                g1.zero_()
        g2 = (out2 if out2 is not None else torch.empty_like(flow)) if need2 else None
        check(LIB.fn2b200_resample2d_backward(_ptr(img), _img_strides(img), _ptr(flow), _ptr(g), _ptr(g1), _ptr(g2),
                                              B, C, iH, iW, H, W, int(kernel_size), int(bool(bilinear)),
                                              int(bool(zero_out1)), _stream(flow)),
              "resample2d_backward")
    return g1, g2


_DT16 = {torch.float16: 1, torch.bfloat16: 2}


def channelnorm_forward(input1, norm_deg=2, out=None):
    _require_cuda(input1)
    if input1.dim() != 4:
        raise ValueError("channelnorm: expected a 4-D tensor, got %s" % (tuple(input1.shape),))
    if input1.dtype in _DT16:        # native 16-bit kernels (the reference dispatches K8/K9 on half too)
        a = input1 if input1.is_contiguous() else input1.contiguous()
        B, C, H, W = a.shape
        with torch.cuda.device_of(a):
            if out is None:
                out = torch.empty((B, 1, H, W), dtype=a.dtype, device=a.device)
            check(LIB.fn2b200_channelnorm_forward_16(_ptr(a), _ptr(out), B, C, H, W, int(norm_deg), _DT16[a.dtype],
                                                     _stream(a)), "channelnorm_forward")
        return out
    a = _f32c(input1)
    B, C, H, W = a.shape
    with torch.cuda.device_of(a):
        if out is None:
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device=a.device)
        check(LIB.fn2b200_channelnorm_forward(_ptr(a), _ptr(out), B, C, H, W, int(norm_deg), _stream(a)),
              "channelnorm_forward")
    return out


def channelnorm_backward(input1, output, grad_output, norm_deg=2, out=None):
    _require_cuda(input1, output, grad_output)
    if input1.dtype in _DT16 and output.dtype == input1.dtype and grad_output.dtype == input1.dtype:
        a, o, g = (t if t.is_contiguous() else t.contiguous() for t in (input1, output, grad_output))
        B, C, H, W = a.shape
        with torch.cuda.device_of(a):
            if out is None:
                out = torch.empty_like(a)
            check(LIB.fn2b200_channelnorm_backward_16(_ptr(a), _ptr(o), _ptr(g), _ptr(out), B, C, H, W, int(norm_deg),
                                                      _DT16[a.dtype], _stream(a)), "channelnorm_backward")
        return out
    a, o, g = _f32c(input1), _f32c(output), _f32c(grad_output)
    B, C, H, W = a.shape
    with torch.cuda.device_of(a):
        if out is None:
            out = torch.empty_like(a)
        check(LIB.fn2b200_channelnorm_backward(_ptr(a), _ptr(o), _ptr(g), _ptr(out), B, C, H, W, int(norm_deg),
                                               _stream(a)),
              "channelnorm_backward")
    return out


def launch_count():
    """Kernel launches issued by libfn2b200 on this thread so far."""
    return int(LIB.fn2b200_launch_count())
