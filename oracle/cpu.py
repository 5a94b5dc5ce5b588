"""numpy front-end of the CPU restatement ``oracle/oracle.c`` (TEST INFRASTRUCTURE ONLY).

The shared object is compiled on first use with ``-march=native`` for the host it runs on (the
GPU box's CPU may differ from the dev container's), into ``oracle/_build/`` (git-ignored), keyed
by the source hash + CPU model so a stale or foreign binary is never loaded.

Function names mirror the reference's extension entry points
(correlation_cuda.cc:169-172, resample2d_cuda.cc:28-31, channelnorm_cuda.cc:27-30); arguments
are numpy fp32 arrays, results are returned rather than written into out-parameters.
"""
import ctypes
import hashlib
import os
import platform
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle.c")
_LIB = None


def _cpu_tag():
    model = platform.machine()
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model += line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model


def build(verbose=False):
    """Compile oracle.c -> oracle/_build/liboracle_<hash>.so and return its path."""
    src = open(_SRC, "rb").read()
    tag = hashlib.sha1(src + _cpu_tag().encode()).hexdigest()[:12]
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle_%s.so" % tag)
    if not os.path.isfile(so):
        tmp = so + ".tmp%d" % os.getpid()
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fno-math-errno", "-shared", "-fPIC",
               "-o", tmp, _SRC, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        os.replace(tmp, so)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        i = ctypes.c_int
        L.orc_correlation_out_shape.argtypes = [i] * 8 + [ip, ip, ip]
        L.orc_correlation_out_shape.restype = None
        L.orc_correlation_forward.argtypes = [fp, fp, fp] + [i] * 9
        L.orc_correlation_backward.argtypes = [fp, fp, fp, fp, fp] + [i] * 9
        L.orc_resample2d_forward.argtypes = [fp, fp, fp] + [i] * 7
        L.orc_resample2d_backward.argtypes = [fp, fp, fp, fp, fp] + [i] * 6
        L.orc_channelnorm_forward.argtypes = [fp, fp] + [i] * 4
        L.orc_channelnorm_backward.argtypes = [fp, fp, fp, fp] + [i] * 4
        _LIB = L
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    D, oH, oW = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().orc_correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1,
                                    stride2, ctypes.byref(D), ctypes.byref(oH), ctypes.byref(oW))
    return D.value, oH.value, oW.value


def correlation_forward(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2,
                        corr_multiply=1):
    a, pa = _f(input1)
    b, pb = _f(input2)
    B, C, H, W = a.shape
    D, oH, oW = correlation_out_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    out = np.zeros((B, D, max(oH, 0), max(oW, 0)), np.float32)
    rc = lib().orc_correlation_forward(pa, pb, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                       B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    if rc:
        raise RuntimeError("oracle correlation_forward rc=%d" % rc)
    return out


def correlation_backward(input1, input2, grad_output, pad_size, kernel_size, max_displacement,
                         stride1, stride2, corr_multiply=1):
    a, pa = _f(input1)
    b, pb = _f(input2)
    g, pg = _f(grad_output)
    B, C, H, W = a.shape
    g1 = np.zeros_like(a)
    g2 = np.zeros_like(a)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib().orc_correlation_backward(pa, pb, pg, g1.ctypes.data_as(fp), g2.ctypes.data_as(fp),
                                        B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    if rc:
        raise RuntimeError("oracle correlation_backward rc=%d (stride1 must be 1)" % rc)
    return g1, g2


def resample2d_forward(input1, input2, kernel_size=1, bilinear=True):
    if kernel_size != 1:
        raise ValueError("oracle restates kernel_size == 1 only")
    a, pa = _f(input1)
    f, pf = _f(input2)
    _, C, iH, iW = a.shape
    B, _, H, W = f.shape
    out = np.zeros((B, C, H, W), np.float32)
    lib().orc_resample2d_forward(pa, pf, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 B, C, iH, iW, H, W, int(bool(bilinear)))
    return out


def resample2d_backward(input1, input2, grad_output, kernel_size=1, bilinear=True):
    if kernel_size != 1:
        raise ValueError("oracle restates kernel_size == 1 only")
    a, pa = _f(input1)
    f, pf = _f(input2)
    g, pg = _f(grad_output)
    _, C, iH, iW = a.shape
    B, _, H, W = f.shape
    g1 = np.zeros_like(a)
    g2 = np.zeros_like(f)
    fp = ctypes.POINTER(ctypes.c_float)
    lib().orc_resample2d_backward(pa, pf, pg, g1.ctypes.data_as(fp), g2.ctypes.data_as(fp),
                                  B, C, iH, iW, H, W)
    return g1, g2


def channelnorm_forward(input1, norm_deg=2):
    a, pa = _f(input1)
    B, C, H, W = a.shape
    out = np.zeros((B, 1, H, W), np.float32)
    lib().orc_channelnorm_forward(pa, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, C, H, W)
    return out


def channelnorm_backward(input1, output, grad_output, norm_deg=2):
    a, pa = _f(input1)
    o, po = _f(output)
    g, pg = _f(grad_output)
    B, C, H, W = a.shape
    gi = np.zeros_like(a)
    lib().orc_channelnorm_backward(pa, po, pg, gi.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B, C, H, W)
    return gi


def upsample4(planes, mode, mul=1.0):
    """[B,C,h,w] -> [B,C,4h,4w]; mode 1 = nn.Upsample(scale_factor=4, mode='bilinear'), 2 = 'nearest'; every input
    value is multiplied by ``mul`` first (models.py:130 ``upsample1(flow2 * div_flow)``)."""
    a, pa = _f(planes)
    B, C, h, w = a.shape
    out = np.zeros((B, C, 4 * h, 4 * w), np.float32)
    L = lib()
    L.orc_upsample4.argtypes = [ctypes.POINTER(ctypes.c_float)] * 2 + [ctypes.c_int] * 4 + [ctypes.c_float]
    rc = L.orc_upsample4(pa, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), B * C, h, w, int(mode), float(mul))
    if rc:
        raise ValueError("oracle upsample4: mode must be 1 (bilinear) or 2 (nearest)")
    return out


def warp_concat_forward(x, flow, C=3, upsample_mode=0, flow_mul=1.0, cat_channels=None, ch_x=0, n_x=None, ch_warped=None,
                        ch_flow=None, flow_div=1.0, ch_flow_norm=-1, ch_diff_norm=None):
    """The chain models.py:130-138 (and :142-150, :154-161, :167-174) builds from nn.Upsample, Resample2d, a
    subtraction, ChannelNorm, a division and torch.cat -- composed from the oracle's own ops, one array out:

        flow_up   = upsample4(flow * flow_mul)                 (or flow itself when upsample_mode == 0)
        warped    = resample2d(x[:, C:2C], flow_up)
        diff_norm = channelnorm(x[:, :C] - warped)
        cat[:, ch_x : ch_x + n_x] = x[:, :n_x];  cat[:, ch_warped ...] = warped;  cat[:, ch_flow ...] = flow_up / flow_div
        cat[:, ch_flow_norm] = channelnorm(flow_up);  cat[:, ch_diff_norm] = diff_norm

    Defaults give models.py:138's 12-channel layout (x | warped | flow / div_flow | diff norm).  Channels that no
    product is assigned to stay 0.  A negative ch_* skips that product."""
    x = np.ascontiguousarray(x, np.float32)
    n_x = 2 * C if n_x is None else n_x
    ch_warped = 2 * C if ch_warped is None else ch_warped
    ch_flow = 3 * C if ch_flow is None else ch_flow
    ch_diff_norm = 3 * C + 2 if ch_diff_norm is None else ch_diff_norm
    cat_channels = 3 * C + 3 if cat_channels is None else cat_channels
    flow_up = upsample4(flow, upsample_mode, flow_mul) if upsample_mode else np.ascontiguousarray(flow, np.float32)
    B, _, H, W = flow_up.shape
    warped = resample2d_forward(x[:, C:2 * C], flow_up)
    cat = np.zeros((B, cat_channels, H, W), np.float32)
    if ch_x >= 0:
        cat[:, ch_x:ch_x + n_x] = x[:, :n_x]
    if ch_warped >= 0:
        cat[:, ch_warped:ch_warped + C] = warped
    if ch_flow >= 0:
        cat[:, ch_flow:ch_flow + 2] = flow_up / np.float32(flow_div)
    if ch_flow_norm >= 0:
        cat[:, ch_flow_norm:ch_flow_norm + 1] = channelnorm_forward(flow_up)
    if ch_diff_norm >= 0:
        cat[:, ch_diff_norm:ch_diff_norm + 1] = channelnorm_forward(x[:, :C] - warped)
    return cat


def warp_concat_backward(x, flow, grad_cat, C=3, flow_div=1.0, ch_x=0, n_x=None, ch_warped=None, ch_flow=None, ch_flow_norm=-1,
                         ch_diff_norm=None):
    """Backward of warp_concat_forward(upsample_mode=0) as the composition of the reference modules' own backward passes:
    torch.cat (slicing), the division, ChannelNorm (channelnorm_kernel.cu:63-96), the subtraction, Resample2d
    (resample2d_kernel.cu:75-198).  Returns (grad_x [B,2C,H,W], grad_flow [B,2,H,W])."""
    x = np.ascontiguousarray(x, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    g = np.ascontiguousarray(grad_cat, np.float32)
    n_x = 2 * C if n_x is None else n_x
    ch_warped = 2 * C if ch_warped is None else ch_warped
    ch_flow = 3 * C if ch_flow is None else ch_flow
    ch_diff_norm = 3 * C + 2 if ch_diff_norm is None else ch_diff_norm
    img0, img1 = x[:, :C], np.ascontiguousarray(x[:, C:2 * C])
    warped = resample2d_forward(img1, flow)
    diff = img0 - warped
    g_diff = np.zeros_like(diff)
    if ch_diff_norm >= 0:
        g_diff = channelnorm_backward(diff, channelnorm_forward(diff), g[:, ch_diff_norm:ch_diff_norm + 1])
    g_warped = -g_diff
    if ch_warped >= 0:
        g_warped = g_warped + g[:, ch_warped:ch_warped + C]
    g_img1, g_flow = resample2d_backward(img1, flow, g_warped)
    gx = np.zeros((x.shape[0], 2 * C) + x.shape[2:], np.float32)
    gx[:, :C] = g_diff
    gx[:, C:] = g_img1
    if ch_x >= 0:
        gx[:, :n_x] += g[:, ch_x:ch_x + n_x]
    if ch_flow >= 0:
        g_flow = g_flow + g[:, ch_flow:ch_flow + 2] / np.float32(flow_div)
    if ch_flow_norm >= 0:
        g_flow = g_flow + channelnorm_backward(flow, channelnorm_forward(flow), g[:, ch_flow_norm:ch_flow_norm + 1])
    return gx, g_flow.astype(np.float32)
