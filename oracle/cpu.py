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
