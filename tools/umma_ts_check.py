#!/usr/bin/env python3
"""Does tcgen05.mma read A from tensor memory the way umma_selftest_ts_kernel writes it (lane = row, column k/2, even k
in the low half)?  D = A @ B^T for K = 64 ... 256 against torch; prints the relative error (expect ~1e-7)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
import testlib
LIB = testlib.load()
check = lambda rc, what: testlib.check(LIB, rc, what)
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for K in (64, 128, 256):
    g = torch.Generator(device=dev).manual_seed(K)
    A = torch.randn(128, K, device=dev, generator=g).bfloat16()
    B = torch.randn(144, K, device=dev, generator=g).bfloat16()
    D = torch.zeros(128, 144, device=dev)
    check(LIB.fn2b200_test_umma_gemm_ts(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()),
                                        K, st), "umma_ts")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    print("K=%3d  max|D - A B^T| / max|ref| = %.3e" % (K, float((D - ref).abs().max() / ref.abs().max())), flush=True)
