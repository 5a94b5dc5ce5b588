#!/usr/bin/env python3
"""Cycles per M128 x N x K16 bf16 tcgen05.mma with both operands in shared memory vs A in tensor memory."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
import testlib
LIB = testlib.load()
check = lambda rc, what: testlib.check(LIB, rc, what)
dev = torch.device("cuda:0")
out = torch.zeros(256, dtype=torch.float32, device=dev)
dummy = torch.zeros(16, dtype=torch.bfloat16, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("mode (0 = SS K-major, 1 = A in TMEM, 2 = SS MN-major)   N   cycles/MMA (mean over SMs)   floor 128*N/256")
for mode in (0, 1, 2):
    for N in (64, 128, 144, 256):
        for rep in range(2):
            check(LIB.fn2b200_test_umma_rate(ctypes.c_void_p(out.data_ptr()), mode, N, 2000, st), "umma_rate")
            torch.cuda.synchronize()
        print("   %d    %3d    %7.1f    %5.0f" % (mode, N, out[:148].mean().item(), 128 * N / 256), flush=True)
