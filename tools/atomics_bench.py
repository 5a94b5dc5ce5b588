#!/usr/bin/env python3
"""Throughput of the reduction flavours a bilinear scatter can be built from (csrc_test/atomics_bench.cu).
Prints lane-operations per microsecond for the whole chip and cycles per lane-op per SM."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import testlib
LIB = testlib.load()
dev = torch.device("cuda:0")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
NAMES = {0: "red.global.f32", 1: "red.global.v2.f32", 2: "red.global.v4.f32", 3: "red.shared.f32 (CAS loop)",
         4: "shared ld+st (non-atomic)", 5: "red.global.f32, lane pairs adjacent"}
grid, iters = 148 * 8, 256
print("mode                                window(floats)  us      Mlane-ops/us   cyc/lane-op/SM   GB/s of payload")
for mode in (0, 5, 1, 2, 3, 4):
    for window in ((16384, 32768) if mode in (3, 4) else (32768, 262144)):
        buf = torch.zeros(grid * window if mode not in (3, 4) else 1024, device=dev)
        cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            testlib.check(LIB, LIB.fn2b200_test_atomics_bench(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(cyc.data_ptr()),
                                                              mode, window, iters, grid, st), "atomics_bench")
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        ops = grid * 256 * iters
        width = {0: 1, 1: 2, 2: 4, 3: 1, 4: 1, 5: 1}[mode]
        print("%-36s %8d   %8.1f   %8.1f        %6.2f          %7.1f" % (NAMES[mode], window, best, ops / best / 1e6 * 1e0,
              best * 1e-6 * 1.965e9 * 148 / ops, ops * width * 4 / best / 1e3), flush=True)
        del buf
