#!/usr/bin/env python3
"""How fast can the SMs pull the correlation kernels' halo boxes through TMA with no consumer?"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flownet2_b200._lib import LIB, check
dev = torch.device("cuda:0")
nimg, Hc, Wc, C = 32, 56, 128, 256
x = torch.randn(nimg, Hc, Wc, C, device=dev).bfloat16()
out = torch.zeros(2 * 148, dtype=torch.int64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("box(w x h)  boxes/stage stages  KB in flight   B/clk/SM   chip TB/s @1.965GHz")
for (bw, bh, per, stages) in [(36, 4, 2, 2), (36, 4, 2, 3), (36, 4, 2, 5), (36, 4, 1, 4), (36, 4, 1, 10), (12, 4, 2, 6),
                              (12, 4, 8, 2), (12, 4, 8, 4), (4, 4, 8, 4), (4, 4, 8, 12), (36, 1, 8, 2), (36, 1, 8, 5), (16, 8, 2, 3),
                              (16, 8, 2, 6)]:
    iters = 600
    for rep in range(2):
        check(LIB.fn2b200_debug_tma_feed(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), nimg, C, Hc, Wc,
                                         bw, bh, stages, per, iters, 148, st), "tma_feed")
        torch.cuda.synchronize()
    o = out.view(148, 2).double()
    rate = (o[:, 1] / o[:, 0]).mean().item()
    print("%3d x %d      %2d        %2d      %6.1f       %6.1f      %6.2f" % (bw, bh, per, stages, per * stages * bw * bh * 128 / 1024, rate,
                                                                              rate * 148 * 1.965e9 / 1e12), flush=True)
