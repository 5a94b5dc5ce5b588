#!/usr/bin/env python3
"""Rows per thread (FN2B200_RS_PY = 1, 2, 4) for the fused warp-concat kernel and the plain Resample2d kernels, cfg3, cold L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flownet2_b200
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
B, H, W, NS = 8, 448, 1024, 5
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.rand(B, 6, H, W, device=dev, generator=g) - 0.5 for _ in range(NS)]
lrs = [torch.randn(B, 2, H // 4, W // 4, device=dev, generator=g) * 0.2 for _ in range(NS)]
fls = [torch.randn(B, 2, H, W, device=dev, generator=g) * 4 for _ in range(NS)]
cats = [torch.empty(B, 12, H, W, device=dev) for _ in range(NS)]
gcs = [torch.randn(B, 12, H, W, device=dev, generator=g) for _ in range(2)]
def rot(fn, n=NS):
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
for py in ("1", "2", "4"):
    os.environ["FN2B200_RS_PY"] = py
    t1 = rot(lambda i: F2.warp_concat_forward(xs[i], lrs[i], upsample="bilinear", flow_mul=20.0, flow_div=20.0, out=cats[i]))
    t2 = rot(lambda i: F2.warp_concat_forward(xs[i], fls[i], flow_div=20.0, out=cats[i]))
    t3 = rot(lambda i: F2.resample2d_forward(xs[i][:, 3:], fls[i]))
    t4 = rot(lambda i: F2.warp_concat_backward(xs[i % 2], fls[i % 2], gcs[i % 2], flow_div=20.0), 2)
    print("PY", py, "fused fwd (quarter-res flow) %.1f us  fused fwd (full-res flow) %.1f us  resample fwd (strided img) %.1f us  fused bwd %.1f us" % (t1, t2, t3, t4), flush=True)
