#!/usr/bin/env python3
"""Tiny driver for ncu: runs each hot-path op a few times at the benchmark shapes.
    ncu ... python tools/prof_ops.py [corr|small|all] [iters]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import flownet2_b200  # noqa: E402
F2 = flownet2_b200.functional
what = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
if what in ("corr", "all"):
    shp = (8, 256, 112, 256)
    a = torch.randn(*shp, device=dev, generator=g)
    b = torch.randn(*shp, device=dev, generator=g)
    go = torch.randn(8, 441, 112, 256, device=dev, generator=g)
    out = torch.empty(8, 441, 112, 256, device=dev)
    g1, g2 = torch.empty_like(a), torch.empty_like(b)
    for _ in range(iters):
        F2.correlation_forward(a, b, 20, 1, 20, 1, 2, out=out)
        F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, out1=g1, out2=g2)
    del a, b, go, out, g1, g2
if what in ("small", "all"):
    B, H, W = 8, 448, 1024
    img = torch.rand(B, 3, H, W, device=dev, generator=g)
    flow = torch.randn(B, 2, H, W, device=dev, generator=g) * 4
    go = torch.randn(B, 3, H, W, device=dev, generator=g)
    o, gi, gf = torch.empty_like(img), torch.empty_like(img), torch.empty_like(flow)
    n = torch.empty(B, 1, H, W, device=dev)
    gn = torch.randn(B, 1, H, W, device=dev, generator=g)
    for _ in range(iters):
        F2.resample2d_forward(img, flow, out=o)
        F2.resample2d_backward(img, flow, go, out1=gi, out2=gf)
        F2.channelnorm_forward(img, out=n)
        F2.channelnorm_backward(img, n, gn, out=gi)
    x = torch.rand(B, 6, H, W, device=dev, generator=g) - 0.5
    lr = torch.randn(B, 2, H // 4, W // 4, device=dev, generator=g) * 0.2
    cat = torch.empty(B, 12, H, W, device=dev)
    gc = torch.randn(B, 12, H, W, device=dev, generator=g)
    for _ in range(iters):
        F2.warp_concat_forward(x, lr, upsample="bilinear", flow_mul=20.0, flow_div=20.0, out=cat)
        F2.warp_concat_backward(x, flow, gc, flow_div=20.0)
torch.cuda.synchronize()
