#!/usr/bin/env python3
"""Backward tensor-core kernel: ring slots (FN2B200_TC_BST) x producer warps (FN2B200_TC_NP), cfg2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flownet2_b200
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
prm = (20, 1, 20, 1, 2)
g = torch.Generator(device=dev).manual_seed(0)
shp = (8, 256, 112, 256)
a = torch.randn(*shp, device=dev, generator=g); b = torch.randn(*shp, device=dev, generator=g)
go = torch.randn(8, 441, 112, 256, device=dev, generator=g)
g1, g2 = torch.empty_like(a), torch.empty_like(b)
_, ws = F2.correlation_forward(a, b, *prm, return_workspace=True)
for np_, bst in ((3, 3), (3, 2), (2, 3), (1, 3), (2, 2), (3, 3)):
    os.environ.update(FN2B200_TC_NP=str(np_), FN2B200_TC_BST=str(bst))
    for _ in range(3):
        F2.correlation_backward(a, b, go, *prm, out1=g1, out2=g2, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F2.correlation_backward(a, b, go, *prm, out1=g1, out2=g2, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print("NP", np_, "BST", bst, "bwd %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3), flush=True)
