#!/usr/bin/env python3
"""Tensor-core correlation tuning sweep: ring depth (FN2B200_TC_BST) x L2 hint (FN2B200_TC_HINT) x C."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import flownet2_b200  # noqa: E402
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
print("torch ok", torch.cuda.get_device_name(0), flush=True)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C in (256, 128, 64):
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(8, C, 112, 256, device=dev, generator=g)
    b = torch.randn(8, C, 112, 256, device=dev, generator=g)
    go = torch.randn(8, 441, 112, 256, device=dev, generator=g)
    out = torch.empty(8, 441, 112, 256, device=dev)
    g1, g2 = torch.empty_like(a), torch.empty_like(b)
    for hint in (1,):
        for bst in (4,):
            os.environ["FN2B200_TC_BST"] = str(bst)
            os.environ["FN2B200_TC_HINT"] = str(hint)
            _, ws = F2.correlation_forward(a, b, 20, 1, 20, 1, 2, out=out, return_workspace=True)
            torch.cuda.synchronize()
            print("  first fwd done", C, hint, bst, flush=True)
            F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, out1=g1, out2=g2, workspace=ws)
            torch.cuda.synchronize()
            print("  first bwd done", flush=True)
            tf = timeit(lambda: F2.correlation_forward(a, b, 20, 1, 20, 1, 2, out=out))
            tb = timeit(lambda: F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, out1=g1, out2=g2, workspace=ws))
            print("C=%3d hint=%d bst<=%d  fwd(incl split) %.3f ms   bwd(2 launches, no split) %.3f ms" % (C, hint, bst, tf, tb), flush=True)
    del a, b, go, out, g1, g2
