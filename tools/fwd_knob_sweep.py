#!/usr/bin/env python3
"""Forward tensor-core kernel: producer warps (FN2B200_TC_NP) x ring slots (FN2B200_TC_BST) x L2 hint, cfg2 and the FlowNet2 shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flownet2_b200
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
prm = (20, 1, 20, 1, 2)
g = torch.Generator(device=dev).manual_seed(0)
for shp in ((8, 256, 112, 256), (8, 256, 56, 128)):
    a = torch.randn(*shp, device=dev, generator=g); b = torch.randn(*shp, device=dev, generator=g)
    out = torch.empty(shp[0], 441, shp[2], shp[3], device=dev)
    for np_, bst, hint in ((4, 8, 1), (4, 7, 1), (4, 6, 1), (4, 5, 1), (4, 4, 1), (2, 8, 1), (1, 8, 1), (4, 8, 0)):
        os.environ.update(FN2B200_TC_NP=str(np_), FN2B200_TC_BST=str(bst), FN2B200_TC_HINT=str(hint))
        for _ in range(3):
            F2.correlation_forward(a, b, *prm, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F2.correlation_forward(a, b, *prm, out=out)
        e1.record(); torch.cuda.synchronize()
        print(shp, "NP", np_, "BST", bst, "HINT", hint, "fwd incl. split %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3), flush=True)
