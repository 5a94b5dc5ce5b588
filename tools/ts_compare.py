#!/usr/bin/env python3
"""Forward correlation on tensor cores: A_hi from shared memory (SS-mode MMAs, 3 accumulator buffers) vs from tensor
memory (tcgen05.cp + TS-mode MMAs, 2 accumulator buffers) -- FN2B200_TC_TS=0/1.  Parity against the oracle on small
shapes (incl. tail-unit segments and C < 256), bit-identity between the two modes, and CUDA-event timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import flownet2_b200
from oracle import cpu as orc
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
prm = (20, 1, 20, 1, 2)
for shape in ((1, 256, 48, 64), (2, 64, 10, 36), (1, 192, 30, 70), (5, 64, 32, 128), (1, 128, 2, 2)):
    g = torch.Generator().manual_seed(7)
    a, b = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    ref = orc.correlation_forward(a.numpy(), b.numpy(), *prm)
    outs = {}
    for ts in ("0", "1"):
        os.environ["FN2B200_TC_TS"] = ts
        outs[ts] = F2.correlation_forward(a.to(dev), b.to(dev), *prm).cpu().numpy()
        err = np.abs(outs[ts] - ref).max() / np.abs(ref).max()
        print(shape, "ts", ts, "rel err vs oracle %.2e" % err, flush=True)
        assert err < 1e-4
    print("   identical:", np.array_equal(outs["0"], outs["1"]), " max diff %.2e" % np.abs(outs["0"] - outs["1"]).max())
g = torch.Generator(device=dev).manual_seed(0)
for shp in ((8, 256, 112, 256), (8, 256, 56, 128), (8, 128, 112, 256)):
    a = torch.randn(*shp, device=dev, generator=g); b = torch.randn(*shp, device=dev, generator=g)
    out = torch.empty(shp[0], 441, shp[2], shp[3], device=dev)
    for ts in ("0", "1", "0", "1"):
        os.environ["FN2B200_TC_TS"] = ts
        for _ in range(3):
            F2.correlation_forward(a, b, *prm, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            F2.correlation_forward(a, b, *prm, out=out)
        e1.record(); torch.cuda.synchronize()
        print(shp, "ts", ts, "fwd incl. split %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3), flush=True)
os.environ.pop("FN2B200_TC_TS")
