#!/usr/bin/env python3
"""Run one correlation case (pad,k,md,s1,s2,B,C,H,W) in-process and report parity; with --all, run a
list of cases each in its own subprocess (a sticky CUDA error must not poison the next case)."""
import subprocess
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["22,1,20,1,2,1,16,10,32", "18,1,20,1,2,1,16,14,32", "20,1,20,1,2,1,16,10,32", "3,1,3,1,1,1,5,8,8",
         "4,1,4,1,1,1,5,8,16", "4,1,4,1,2,2,7,9,12", "8,1,8,1,2,1,9,11,16", "5,1,5,1,2,1,6,8,8", "21,1,20,1,2,1,16,10,32"]


def one(spec, what):
    import numpy as np
    import torch
    import flownet2_b200 as f
    from oracle import cpu as orc
    pad, k, md, s1, s2, B, C, H, W = [int(v) for v in spec.split(",")]
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    if what in ("fwd", "both"):
        out = f.functional.correlation_forward(a.cuda(), b.cuda(), pad, k, md, s1, s2)
        torch.cuda.synchronize()
        ref = orc.correlation_forward(a.numpy(), b.numpy(), pad, k, md, s1, s2)
        print(spec, "fwd err", float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max()))
    if what in ("bwd", "both"):
        ref = orc.correlation_forward(a.numpy(), b.numpy(), pad, k, md, s1, s2)
        go = torch.randn(ref.shape, generator=g)
        g1, g2 = f.functional.correlation_backward(a.cuda(), b.cuda(), go.cuda(), pad, k, md, s1, s2)
        torch.cuda.synchronize()
        r1, r2 = orc.correlation_backward(a.numpy(), b.numpy(), go.numpy(), pad, k, md, s1, s2)
        print(spec, "bwd err", float(np.abs(g1.cpu().numpy() - r1).max() / np.abs(r1).max()),
              float(np.abs(g2.cpu().numpy() - r2).max() / np.abs(r2).max()))


if __name__ == "__main__":
    if sys.argv[1] == "--all":
        for c in CASES:
            for what in ("fwd", "bwd"):
                r = subprocess.run([sys.executable, __file__, c, what], capture_output=True, text=True, timeout=120)
                tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1]
                print("rc=%d %s %s :: %s" % (r.returncode, c, what, tail[:200]))
    else:
        one(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "both")
