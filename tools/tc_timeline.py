#!/usr/bin/env python3
"""Per-unit timeline of the tensor-core backward kernel (CTA 0): builder thread 0 and the MMA thread
record clock64() at their synchronisation points (FN2B200_TC_DBG = device pointer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flownet2_b200
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(8, C, 112, 256, device=dev, generator=g)
b = torch.randn(8, C, 112, 256, device=dev, generator=g)
go = torch.randn(8, 441, 112, 256, device=dev, generator=g)
_, ws = F2.correlation_forward(a, b, 20, 1, 20, 1, 2, return_workspace=True)
F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, need2=False, workspace=ws)   # warm
if len(sys.argv) > 2 and sys.argv[2] == "fwd":
    dbg = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    os.environ["FN2B200_TC_DBG"] = str(dbg.data_ptr())
    F2.correlation_forward(a, b, 20, 1, 20, 1, 2)
    torch.cuda.synchronize()
    del os.environ["FN2B200_TC_DBG"]
    d = dbg.cpu().view(64, 8).tolist()
    t0 = d[0][0]
    print("C=%d forward, per unit (cycles rel. to first): MMA[start acc_ok end | acc.wait b.wait total]  EPI[start full_ok done | wait work]" % C)
    for u in range(0, 44):
        r = d[u]
        print("%3d  M %7d %7d %7d | %5d %5d %5d   E %7d %7d %7d | %5d %5d" % (
            u, r[0] - t0, r[1] - t0, r[3] - t0, r[1] - r[0], r[2], r[3] - r[0], r[4] - t0, r[5] - t0, r[6] - t0,
            r[5] - r[4], r[6] - r[5]))
    sys.exit(0)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
os.environ["FN2B200_TC_DBG"] = str(dbg.data_ptr())
F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, need2=False, workspace=ws)
torch.cuda.synchronize()
del os.environ["FN2B200_TC_DBG"]
d = dbg.cpu().view(64, 8).tolist()
t0 = d[0][0]
print("C=%d  unit: builder[wait_start wait_end fence_done arrived]  mma[wait_start wait_end issued]  (cycles rel. to first)" % C)
for u in range(0, 40):
    r = d[u]
    print("%3d  B %7d %7d %7d %7d   M %7d %7d %7d | b.wait %5d b.scatter %5d  m.wait(A) %5d m.issue %5d of which wait(B) %5d" % (
        u, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0,
        r[1] - r[0], r[2] - r[1], r[5] - r[4], r[6] - r[5], r[7]))
