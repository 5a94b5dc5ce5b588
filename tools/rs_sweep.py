#!/usr/bin/env python3
"""Resample2d at cfg3 ([8,3,448,1024]): round 1's row kernels vs the TMA-staged kernels vs the 2-D tile kernels (rows per
thread, backward scatter flavour, halo of the shared-memory accumulation box),
for sigma = 4 and 64 px random flows and a smooth large-displacement flow; plus the fused warp->diff->norm->concat
kernel against the chain of individual ops.  Rotating buffer sets (cold L2), median of 3; GB/s of algorithmic bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flownet2_b200
F2 = flownet2_b200.functional
dev = torch.device("cuda:0")
PEAK = 6587.7
B, H, W, NS = 8, 448, 1024, 6
hw = B * H * W * 4
g = torch.Generator(device=dev).manual_seed(0)


def rotating(make_call, nsets=NS, reps=3):
    calls = [make_call(i) for i in range(nsets)]
    for c in calls:
        c()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in calls:
            c()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / nsets)
    out.sort()
    return out[len(out) // 2]


imgs = [torch.rand(B, 3, H, W, device=dev, generator=g) for _ in range(NS)]
gos = [torch.randn(B, 3, H, W, device=dev, generator=g) for _ in range(NS)]
o3 = [torch.empty(B, 3, H, W, device=dev) for _ in range(NS)]
o2 = [torch.empty(B, 2, H, W, device=dev) for _ in range(NS)]
yy, xx = torch.meshgrid(torch.arange(H, device=dev).float(), torch.arange(W, device=dev).float(), indexing="ij")
smooth = torch.stack((40 + 10 * torch.sin(yy / 50) + 0.02 * xx, -25 + 8 * torch.cos(xx / 70)), 0).unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
flowsets = {"sigma4": [torch.randn(B, 2, H, W, device=dev, generator=g) * 4 for _ in range(NS)],
            "sigma64": [torch.randn(B, 2, H, W, device=dev, generator=g) * 64 for _ in range(NS)],
            "smooth40": [smooth + 0.3 * torch.randn(B, 2, H, W, device=dev, generator=g) for _ in range(NS)]}


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


ALLK = ("FN2B200_RESAMPLE", "FN2B200_RS_TILE_H", "FN2B200_RS_HALO", "FN2B200_RS_PY", "FN2B200_RS_BWD")
configs = [("row (round 1)", dict(FN2B200_RESAMPLE="row"))]
for py in (1, 2, 4):
    configs.append(("tile py=%d planar" % py, dict(FN2B200_RS_PY=py, FN2B200_RS_BWD="planar")))
for py in (2, 4):
    configs.append(("tile py=%d vec" % py, dict(FN2B200_RS_PY=py, FN2B200_RS_BWD="vec")))
print("%-26s %-9s  fwd us  frac   bwd us  frac   bwd(flow only) us   bwd(img only) us" % ("config", "flow"))
for name, env in configs:
    setenv(**{k: env.get(k) for k in ALLK})
    for fname, flows in flowsets.items():
        if fname != "sigma4" and "py=1" in name:
            continue
        f = rotating(lambda i: (lambda: F2.resample2d_forward(imgs[i], flows[i], out=o3[i])))
        bw = rotating(lambda i: (lambda: F2.resample2d_backward(imgs[i], flows[i], gos[i], out1=o3[i], out2=o2[i])))
        bf = rotating(lambda i: (lambda: F2.resample2d_backward(imgs[i], flows[i], gos[i], need1=False, out2=o2[i])))
        bi = rotating(lambda i: (lambda: F2.resample2d_backward(imgs[i], flows[i], gos[i], need2=False, out1=o3[i])))
        print("%-26s %-9s %7.1f  %.3f %7.1f  %.3f   %7.1f             %7.1f" % (name, fname, f * 1e3, hw * 8 / f / 1e6 / PEAK, bw * 1e3,
                                                                               hw * 13 / bw / 1e6 / PEAK, bf * 1e3, bi * 1e3), flush=True)
setenv(**{k: None for k in ALLK})
del o3, o2, gos
# fused warp -> diff -> norm -> concat (models.py:130-138) vs the chain of individual ops
xs = [torch.rand(B, 6, H, W, device=dev, generator=g) - 0.5 for _ in range(NS)]
lrs = [torch.randn(B, 2, H // 4, W // 4, device=dev, generator=g) * 0.2 for _ in range(NS)]
cats = [torch.empty(B, 12, H, W, device=dev) for _ in range(NS)]
up = torch.nn.Upsample(scale_factor=4, mode="bilinear")
rs, cn = flownet2_b200.Resample2d(), flownet2_b200.ChannelNorm()


def chain(i):
    def run():
        x = xs[i]
        fl = up(lrs[i] * 20.0)
        warped = rs(x[:, 3:], fl)
        torch.cat((x, warped, fl / 20.0, cn(x[:, :3] - warped)), dim=1, out=cats[i])
    return run


with torch.no_grad():
    t_chain = rotating(chain)
    t_fused = rotating(lambda i: (lambda: F2.warp_concat_forward(xs[i], lrs[i], upsample="bilinear", flow_mul=20.0, flow_div=20.0, out=cats[i])))
alg = hw * (6 + 12) + B * 2 * (H // 4) * (W // 4) * 4
print("\nwarp->diff->norm->concat [8,6,448,1024] + quarter-res flow: chain of ops %.1f us, fused kernel %.1f us (%.2fx); fused moves "
      "%.0f MB algorithmic = %.0f GB/s = %.3f of HBM peak" % (t_chain * 1e3, t_fused * 1e3, t_chain / t_fused, alg / 1e6, alg / t_fused / 1e6,
                                                               alg / t_fused / 1e6 / PEAK))
