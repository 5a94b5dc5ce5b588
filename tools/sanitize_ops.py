#!/usr/bin/env python3
"""Small invocations of every product kernel, meant to run UNDER compute-sanitizer
(tools/gpu_sanitize.sh: memcheck / synccheck / racecheck / initcheck; SURVEY section 5 "race detection").

    compute-sanitizer --tool racecheck --kernel-name regex:'corr_|resample2d|channelnorm|warp_' python tools/sanitize_ops.py [group]

Inputs are created on the host and copied (so torch launches as few of its own kernels as possible);
results are compared with the CPU oracle so that a sanitizer-clean run is also a correct run.
groups: tc (tensor-core correlation fwd+bwd), fma (TMA-tiled + generic correlation), rs (Resample2d),
cn (ChannelNorm), fused (warp->diff->norm->concat), all.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import flownet2_b200 as f
from oracle import cpu as orc

F2 = f.functional
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
errs = {}


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a.detach().cpu().numpy() if torch.is_tensor(a) else a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).float()


def corr_case(tag, shape, prm):
    a, b = rnd(shape, 1), rnd(shape, 2)
    out = F2.correlation_forward(a.to(dev), b.to(dev), *prm)
    ref = orc.correlation_forward(a.numpy(), b.numpy(), *prm)
    errs[tag + "_fwd"] = rel(out, ref)
    go = rnd(ref.shape, 3)
    g1, g2 = F2.correlation_backward(a.to(dev), b.to(dev), go.to(dev), *prm)
    r1, r2 = orc.correlation_backward(a.numpy(), b.numpy(), go.numpy(), *prm)
    errs[tag + "_g1"], errs[tag + "_g2"] = rel(g1, r1), rel(g2, r2)


if which in ("tc", "all"):
    # tensor-core path: ragged tiles, several k-blocks, a tile count that exercises the tail-unit split
    corr_case("tc_c64", (2, 64, 10, 36), (20, 1, 20, 1, 2))
    corr_case("tc_c256", (1, 256, 20, 36), (20, 1, 20, 1, 2))
    corr_case("tc_c192", (1, 192, 6, 70), (20, 1, 20, 1, 2))
if which in ("fma", "all"):
    corr_case("tiled", (1, 20, 13, 64), (20, 1, 20, 1, 2))
    corr_case("generic", (1, 6, 10, 12), (4, 3, 4, 1, 2))
if which in ("rs", "all"):
    for sigma, shp in ((4.0, (2, 3, 40, 72)), (64.0, (1, 3, 33, 50)), (2.0, (1, 5, 17, 23))):
        B, C, H, W = shp
        g = torch.Generator().manual_seed(5)
        img, flow = torch.rand(B, C, H, W, generator=g), torch.randn(B, 2, H, W, generator=g) * sigma
        go = torch.randn(B, C, H, W, generator=g)
        out = F2.resample2d_forward(img.to(dev), flow.to(dev))
        errs["rs%g_fwd" % sigma] = rel(out, orc.resample2d_forward(img.numpy(), flow.numpy()))
        g1, g2 = F2.resample2d_backward(img.to(dev), flow.to(dev), go.to(dev))
        r1, r2 = orc.resample2d_backward(img.numpy(), flow.numpy(), go.numpy())
        errs["rs%g_gimg" % sigma], errs["rs%g_gflow" % sigma] = rel(g1, r1), rel(g2, r2)
if which in ("cn", "all"):
    for shp in ((2, 3, 19, 36), (1, 2, 8, 8), (1, 5, 7, 9)):
        x = rnd(shp, 9)
        o = F2.channelnorm_forward(x.to(dev))
        ref = orc.channelnorm_forward(x.numpy())
        errs["cn%d_fwd" % shp[1]] = rel(o, ref)
        go = rnd(ref.shape, 10)
        errs["cn%d_bwd" % shp[1]] = rel(F2.channelnorm_backward(x.to(dev), o, go.to(dev)), orc.channelnorm_backward(x.numpy(), ref, go.numpy()))
        for dt in (torch.float16, torch.bfloat16):
            xh = x.to(dev).to(dt)
            oh = F2.channelnorm_forward(xh)
            F2.channelnorm_backward(xh, oh, go.to(dev).to(dt))
if which in ("fused", "all"):
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 6, 24, 40, generator=g)
    lr = torch.randn(2, 2, 6, 10, generator=g) * 0.3
    cat = F2.warp_concat_forward(x.to(dev), lr.to(dev), upsample="bilinear", flow_mul=20.0, flow_div=20.0)
    errs["fused_fwd"] = rel(cat, orc.warp_concat_forward(x.numpy(), lr.numpy(), upsample_mode=1, flow_mul=20.0, flow_div=20.0))
    fl = torch.randn(2, 2, 24, 40, generator=g) * 3
    gc = torch.randn(2, 12, 24, 40, generator=g)
    gx, gf = F2.warp_concat_backward(x.to(dev), fl.to(dev), gc.to(dev), flow_div=20.0)
    rx, rf = orc.warp_concat_backward(x.numpy(), fl.numpy(), gc.numpy(), flow_div=20.0)
    errs["fused_bwd_gx"], errs["fused_bwd_gflow"] = rel(gx, rx), rel(gf, rf)
    up = F2.resample2d_forward_up(x[:, 3:].to(dev), lr.to(dev), "nearest", 20.0)
    errs["resample_up"] = rel(up, orc.resample2d_forward(np.ascontiguousarray(x[:, 3:].numpy()), orc.upsample4(lr.numpy(), 2, 20.0)))
    a, b = rnd((1, 64, 12, 20), 12), rnd((1, 64, 12, 20), 13)
    buf = torch.zeros(1, 32 + 441, 12, 20, device=dev)
    F2.correlation_forward_cat(a.to(dev), b.to(dev), buf, 32, 0.1, 20, 1, 20, 1, 2)
    ref = orc.correlation_forward(a.numpy(), b.numpy(), 20, 1, 20, 1, 2)
    errs["corr_cat_leaky"] = rel(buf[:, 32:], np.where(ref > 0, ref, ref * np.float32(0.1)))
torch.cuda.synchronize()
bad = {k: v for k, v in errs.items() if not v <= 1e-4}
print("[sanitize_ops %s] launches=%d max_err=%.2e bad=%s" % (which, F2.launch_count(), max(errs.values()) if errs else 0.0, bad), flush=True)
sys.exit(1 if bad else 0)
