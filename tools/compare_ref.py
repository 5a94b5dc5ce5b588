#!/usr/bin/env python3
"""Side-by-side kernel timings: ours vs the reference's own kernels rebuilt for sm_100a (oracle/_ref),
same inputs, cold L2 (256 MB flush between iterations), CUDA events.  Writes JSON to stdout.

    python tools/compare_ref.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import flownet2_b200  # noqa: E402
from bench import measured_peak, time_cold  # noqa: E402
from oracle import ref as oref  # noqa: E402

F2 = flownet2_b200.functional


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda:0")
    peak, _ = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    res = {}
    corr = oref.load_extension("correlation_cuda")
    rs = oref.load_extension("resample2d_cuda")
    cn = oref.load_extension("channelnorm_cuda")

    def rec(name, ours, ref, nbytes, it_ours=10, it_ref=3):
        ours()
        mo = time_cold(ours, it_ours, flush)
        e = {"ours_ms": round(mo, 4), "ours_GBps": round(nbytes / mo / 1e6, 1), "ours_frac_hbm": round(nbytes / mo / 1e6 / peak, 4)}
        if ref is not None:
            ref()
            mr = time_cold(ref, it_ref, flush)
            e.update({"ref_ms": round(mr, 4), "ref_GBps": round(nbytes / mr / 1e6, 1), "speedup": round(mr / mo, 1)})
        res[name] = e
        print(name, e, file=sys.stderr)

    for tag, shp in (("cfg2_8x256x112x256", (8, 256, 112, 256)), ("e2e_8x256x56x128", (8, 256, 56, 128)),
                     ("cfg1_1x256x48x64", (1, 256, 48, 64))):
        B, C, H, W = shp
        a = torch.randn(*shp, device=dev, generator=g)
        b = torch.randn(*shp, device=dev, generator=g)
        out = torch.empty(B, 441, H, W, device=dev)
        go = torch.randn(B, 441, H, W, device=dev, generator=g)
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        s1, s2, ro, rg1, rg2 = (a.new_empty(0) for _ in range(5))
        fb = 4 * (2 * B * C * H * W + B * 441 * H * W)
        bb = 4 * (B * 441 * H * W + 4 * B * C * H * W)
        rec("corr_fwd_" + tag, lambda: F2.correlation_forward(a, b, 20, 1, 20, 1, 2, out=out),
            (lambda: corr.forward(a, b, s1, s2, ro, 20, 1, 20, 1, 2, 1)) if corr else None, fb)
        rec("corr_bwd_" + tag, lambda: F2.correlation_backward(a, b, go, 20, 1, 20, 1, 2, out1=g1, out2=g2),
            (lambda: corr.backward(a, b, s1, s2, go, rg1, rg2, 20, 1, 20, 1, 2, 1)) if (corr and not (quick and B > 1)) else None,
            bb, it_ref=2)
        del a, b, out, go, g1, g2, s1, s2, ro, rg1, rg2
    B, H, W = 8, 448, 1024
    hw = B * H * W * 4
    for sigma in (4.0, 64.0):
        x6 = torch.rand(B, 6, H, W, device=dev, generator=g)
        img_nc = x6[:, 3:]
        img = img_nc.contiguous()
        flow = torch.randn(B, 2, H, W, device=dev, generator=g) * sigma
        go = torch.randn(B, 3, H, W, device=dev, generator=g)
        o, gi, gf = torch.empty_like(img), torch.empty_like(img), torch.empty_like(flow)

        def ref_fwd():
            ic = img_nc.contiguous()                      # resample2d.py:48
            oo = ic.new(B, 3, H, W).zero_()               # resample2d.py:18
            rs.forward(ic, flow, oo, 1, True)

        def ref_bwd():
            a1, a2 = torch.zeros_like(img), torch.zeros_like(flow)   # resample2d.py:31-32
            rs.backward(img, flow, go, a1, a2, 1, True)
        rec("resample2d_fwd_sigma%g" % sigma, lambda: F2.resample2d_forward(img_nc, flow, out=o), ref_fwd if rs else None, hw * 8)
        rec("resample2d_bwd_sigma%g" % sigma, lambda: F2.resample2d_backward(img, flow, go, out1=gi, out2=gf),
            ref_bwd if rs else None, hw * 13)
    for C in (3, 2):
        x = torch.randn(B, C, H, W, device=dev, generator=g)
        o = torch.empty(B, 1, H, W, device=dev)
        gon = torch.randn(B, 1, H, W, device=dev, generator=g)
        gi = torch.empty_like(x)

        def ref_fwd():
            oo = x.new(B, 1, H, W).zero_()                # channelnorm.py:11
            cn.forward(x, oo, 2)

        def ref_bwd():
            gg = torch.zeros_like(x)                      # channelnorm.py:23
            cn.backward(x, o, gon, gg, 2)
        rec("channelnorm_fwd_c%d" % C, lambda: F2.channelnorm_forward(x, out=o), ref_fwd if cn else None, hw * (C + 1))
        rec("channelnorm_bwd_c%d" % C, lambda: F2.channelnorm_backward(x, o, gon, out=gi), ref_bwd if cn else None, hw * (2 * C + 2))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
