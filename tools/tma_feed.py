#!/usr/bin/env python3
"""How fast can the SMs pull the correlation kernels' halo boxes through TMA with no consumer?
Three sweeps: box shape / ring depth (all SMs), number of active SMs (per-SM or chip-wide ceiling?),
and thread-block clusters with TMA multicast (does one L2 read feeding 2/4 SMs lift the per-SM rate?)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
import testlib
LIB = testlib.load()
check = lambda rc, what: testlib.check(LIB, rc, what)
dev = torch.device("cuda:0")
nimg, Hc, Wc, C = 32, 56, 128, 256
x = torch.randn(nimg, Hc, Wc, C, device=dev).bfloat16()
out = torch.zeros(2 * 148, dtype=torch.int64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(bw, bh, per, stages, grid=148, cluster=1, iters=600, warps=1):
    out.zero_()
    for rep in range(2):
        check(LIB.fn2b200_test_tma_feed(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), nimg, C, Hc, Wc,
                                         bw, bh, stages, per, iters, grid, cluster, warps, st), "tma_feed")
        torch.cuda.synchronize()
    o = out.view(148, 2)[:grid].double()
    return (o[:, 1] / o[:, 0]).mean().item()


if which in ("all", "boxes"):
    print("box(w x h)  boxes/stage stages  KB in flight   B/clk/SM   chip TB/s @1.965GHz")
    for (bw, bh, per, stages) in [(36, 4, 2, 2), (36, 4, 2, 3), (36, 4, 1, 4), (12, 4, 8, 2), (12, 4, 8, 4), (4, 4, 8, 4),
                                  (16, 8, 2, 3)]:
        rate = run(bw, bh, per, stages)
        print("%3d x %d      %2d        %2d      %6.1f       %6.1f      %6.2f" % (bw, bh, per, stages, per * stages * bw * bh * 128 / 1024,
                                                                                  rate, rate * 148 * 1.965e9 / 1e12), flush=True)
if which in ("all", "stage"):
    print("\nbytes per stage vs ring depth (same bytes in flight arranged differently)")
    print("box(w x h)  boxes/stage stages  KB/stage  KB in flight   B/clk/SM")
    for (bw, bh, per, stages) in [(36, 4, 1, 8), (36, 4, 2, 4), (36, 4, 4, 2), (36, 4, 3, 3), (36, 4, 6, 2), (36, 4, 1, 2), (36, 4, 1, 12),
                                  (12, 4, 2, 8), (12, 4, 4, 4), (12, 4, 8, 2), (12, 4, 16, 2), (12, 4, 12, 3), (12, 4, 4, 8)]:
        rate = run(bw, bh, per, stages)
        print("%3d x %d      %2d        %2d     %6.1f     %6.1f       %6.1f" % (bw, bh, per, stages, per * bw * bh * 128 / 1024,
                                                                             per * stages * bw * bh * 128 / 1024, rate), flush=True)
if which in ("all", "warps"):
    print("\nproducer warps, each with a private ring (36x4 box = 18 KB; 12x4 = 6 KB)")
    print("box(w x h)  boxes/stage stages warps  KB in flight   B/clk/SM")
    for (bw, bh, per, stages, warps) in [(36, 4, 1, 2, 1), (36, 4, 1, 2, 2), (36, 4, 1, 2, 4), (36, 4, 1, 3, 4), (36, 4, 2, 2, 2),
                                         (36, 4, 2, 1, 4), (36, 4, 1, 1, 8), (12, 4, 4, 2, 1), (12, 4, 4, 2, 2), (12, 4, 4, 2, 4),
                                         (12, 4, 2, 2, 8), (12, 4, 4, 1, 8)]:
        rate = run(bw, bh, per, stages, warps=warps, iters=400)
        print("%3d x %d      %2d        %2d     %d      %6.1f       %6.1f" % (bw, bh, per, stages, warps,
                                                                          warps * per * stages * bw * bh * 128 / 1024, rate), flush=True)
if which in ("all", "grid"):
    print("\nactive SMs (36x4 boxes, 2 per stage, 3 stages)   B/clk/SM   aggregate TB/s")
    for grid in (8, 16, 37, 74, 111, 148):
        rate = run(36, 4, 2, 3, grid=grid)
        print("   %3d                                           %6.1f     %6.2f" % (grid, rate, rate * grid * 1.965e9 / 1e12), flush=True)
if which in ("all", "cluster"):
    print("\ncluster multicast (36x4 boxes, 2 per stage, 3 stages; 12x4 boxes, 8 per stage, 3 stages)")
    print("cluster  grid   box     delivered B/clk/SM   delivered chip TB/s   L2-side TB/s (delivered / cluster)")
    for (bw, bh, per, stages) in [(36, 4, 4, 2), (12, 4, 8, 3)]:
        for cs in (1, 2, 4, -2, -4):       # negative: every rank issues its share of each stage's boxes
            grid = 148 - 148 % abs(cs)
            rate = run(bw, bh, per, stages, grid=grid, cluster=cs)
            tb = rate * grid * 1.965e9 / 1e12
            print("  %2d     %3d   %2dx%d        %6.1f               %6.2f                %6.2f" % (cs, grid, bw, bh, rate, tb, tb / abs(cs)), flush=True)
