#!/usr/bin/env python3
"""Pinned-host <-> device copy bandwidth, alone and with every rank copying at once, with the staging buffers on the
GPU's own NUMA node vs. wherever torchrun happened to start the rank.

    python tools/hostcopy_probe.py                      # 1 GPU: each NUMA node in turn
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/hostcopy_probe.py

Prints one JSON line (rank 0): per-rank H2D / D2H / simultaneous GB/s for policy = "none" (no binding), "local"
(flownet2_b200.numa.bind_to_device_node) and, single-GPU only, every explicit node.
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch

spec = importlib.util.spec_from_file_location("_fn2_numa", os.path.join(ROOT, "flownet2-pytorch_b200", "numa.py"))
numa = importlib.util.module_from_spec(spec)
spec.loader.exec_module(numa)

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 512 << 20
ALL_CPUS = sorted(os.sched_getaffinity(0))


def measure():
    h_in = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    d_out = torch.ones(NBYTES, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def timed(fn, reps=6):
        fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def h2d():
        s1.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()
    res["h2d_GBps"] = round(NBYTES / timed(h2d) / 1e6, 1)
    res["d2h_GBps"] = round(NBYTES / timed(d2h) / 1e6, 1)
    res["duplex_GBps_each_way"] = round(NBYTES / timed(both) / 1e6, 1)
    del h_in, h_out
    return res


out = {"world": world, "gpu_node": None, "nodes": numa.online_nodes(), "policies": {}}
try:
    out["gpu_node"] = numa.device_node(local)
except Exception as e:
    out["gpu_node_error"] = str(e)[:100]
policies = [("none", None)]
if world == 1:
    policies += [("node%d" % n, n) for n in numa.online_nodes()]
policies.append(("local", "local"))
for name, node in policies:
    os.sched_setaffinity(0, ALL_CPUS)
    if node == "local":
        info = numa.bind_to_device_node(local)
    elif node is not None:
        info = numa.bind_to_node(node)
    else:
        info = {"node": None}
    r = measure()
    r["bind"] = info
    if dist:
        t = torch.tensor([r["h2d_GBps"], r["d2h_GBps"], r["duplex_GBps_each_way"]], device=dev, dtype=torch.float64)
        lo, sm = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        r = {"min_over_ranks": [round(float(x), 1) for x in lo], "sum_over_ranks": [round(float(x), 1) for x in sm],
             "rank0": r, "order": ["h2d", "d2h", "duplex_each_way"]}
    out["policies"][name] = r
if rank == 0:
    print(json.dumps(out))
if dist:
    dist.barrier()
    dist.destroy_process_group()
