#!/usr/bin/env python3
"""Measure pinned-memory PCIe bandwidth on this box: H2D alone, D2H alone, both concurrently (the ceiling of bench.py's e2e).

Single process:            python tools/pcie_peak.py
All GPUs of the box at once (the box's AGGREGATE host-side ceiling, which bounds the N-GPU e2e figure):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_peak.py [--json out.json]
Every rank copies 1 GiB buffers to / from its own GPU between barriers; rank 0 prints per-rank and summed GB/s."""
import argparse
import json
import os

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--bind", action="store_true", help="bind each rank to the CPUs NVML reports as local to its GPU")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")

affinity = "unbound"
if args.bind:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i for i in range(os.cpu_count()) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            affinity = "%d cpus" % len(cpus)
    except Exception as e:       # noqa: BLE001 - diagnostics only
        affinity = "bind failed: %s" % e

n = args.mib << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 1)
res = {"h2d": run(True, False), "d2h": run(False, True), "both_each_way": run(True, True)}
if world > 1:
    allres = [None] * world
    dist.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    table = {"ranks": world, "mib_per_copy": args.mib, "affinity": affinity, "per_rank": allres,
             "sum": {k: sum(r[k] for r in allres) for k in res}}
    for k in res:
        print("%-14s per rank %s  | sum %.1f GB/s" % (k, " ".join("%.1f" % r[k] for r in allres), table["sum"][k]))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(table, f, indent=1)
