#!/usr/bin/env python
"""Concurrent host<->device copy ceiling over ALL GPUs of the box: one process per GPU (torchrun), every rank copies
2 GiB each way from / to its own pinned buffers at the same time; rank 0 prints per-GPU and aggregate GB/s.  Answers
whether the e2e (host-buffer) scaling of bench.py is bound by the host side.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/pcie_probe_concurrent.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total, chunk = 2 << 30, 32 << 20
h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
d_a = torch.empty(total, dtype=torch.uint8, device="cuda")
d_b = torch.empty(total, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for off in range(0, total, chunk):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


run(True, True, 1)
a, b, c = run(True, False), run(False, True), run(True, True)
if rank == 0:
    g = total / 1e9
    try:
        numa = open("/sys/devices/system/node/online").read().strip()
    except Exception:
        numa = "?"
    print(f"{world} GPU(s) concurrently, 2 GiB per direction per GPU, 32 MiB chunks, host cores {os.cpu_count()}, NUMA nodes online {numa}")
    print(f"  H2D only : {g / a:6.1f} GB/s per GPU, {world * g / a:7.1f} GB/s aggregate")
    print(f"  D2H only : {g / b:6.1f} GB/s per GPU, {world * g / b:7.1f} GB/s aggregate")
    print(f"  both     : {g / c:6.1f} GB/s per GPU each way, {2 * world * g / c:7.1f} GB/s aggregate (in + out)")
if world > 1:
    dist.destroy_process_group()
