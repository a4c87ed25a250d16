"""Concurrent host-to-device copy probe: every rank (one process per GPU, torchrun) copies a
page-locked buffer to its GPU at the same time; rank 0 prints the per-rank and aggregate rates.
Run with and without NUMA-local CPU affinity (PKM_PROBE_AFFINITY=1) to find the box's ceiling for
the `e2e` leg of bench.py, whose input (20.8 GB per cube) is PCIe bound.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popkinmocks_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
aff = None
if os.environ.get("PKM_PROBE_AFFINITY", "0") == "1":
    aff = sharding.bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gb = float(os.environ.get("PKM_PROBE_GB", "2"))
n = int(gb * (1 << 30) // 8)
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
host.fill_(1.0)
devbuf = torch.empty(n, dtype=torch.float64, device=dev)
out = {}
for direction in ("h2d", "d2h"):
    for _ in range(2):
        (devbuf.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(devbuf, non_blocking=True))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        (devbuf.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(devbuf, non_blocking=True))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rate = torch.tensor([reps * n * 8 / dt / 1e9], dtype=torch.float64, device=dev)
    rates = [torch.zeros_like(rate) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, rate)
    else:
        rates = [rate]
    out[direction] = [round(float(r), 2) for r in rates]
if rank == 0:
    print(json.dumps({"n_gpus": world, "affinity": aff is not None, "cpus_rank0": aff, "gb_per_copy": gb,
                      "h2d_GBps_per_rank": out["h2d"], "h2d_GBps_total": round(sum(out["h2d"]), 1),
                      "d2h_GBps_per_rank": out["d2h"], "d2h_GBps_total": round(sum(out["d2h"]), 1)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
