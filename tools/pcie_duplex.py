#!/usr/bin/env python
"""PCIe copy bandwidth of the box: host->device alone, device->host alone, both at once (two streams, pinned memory).
Context for bench.py's e2e figure: lbm_run_from_host overlaps the two directions, so its floor is the duplex figure.

  python tools/pcie_duplex.py                                         one GPU
  torchrun --nproc-per-node 8 tools/pcie_duplex.py                    all GPUs of the box AT THE SAME TIME: what the ranks of an e2e run share
"""
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = (1 << 30) if world == 1 else (1 << 28)        # floats per direction: 4 GiB alone, 1 GiB per rank when the ranks run together
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float32, device="cuda"), torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=3):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:                                   # the slowest rank sets the pace of a joint transfer
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


for _ in range(2):
    t_up, t_down = timed(h2d), timed(d2h)
    t_both = timed(lambda: (h2d(), d2h()))
    gb = 4 * n / 1e9
    if rank == 0:
        print(f"{world} GPU(s) at once, per GPU: H2D {gb / t_up:.1f} GB/s, D2H {gb / t_down:.1f} GB/s, both directions {gb / t_both:.1f} GB/s each; "
              f"box aggregate: H2D {world * gb / t_up:.1f}, D2H {world * gb / t_down:.1f}, duplex {2 * world * gb / t_both:.1f} GB/s total", flush=True)
if world > 1:
    dist.destroy_process_group()
