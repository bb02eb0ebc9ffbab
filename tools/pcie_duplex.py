#!/usr/bin/env python
"""PCIe copy bandwidth of the box: host->device alone, device->host alone, both at once (two streams, pinned memory).
Context for bench.py's e2e figure: lbm_run_from_host overlaps the two directions, so its floor is the duplex figure."""
import torch

n = 1 << 30                                   # 4 GiB of float32 per direction
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.float32, device="cuda"), torch.zeros(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


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
    print(f"H2D {gb / t_up:.1f} GB/s, D2H {gb / t_down:.1f} GB/s, both at once {2 * gb / t_both:.1f} GB/s total ({gb / t_both:.1f} per direction)")
