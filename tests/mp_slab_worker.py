"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), y-slabs through cuda_lbm_b200.slab.SlabSolver.

Runs a scenario slab-decomposed over WORLD_SIZE GPUs (mode "direct": peer-mapped neighbours over NVLink, CUDA IPC between
the processes; mode "nccl": packed halo rows and IBM node states through NCCL), gathers rho, u on rank 0 and compares them
with the single-handle run of the same scenario on GPU 0.  Bar: bit-identical (OptimalAdapter: 2e-7).  Exit code 0 = parity.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402
from cuda_lbm_b200.slab import SlabSolver  # noqa: E402


def build_case(kind, coll):
    if kind == "tg":
        return cases.Case("mp_tg", 512, 384, coll, 1.0 / 6.0, (True, True), 0.04, "tg", scale=4)
    from oracle import oracle as O
    c = cases.Case("mp_ibm", 512, 256, coll, cases._cyl_nu(256), (False, False), 0.05, "cyl_ibm")
    # one cylinder across the face between the two middle slabs, a second one overlapping it, a third one elsewhere
    c.bodies = [O.create_cylinder(96.0, 128.0, 16.0, 64), O.create_cylinder(110.0, 120.5, 6.0, 24), O.create_cylinder(300.5, 40.25, 8.0, 40)]
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="ibm")
    ap.add_argument("--mode", default="direct")
    ap.add_argument("--coll", type=int, default=cases.MRT)
    ap.add_argument("--steps", type=int, default=21)
    ap.add_argument("--adapter", type=int, default=0, help="0 = LBM_ADAPTER_EXACT, 1 = LBM_ADAPTER_LAGGED (OptimalAdapter only)")
    ap.add_argument("--from-host", action="store_true", help="one lbm_run_from_host call per rank instead of init + steps + read-back")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    case = build_case(a.kind, a.coll)
    rho0, u0 = case.init_fields()
    e = cases.make_engine(case, adapter_mode=a.adapter, device=local, rank=rank, world=world)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    e.set_stream(stream.cuda_stream)
    s = SlabSolver(e, case.nx, case.periodic[1], dev, optimal_adapter=(case.coll == cases.CM_OPT), adapter_exact=(a.adapter == 0), mode=a.mode)
    if a.from_host:
        # pinned host slabs in, pinned host slabs out: the band pipeline of every rank, slab faces synchronised on the device
        sl = slice(e.y0, e.y0 + e.ny_local)
        h_rho, h_u = torch.from_numpy(np.ascontiguousarray(rho0[sl])).pin_memory(), torch.from_numpy(np.ascontiguousarray(u0[sl])).pin_memory()
        o_rho, o_u = torch.empty_like(h_rho).pin_memory(), torch.empty_like(h_u).pin_memory()
        dist.barrier()
        s.run_from_host(h_rho.data_ptr(), h_u.data_ptr(), a.steps, o_rho.data_ptr(), o_u.data_ptr())
        rho, u = o_rho.numpy(), o_u.numpy()
    else:
        e.init_fields(rho0, u0)
        s.barrier_after_init()
        s.step(a.steps, macroscopics=True)
        e.sync()
        rho, u = e.macroscopics()
    # gather the slabs on rank 0 (equal row counts: ny % world == 0 for both cases)
    t = torch.from_numpy(np.concatenate([rho[..., None], u], axis=-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    ok = True
    if rank == 0:
        full = torch.cat(parts, dim=0).cpu().numpy()
        one = cases.make_engine(case, adapter_mode=a.adapter, device=local)
        one.init_fields(rho0, u0)
        one.step(a.steps, macroscopics=True)
        r1, u1 = one.macroscopics()
        one.close()
        d_rho, d_u = float(np.abs(full[..., 0] - r1).max()), float(np.abs(full[..., 1:] - u1).max())
        tol = 0.0 if case.coll != cases.CM_OPT else 2e-7
        ok = bool(np.isfinite(full).all()) and d_rho <= tol and d_u <= tol
        print(f"MP_PARITY kind={a.kind} from_host={a.from_host} mode={s.mode} world={world} coll={a.coll} steps={a.steps} max|drho|={d_rho:.3e} max|du|={d_u:.3e} "
              f"collectives={s.collectives} {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    e.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
