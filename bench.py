#!/usr/bin/env python
"""bench.py — MLUPS of the fused D2Q9 stream+collide on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...                the reference arm (CPU, see below)
  torchrun --nproc-per-node N bench.py --gpus N ...              N > 1: one rank per GPU, y-slabs

Workload (config.workload): BASELINE.json configs[3], the configuration its metric and targets are quoted on —
Taylor-Green vortex, D2Q9, BGK, 32768 x 32768, periodic, y-slab decomposed over N GPUs of one box (strong
scaling: the grid is fixed, each rank owns 32768/N rows).  It fits one B200 (38.7 GB of populations).
A "step" is one time step of the whole grid = one launch of the fused kernel per rank (+ halo rows on odd steps).

  value    MLUPS with the populations resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e      the same K steps through the public API with HOST buffers inside the timed region: pinned host rho/u -> H2D ->
           K steps -> D2H of rho,u, i.e. one segment of the reference's driver loop (init once, update_macroscopics at the
           save interval, src/main.cu:77-147).  lbm_run_from_host per rank (one call, copies and kernels pipelined over
           row bands; peer-mapped slabs synchronise their faces level by level on the device)
  roofline 72 B per cell-update (9 fp32 reads + 9 writes) / measured kernel time vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/lbm_oracle.c, OpenMP) on a bounded sample, rank 0, N = 1 only

--impl reference: the reference's algorithm on the box's host cores.  The reference's solver is CUDA-only; its
CPU form is the statement-by-statement restatement in oracle/ (kind "port"), run with all host threads on a
bounded sample (a 4096 x 4096 sub-grid of the same workload per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 32768
NU = 1.0 / 6.0
U0 = 0.04 / 256.0          # TaylorGreenInit: u_max / SCALE, SCALE = NX/128 (taylorGreenFunctors.cuh:11-13, defines.hpp:20-23)
BYTES_PER_UPDATE = 72.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_mlups(n, steps, warm=1):
    """TG BGK on an n x n periodic grid with the CPU oracle, all OpenMP threads.  Returns (MLUPS, threads, seconds)."""
    from oracle import oracle as O
    import numpy as np
    rho, u = O.taylor_green_init(n, n, NU, 0.04 / (n / 128.0))
    o = O.Oracle(n, n, coll=O.BGK, viscosity=NU, periodic=(True, True), u_max=0.04)
    o.init(rho, u)
    o.step(warm)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    assert np.isfinite(o.macroscopics()[0]).all()
    return n * n * steps / dt / 1e6, O.num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 4096
    mlups_w, threads, _ = cpu_oracle_mlups(n, max(1, args.warmup), warm=0)
    mlups, threads, dt = cpu_oracle_mlups(n, args.steps, warm=0)
    sample = f"Taylor-Green BGK {n}x{n} sub-grid per step (bounded sample of the {NX}x{NY} workload), {args.steps} steps, {threads} OpenMP threads"
    line = {"impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"taylor_green_d2q9_bgk_{NX}x{NY}_periodic_yslab", "nx": NX, "ny": NY, "collision": "BGK",
                       "sampled_grid": [n, n]},
            "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--ny", type=int, default=NY)
    ap.add_argument("--collision", default="BGK")
    ap.add_argument("--rows-per-gpu", type=int, default=0,
                    help="weak scaling: ny = rows-per-gpu * N (BASELINE configs[3]: 32768 x 4096*G slabs); default 0 = strong scaling on the fixed grid")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import cuda_lbm_b200 as L
    from cuda_lbm_b200.slab import SlabSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx, ny = args.nx, args.ny
    scaling = "strong"
    if args.rows_per_gpu > 0:
        ny, scaling = args.rows_per_gpu * world, "weak"
    coll = {"BGK": L.BGK, "MRT": L.MRT, "CM": L.CM, "CM_OPT": L.CM_OPTIMAL}[args.collision]
    scale = nx / 128.0
    eng = L.Engine(nx, ny, collision=coll, viscosity=NU, periodic=(True, True), u_max=0.04, device=local, rank=rank, world=world,
                   adapter_mode=L.ADAPTER_LAGGED)
    # one explicit stream for the engine's kernels, the NCCL halo traffic and the timing events
    # (torch's legacy default stream has handle 0, which lbm_set_stream reads as "use the handle's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    slab_mode = os.environ.get("LBM_SLAB_MODE", "direct")        # "direct": peer-mapped edge rows over NVLink; "nccl": packed halo rows
    solver = SlabSolver(eng, nx, True, dev, optimal_adapter=(coll == L.CM_OPTIMAL), adapter_exact=False, mode=slab_mode)
    nloc = eng.ny_local * nx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident measurement ----------------
    eng.init_taylor_green(NU, 0.04 / scale)
    solver.barrier_after_init()
    solver.step(args.warmup)
    barrier()
    l0 = eng.info().kernel_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    solver.step(args.steps)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    launches = eng.info().kernel_launches - l0
    mass = eng.total_mass()
    assert np.isfinite(mass), "non-finite state after the timed region"
    mlups = nx * ny * args.steps / (ms * 1e-3) / 1e6
    # dominant kernel = the fused step kernel: one launch per step per rank over nloc cells
    kern_ms = ms / args.steps
    achieved = BYTES_PER_UPDATE * nloc / (kern_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this very command
    # (profiles/r01_ncu_bench_kernel.md, second capture: 77.261 GB odd phase, 77.259 GB even phase); other sizes were not captured
    traffic = 77.260e9 * (nloc / float(NX * NY)) if (nx == NX and args.collision == "BGK") else None

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        from cuda_lbm_b200._capi import lib, check
        import ctypes as C
        h_rho, h_u = C.c_void_p(), C.c_void_p()
        check(lib().lbm_host_alloc(C.byref(h_rho), nloc * 4))
        check(lib().lbm_host_alloc(C.byref(h_u), nloc * 8))
        # host-resident input: the Init functor's rho,u for this slab, produced once outside the timed region
        check(lib().lbm_reserve_macroscopics(eng._h))
        eng.init_taylor_green(NU, 0.04 / scale)
        eng.macroscopics_into(h_rho.value, h_u.value)
        solver.barrier_after_init()
        barrier()
        t0 = time.perf_counter()
        # one call per rank = one driver segment: H2D of the inputs, the steps and D2H of the result overlap band by band
        # (lbm_run_from_host; peer-mapped slabs synchronise their faces level by level on the device)
        api, failed = None, 0.0
        try:
            solver.run_from_host(h_rho.value, h_u.value, args.steps, h_rho.value, h_u.value)
            api = ("lbm_run_from_host (copies and kernels pipelined over row bands)" if solver.mode in ("single", "direct")
                   else "lbm_init_fields_local + lbm_step_with_macroscopics + lbm_get_macroscopics")
        except L.LbmError as ex:        # never lose the bench line over the e2e leg: every rank falls back together
            failed, api = 1.0, f"fallback after: {ex}"
        if max_over_ranks(failed) > 0:
            eng.init_taylor_green(NU, 0.04 / scale)
            eng.macroscopics_into(h_rho.value, h_u.value)
            solver.barrier_after_init()
            barrier()
            t0 = time.perf_counter()
            check(lib().lbm_init_fields_local(eng._h, h_rho, h_u))
            solver.barrier_after_init()
            solver.step(args.steps, macroscopics=True)
            eng.macroscopics_into(h_rho.value, h_u.value)
            api = "lbm_init_fields_local + lbm_step_with_macroscopics + lbm_get_macroscopics (" + (api or "another rank's lbm_run_from_host failed") + ")"
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": nx * ny * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": 12.0 * nloc / args.steps, "d2h_bytes_per_step": 12.0 * nloc / args.steps,
               "segment": f"pinned host rho,u -> H2D -> {args.steps} steps -> D2H rho,u ({12 * nloc / 1e9:.2f} GB each way per rank)", "api": api,
               "seconds": dt}
        lib().lbm_host_free(h_rho)
        lib().lbm_host_free(h_u)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = 2048
        v, threads, dt = cpu_oracle_mlups(n, 60)
        cpu = {"value": v, "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": f"CPU oracle (oracle/lbm_oracle.c, OpenMP x{threads}), Taylor-Green BGK {n}x{n}, 60 steps, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"taylor_green_d2q9_{args.collision.lower()}_{nx}x{ny}_periodic_yslab", "nx": nx, "ny": ny,
                           "collision": args.collision, "rows_per_gpu": eng.ny_local, "quirks": "reference-compatible", "slab_coupling": solver.mode,
                           "l2_policy": f"populations per GPU {36.0 * nloc / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": "ncu dram bytes per launch, mean of the odd/even phase captures (profiles/r01_ncu_bench_kernel.md)",
                             "peak_source": peak_src, "kernel": "lbm::step_vec_kernel<BGK, odd|even> (fused pull + collide + push, 4 cells/thread)",
                             "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * nloc, "kernel_ms": kern_ms},
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
