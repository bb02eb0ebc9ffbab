#!/usr/bin/env python
"""bench.py — MLUPS of the fused D2Q9 stream+collide on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--gpus N] ...                the reference arm (CPU, see below)
  torchrun --nproc-per-node N bench.py --gpus N ...              N > 1: one rank per GPU, y-slabs

Workload (config.workload): BASELINE.json configs[3], the configuration its metric and targets are quoted on —
Taylor-Green vortex, D2Q9, BGK, 32768 x 32768, periodic, y-slab decomposed over N GPUs of one box (strong
scaling: the grid is fixed, each rank owns 32768/N rows; --rows-per-gpu R = weak scaling).  It fits one B200 (38.7 GB of populations).
A "step" is one time step of the whole grid = one launch of the fused kernel per rank (+ halo rows on odd steps).

  value    MLUPS with the populations resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e      the same K steps through the public API with HOST buffers inside the timed region: pinned host rho/u -> H2D ->
           K steps -> D2H of rho,u, i.e. one segment of the reference's driver loop (init once, update_macroscopics at the
           save interval, src/main.cu:77-147).  lbm_run_from_host per rank (one call, copies and kernels pipelined over
           row bands; peer-mapped slabs synchronise their faces level by level on the device)
  roofline 72 B per cell-update (9 fp32 reads + 9 writes) / measured kernel time vs MEASURED_PEAKS.json
  check    correctness of the state the timed region left behind, computed outside it: relative L2 error of u against the
           analytic Taylor-Green decay (the reference's metric, taylorGreenScenario.cuh:59-88; sums taken on the device and
           all-reduced over the ranks) and the mass drift since init
  cpu_baseline  the CPU oracle (oracle/lbm_oracle.c, OpenMP) on a bounded sample, rank 0
  reference_cuda (N = 1)  the reference's OWN CUDA solver (oracle/_ref/bin/t_tg_bgk_8192, built from its sources by oracle/build_ref.sh)
           on the largest Taylor-Green box it can hold, timed by its own CUDA events outside this script's timed region
  configs  (N = 1)  the other four BASELINE.json configurations through the C++ ScenarioTrait / LBM<2> header shim
           (examples/_bin/ex_c*, built from examples/main.cu) with the reference's CUDA solver on the same scenario beside each

--impl reference: the reference's algorithm on the box's host cores.  The reference's solver is CUDA-only; its
CPU form is the statement-by-statement restatement in oracle/ (kind "port"), run with all host threads on a
bounded sample (a 4096 x 4096 sub-grid of the same workload per step).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 32768
NU = 1.0 / 6.0
U0 = 0.04 / 256.0          # TaylorGreenInit: u_max / SCALE, SCALE = NX/128 (taylorGreenFunctors.cuh:11-13, defines.hpp:20-23)
BYTES_PER_UPDATE = 72.0
NCU_TRAFFIC_PER_LAUNCH = 77.260e9       # dram bytes per launch of the 32768^2 BGK step kernels, ncu --set full (profiles/r02_ncu_bench_kernel.md)
EX = os.path.join(ROOT, "examples", "_bin")
REF = os.path.join(ROOT, "oracle", "_ref", "bin")


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
    """TG BGK on an n x n periodic grid with the CPU oracle on every core this process may use.  Returns (MLUPS, threads, seconds)."""
    from oracle import oracle as O
    import numpy as np
    threads = O.set_num_threads()           # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for the host's cores explicitly
    rho, u = O.taylor_green_init(n, n, NU, 0.04 / (n / 128.0))
    o = O.Oracle(n, n, coll=O.BGK, viscosity=NU, periodic=(True, True), u_max=0.04)
    o.init(rho, u)
    o.step(warm)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    assert np.isfinite(o.macroscopics()[0]).all()
    return n * n * steps / dt / 1e6, threads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 4096
    cpu_oracle_mlups(n, max(1, args.warmup), warm=0)
    mlups, threads, dt = cpu_oracle_mlups(n, args.steps, warm=0)
    sample = (f"Taylor-Green BGK {n}x{n} sub-grid per step (bounded sample of the {NX}x{NY} workload: MLUPS is per cell), {args.steps} steps, "
              f"{threads} OpenMP threads (set explicitly; {os.cpu_count()} host cores)")
    line = {"impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"taylor_green_d2q9_bgk_{NX}x{NY}_periodic_yslab", "nx": NX, "ny": NY, "collision": "BGK",
                       "sampled_grid": [n, n]},
            "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ the reference's CUDA build and the other BASELINE configurations
def reference_cuda_mlups(binary, steps, timeout=600):
    """Runs one of the reference's own CUDA solver binaries (oracle/_ref/bin) and returns its REF_MLUPS line as a dict, or None."""
    path = os.path.join(REF, binary)
    if not os.path.exists(path):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        try:
            r = subprocess.run([path, str(steps), tmp, "x"], capture_output=True, text=True, timeout=timeout)
        except Exception as ex:  # noqa: BLE001
            return {"binary": binary, "error": str(ex)[:200]}
    m = re.search(r"REF_MLUPS ([0-9.]+) steps (\d+) ms_per_step ([0-9.]+) nx (\d+) ny (\d+)", r.stdout)
    if not m:
        return {"binary": binary, "error": (r.stdout[-200:] + r.stderr[-200:]).strip()}
    return {"binary": "oracle/_ref/bin/" + binary, "grid": [int(m.group(4)), int(m.group(5))], "mlups": float(m.group(1)),
            "ms_per_step": float(m.group(3)), "timed_steps": int(m.group(2)),
            "what": "the reference's own translation units (src/core, src/IBM) compiled for sm_100a by oracle/build_ref.sh; median-free mean of its per-step CUDA-event times"}


# name -> (shim binary, reference binary, nx, ny, operator, warm-up steps, timed steps, reference steps, description); c3 repeats its
# (init, warm-up, timed) segment C3_REPEAT times
CONFIGS = {
    "c1": ("ex_c1_tg_256", "c1_tg_bgk_256", 256, 256, "BGK", 64, 4000, 200, "Taylor-Green 256x256 BGK (configs[0])"),
    "c2": ("ex_c2_pois_1024x256", "c2_pois_mrt_1024x256", 1024, 256, "MRT", 64, 4000, 200, "Poiseuille 1024x256 MRT, body force, bounce-back walls (configs[1])"),
    # 18 + 17 steps: the reference's adapter keeps this cavity physical for ~40 steps (tools/c3_probe.py, profiles/r02_c3_probe.txt); the
    # warm-up covers module load and the capture of the 16-step CUDA graph that the timed steps replay (same parity)
    "c3": ("ex_c3_lid_4096", "c3_lid_cmopt_4096", 4096, 4096, "CM<OptimalAdapter> (exact grid means)", 18, 17, 30,
           "lid-driven cavity 4096x4096 CM<OptimalAdapter> (configs[2]); run inside the window in which the reference's adapter keeps the field finite"),
    # lagged sums: the first step takes the sums by the pre-pass, the graph of the steady state is captured at step 2 — 19 warm-up steps put
    # the timed steps on the same parity (18 would put one graph capture, ~2.5 ms, into the first of the six timed segments)
    "c3_lagged": ("ex_c3_lid_4096", None, 4096, 4096, "CM<OptimalAdapter> (grid means of the previous step)", 19, 17, 0,
                  "configs[2] with LBM_ADAPTER_LAGGED (72 B/cell; deviation from the exact mode: tests/test_reference_fullsize_gpu.py)"),
    "c5": ("ex_c5_cyl_8192x2048", "c5_cyl_ibm_mrt_8192x2048", 8192, 2048, "MRT + IBM (256 markers)", 32, 400, 60,
           "flow past cylinder 8192x2048 MRT, IBM direct forcing (configs[4])"),
}


C3_REPEAT = 6


def run_config(name, peak):
    shim, ref, nx, ny, op, warm, steps, ref_steps, what = CONFIGS[name]
    path = os.path.join(EX, shim)
    if not os.path.exists(path):
        return {"name": name, "unavailable": f"examples/_bin/{shim} not built"}
    env = dict(os.environ)
    if name == "c3_lagged":
        env["LBM_B200_ADAPTER"] = "1"
    with tempfile.TemporaryDirectory() as tmp:
        try:
            rep = C3_REPEAT if name.startswith("c3") else 1
            r = subprocess.run([path, "--steps", str(steps), "--save-int", str(steps), "--warmup", str(warm), "--repeat", str(rep), "--fast"], cwd=tmp, env=env,
                               capture_output=True, text=True, timeout=600)
        except Exception as ex:  # noqa: BLE001
            return {"name": name, "error": str(ex)[:200]}
    m = re.search(r"SHIM_RESULT (.*)", r.stdout)
    if not m:
        return {"name": name, "error": (r.stdout[-300:] + r.stderr[-300:]).strip()}
    res = dict(kv.split("=") for kv in m.group(1).split())
    mlups, mass = float(res["mlups"]), float(res["mass_per_node"])
    cells = nx * ny
    out = {"name": name, "workload": what, "nx": nx, "ny": ny, "collision": op, "warmup_steps": warm, "steps": steps, "segments": rep,
           "mlups": mlups, "ms_per_step": float(res["ms_per_step"]),
           "physical": bool(abs(mass - 1.0) < 0.05 and float(res["sum_u2"]) == float(res["sum_u2"])), "mass_per_node": mass,
           "api": "examples/main.cu over the LBM<2> / ScenarioTrait header shim (include/cuda-lbm), LBM::run<Scenario>(n)"}
    # the HBM roofline applies where the populations (36 B/cell) do not fit the 126 MB L2
    out["roofline_frac"] = (BYTES_PER_UPDATE * mlups * 1e6 / 1e9 / peak) if 36.0 * cells > 4 * 126e6 else None
    if out["roofline_frac"] is None:
        out["roofline_note"] = "populations are L2-resident: launch / latency bound, no HBM roofline claim"
    if ref:
        rc = reference_cuda_mlups(ref, ref_steps)
        out["reference_cuda"] = rc
        if rc and "mlups" in rc:
            out["speedup_vs_reference_cuda"] = mlups / rc["mlups"]
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the N = 1 legs that run the other BASELINE configurations and the reference's CUDA build")
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
    u0 = 0.04 / scale
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

    def reduce_ranks(vals, op):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return [float(v) for v in t.tolist()]

    def max_over_ranks(v):
        return reduce_ranks([v], dist.ReduceOp.MAX if world > 1 else None)[0]

    def sum_over_ranks(vals):
        return reduce_ranks(vals, dist.ReduceOp.SUM if world > 1 else None)

    # ---------------- device-resident measurement ----------------
    eng.init_taylor_green(NU, u0)
    solver.barrier_after_init()
    mass0 = sum_over_ranks([eng.total_mass()])[0]
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
    mlups = nx * ny * args.steps / (ms * 1e-3) / 1e6
    # dominant kernel = the fused step kernel: one launch per step per rank over nloc cells
    kern_ms = ms / args.steps
    achieved = BYTES_PER_UPDATE * nloc / (kern_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # dram__bytes_read.sum + dram__bytes_write.sum per launch: a CITATION of the ncu --set full capture of this command at N = 1
    # (profiles/r02_ncu_bench_kernel.md, mean of the odd- and even-phase kernels), scaled by the rows this rank owns; not re-measured per run
    traffic = NCU_TRAFFIC_PER_LAUNCH * (nloc / float(NX * NY)) if (nx == NX and args.collision == "BGK") else None

    # ---------------- correctness of the state the timed region left behind (outside the timed region) ----------------
    # one more step that also stores rho / u, then the reference's Taylor-Green metric with both sums taken on the device
    # (lbm_taylor_green_error_sums, taylorGreenScenario.cuh:59-88) and the mass, all-reduced over the ranks
    solver.step(1, macroscopics=True)
    t_total = args.warmup + args.steps + 1
    s = sum_over_ranks(list(eng.taylor_green_error_sums(NU, u0, float(t_total))))
    mass1 = sum_over_ranks([eng.total_mass()])[0]
    finite = bool(np.isfinite(s).all() and np.isfinite(mass1) and s[1] > 0)
    check = {"finite": finite, "steps_total": t_total,
             "tg_l2_error_pct": (100.0 * (s[0] / s[1]) ** 0.5) if finite else None,
             "mass_per_cell": mass1 / (float(nx) * ny), "mass_rel_drift": mass1 / mass0 - 1.0,
             "how": "lbm_taylor_green_error_sums + lbm_total_mass after the timed region, summed over the ranks (fp64); the reference's stale rest population (A-D1, reproduced) makes mass drift at the 1e-6 level"}
    assert finite, "non-finite state after the timed region"

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        from cuda_lbm_b200._capi import lib, check as chk
        import ctypes as C
        # pinned host buffers on the NUMA node of this rank's GPU (lbm_host_alloc), separate ones for the input and the result
        h_rho, h_u, o_rho, o_u = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        chk(lib().lbm_host_alloc(C.byref(h_rho), nloc * 4))
        chk(lib().lbm_host_alloc(C.byref(h_u), nloc * 8))
        chk(lib().lbm_host_alloc(C.byref(o_rho), nloc * 4))
        chk(lib().lbm_host_alloc(C.byref(o_u), nloc * 8))
        # host-resident input: the Init functor's rho,u for this slab, produced once outside the timed region
        chk(lib().lbm_reserve_macroscopics(eng._h))
        eng.init_taylor_green(NU, u0)
        eng.macroscopics_into(h_rho.value, h_u.value)
        solver.barrier_after_init()
        barrier()
        t0 = time.perf_counter()
        # one call per rank = one driver segment: H2D of the inputs, the steps and D2H of the result overlap band by band
        # (lbm_run_from_host; peer-mapped slabs synchronise their faces level by level on the device)
        api, failed = None, 0.0
        try:
            solver.run_from_host(h_rho.value, h_u.value, args.steps, o_rho.value, o_u.value)
            api = ("lbm_run_from_host (copies and kernels pipelined over row bands)" if solver.mode in ("single", "direct")
                   else "lbm_init_fields_local + lbm_step_with_macroscopics + lbm_get_macroscopics")
        except L.LbmError as ex:        # never lose the bench line over the e2e leg: every rank falls back together
            failed, api = 1.0, f"fallback after: {ex}"
        if max_over_ranks(failed) > 0:
            eng.init_taylor_green(NU, u0)
            eng.macroscopics_into(h_rho.value, h_u.value)
            solver.barrier_after_init()
            barrier()
            t0 = time.perf_counter()
            chk(lib().lbm_init_fields_local(eng._h, h_rho, h_u))
            solver.barrier_after_init()
            solver.step(args.steps, macroscopics=True)
            eng.macroscopics_into(o_rho.value, o_u.value)
            api = "lbm_init_fields_local + lbm_step_with_macroscopics + lbm_get_macroscopics (" + (api or "another rank's lbm_run_from_host failed") + ")"
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        # the result in the HOST buffers is what the user gets: check it too (mean of rho over this rank's rows, summed over ranks)
        h_sum = float(np.ctypeslib.as_array(C.cast(o_rho, C.POINTER(C.c_float)), shape=(nloc,)).sum(dtype=np.float64))
        h_mean = sum_over_ranks([h_sum])[0] / (float(nx) * ny)
        e2e = {"value": nx * ny * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": 12.0 * nloc / args.steps, "d2h_bytes_per_step": 12.0 * nloc / args.steps,
               "segment": f"pinned host rho,u -> H2D -> {args.steps} steps -> D2H rho,u ({12 * nloc / 1e9:.2f} GB each way per rank; separate input and output buffers, allocated on the GPU's NUMA node)", "api": api,
               "seconds": dt, "host_result_mean_rho": h_mean}
        for b in (h_rho, h_u, o_rho, o_u):
            lib().lbm_host_free(b)
    eng.close()

    cpu = None
    if rank == 0 and not args.no_cpu:
        n = 2048
        v, threads, dt = cpu_oracle_mlups(n, 60)
        cpu = {"value": v, "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": f"CPU oracle (oracle/lbm_oracle.c, OpenMP x{threads}), Taylor-Green BGK {n}x{n}, 60 steps, {dt:.1f} s"}

    # ---------------- N = 1: the reference's CUDA solver and the other BASELINE configurations, outside every timed region ----------------
    ref_cuda, configs = None, None
    if rank == 0 and world == 1 and not args.no_configs:
        ref_cuda = reference_cuda_mlups("t_tg_bgk_8192", 40)
        if ref_cuda and "mlups" in ref_cuda:
            ref_cuda["note"] = ("Taylor-Green BGK on the largest square box the reference holds comfortably (int indexing and 144 B/cell stop it near 15000^2; "
                                "32768^2 is out of its reach); MLUPS is per cell, so the figure compares directly with `value`")
            ref_cuda["speedup_of_value"] = mlups / ref_cuda["mlups"]
        configs = [run_config(name, peak) for name in ("c1", "c2", "c3", "c3_lagged", "c5")]

    if rank == 0:
        line = {"metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"taylor_green_d2q9_{args.collision.lower()}_{nx}x{ny}_periodic_yslab", "nx": nx, "ny": ny,
                           "collision": args.collision, "rows_per_gpu": eng.ny_local, "quirks": "reference-compatible", "slab_coupling": solver.mode,
                           "l2_policy": f"populations per GPU {36.0 * nloc / 1e9:.1f} GB >> 126 MB L2 (no flush needed)"},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "check": check,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic,
                             "traffic_source": "citation: ncu --set full of this command at N = 1, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the odd / even phase captures (profiles/r02_ncu_bench_kernel.md), scaled by rows per rank; not re-measured by this run",
                             "peak_source": peak_src, "kernel": "lbm::step_odd_kernel<BGK> (odd AA phase) / lbm::step_vec_kernel<BGK, even> — fused pull + collide + push, 4 cells per thread, one launch per step",
                             "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * nloc, "kernel_ms": kern_ms},
                "cpu_baseline": cpu, "reference_cuda": ref_cuda, "configs": configs}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
