#!/usr/bin/env python
"""MLUPS of the five BASELINE.json configurations on one B200 through cuda_lbm_b200.Engine and the Python scenario mirror
(tools/pyscenarios.py), next to the reference's own CUDA solver where its binary is present (oracle/_ref/bin, built by
oracle/build_ref.sh).  Development / reporting tool; bench.py is the contract benchmark.

  python tools/config_bench.py [c1 c2 c3 c4 c5] [--steps K]
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cuda_lbm_b200 as L  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyscenarios as S  # noqa: E402

f32 = np.float32


def cfg(name):
    if name == "c1":
        return dict(nx=256, ny=256, sc=S.TaylorGreenScenario(scale=2, collision=L.BGK), steps=1000, ref="c1_tg_bgk_256", adapter=L.ADAPTER_EXACT)
    if name == "c2":
        return dict(nx=1024, ny=256, sc=S.PoiseuilleScenario(collision=L.MRT), steps=1000, ref="c2_pois_mrt_1024x256", adapter=L.ADAPTER_EXACT)
    if name == "c3":
        return dict(nx=4096, ny=4096, sc=S.LidDrivenScenario(collision=L.CM_OPTIMAL, u_max=0.1, viscosity=0.4096), steps=32,
                    ref=None, adapter=L.ADAPTER_EXACT)    # 6 + 32 steps: inside the ~40 steps for which the reference's adapter keeps this cavity finite (tools/c3_probe.py)
    if name == "c3l":
        return dict(nx=4096, ny=4096, sc=S.LidDrivenScenario(collision=L.CM_OPTIMAL, u_max=0.1, viscosity=0.4096), steps=32,
                    ref=None, adapter=L.ADAPTER_LAGGED)
    if name == "c4":
        return dict(nx=32768, ny=32768, sc=S.TaylorGreenScenario(scale=256, collision=L.BGK), steps=40, ref=None, adapter=L.ADAPTER_EXACT)
    if name == "c4s":
        return dict(nx=8192, ny=8192, sc=S.TaylorGreenScenario(scale=64, collision=L.BGK), steps=100, ref="t_tg_bgk_8192", adapter=L.ADAPTER_EXACT)
    if name == "c5":
        return dict(nx=8192, ny=2048, sc=S.FlowPastCylinderScenario(2048, collision=L.MRT, num_pts=256), steps=200,
                    ref="c5_cyl_ibm_mrt_8192x2048", adapter=L.ADAPTER_EXACT)
    if name == "c5d":
        return dict(nx=8192, ny=2048, sc=S.FlowPastCylinderScenario(2048, collision=L.MRT, num_pts=804), steps=200, ref=None, adapter=L.ADAPTER_EXACT)
    raise SystemExit("unknown config " + name)


def run(name, steps_override=None):
    c = cfg(name)
    nx, ny, sc = c["nx"], c["ny"], c["sc"]
    steps = steps_override or c["steps"]
    lbm = S.LBM(nx, ny, adapter_mode=c["adapter"])
    lbm.allocate(sc)
    if name.startswith("c4"):
        # the Init functor on the device (a 32768^2 host field would be 12.9 GB)
        lbm.engine.init_taylor_green(sc.viscosity, f32(sc.u_max) / f32(sc.scale))
    else:
        lbm.init(sc)
    eng = lbm.engine
    eng.step(6)
    eng.sync()
    l0 = eng.info().kernel_launches
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.step(steps)
    e1.record(stream)
    eng.sync()
    ms = e0.elapsed_time(e1)
    inf = eng.info()
    out = {"config": name, "scenario": sc.name(), "nx": nx, "ny": ny, "collision": ["BGK", "MRT", "CM", "CM_OPT"][sc.collision],
           "steps": steps, "ms_per_step": ms / steps, "mlups": nx * ny * steps / ms / 1e3,
           "gbs_72B": 72.0 * nx * ny * steps / ms / 1e6, "launches_per_step": (inf.kernel_launches - l0) / steps,
           "markers": inf.num_markers, "ibm_nodes": inf.num_ibm_nodes, "bytes_per_cell": inf.bytes_per_cell,
           "mass_per_cell": eng.total_mass() / (nx * ny)}
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin", c["ref"]) if c["ref"] else None
    if ref_bin and os.path.exists(ref_bin):
        try:
            r = subprocess.run([ref_bin, "60", "/tmp", "x"], capture_output=True, text=True, timeout=600)
            m = re.search(r"REF_MLUPS ([0-9.]+)", r.stdout)
            if m:
                out["reference_cuda_mlups"] = float(m.group(1))
                out["speedup_vs_reference_cuda"] = out["mlups"] / out["reference_cuda_mlups"]
        except Exception as ex:  # noqa: BLE001
            out["reference_cuda_error"] = str(ex)
    lbm.free()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c3", "c3l", "c4s", "c5", "c5d", "c4"])
    ap.add_argument("--steps", type=int, default=None)
    a = ap.parse_args()
    for n in a.configs:
        print(json.dumps(run(n, a.steps)), flush=True)
