#!/usr/bin/env python
"""Development aid: how long does BASELINE config 3 (4096^2 cavity, CM<2,OptimalAdapter>) stay finite?  Steps the engine in
chunks and prints the mass per cell and the extreme densities; both adapter modes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for mode in (0, 1):
    case = cases.Case("c3", n, n, cases.CM_OPT, 0.1 * n / 1000.0, (False, False), 0.1, "lid")
    e = cases.make_engine(case, quirks=127, adapter_mode=mode)
    e.init_fields(*case.init_fields())
    done = 0
    for chunk in [10] * 10 + [50] * 8 + [100] * 5:
        e.step(chunk, macroscopics=True)
        done += chunk
        rho, u = e.macroscopics()
        fin = bool(np.isfinite(rho).all() and np.isfinite(u).all())
        print(f"C3_PROBE n={n} adapter={'exact' if mode == 0 else 'lagged'} step={done} finite={fin} mass/N={e.total_mass() / (n * n):.6f} "
              f"rho[min,max]=[{np.nanmin(rho):.4f},{np.nanmax(rho):.4f}] max|u|={np.nanmax(np.abs(u)):.4f} avg={e.moment_avg()}", flush=True)
        if not fin:
            break
    e.close()
