"""Packs the raw dumps of the reference CUDA solver (gpurun_out/golden/*.bin, produced on a B200 by
tests/golden/make_golden.sh) into the committed fixtures tests/golden/<case>.npz.

Per case: f_t<k> (post-collision populations, [ny,nx,9]) for k in case.steps_f, rho_t<k> / u_t<k>
(the reference's d_rho / d_u after step k) for k in case.steps_m and k = 0.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import CASES  # noqa: E402


def main(src):
    for c in CASES:
        out = {}
        n = c.nx * c.ny

        def load(step, kind, cnt):
            p = os.path.join(src, f"{c.name}_t{step}.{kind}.bin")
            a = np.fromfile(p, np.float32)
            assert a.size == cnt, (p, a.size, cnt)
            return a

        try:
            for k in c.steps_f:
                out[f"f_t{k}"] = load(k, "f", 9 * n).reshape(c.ny, c.nx, 9)
            for k in (0,) + tuple(c.steps_m):
                out[f"rho_t{k}"] = load(k, "rho", n).reshape(c.ny, c.nx)
                out[f"u_t{k}"] = load(k, "u", 2 * n).reshape(c.ny, c.nx, 2)
        except (FileNotFoundError, AssertionError) as e:
            print("skip", c.name, e)
            continue
        np.savez_compressed(os.path.join(HERE, c.name + ".npz"), **out)
        print("packed", c.name, {k: v.shape for k, v in list(out.items())[:2]})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(HERE)), "gpurun_out", "golden"))
