"""Prints the engine-vs-oracle deviation table for every parity case (run on the GPU box; no assertions)."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import cases
from test_parity_gpu import _run_pair

if __name__ == "__main__":
    for c in cases.CASES:
        for qk in (63, 0):
            try:
                steps = [1, 2, 3, 4, 10, 50] if c.coll != cases.CM_OPT else [1, 2, 3, 4, 10, 20]
                res = _run_pair(c, steps, quirks=qk)
                print(f"{c.name:16s} quirks={qk:2d} " + " | ".join(f"t{n}: df={df:.1e} drho={dr:.1e} du={du:.1e}{'' if fin else ' NONFINITE'}" for n, df, dr, du, fin in res), flush=True)
            except Exception:
                print(c.name, qk, "EXCEPTION"); traceback.print_exc()
