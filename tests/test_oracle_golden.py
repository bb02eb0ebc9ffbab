"""CPU: the oracle (oracle/lbm_oracle.c) against the golden vectors in tests/golden/*.npz.

The golden vectors are outputs of the REFERENCE'S OWN CUDA SOLVER (oracle/build_ref.sh builds it from
/root/reference with the SURVEY.md Appendix-B build patches; tests/golden/make_golden.sh ran it on a B200;
tests/golden/pack_golden.py packed the dumps).  They pin the oracle: same Scenario structs, same step order,
state dumped after init and after steps 1, 2, 3, 10, 100 (30 for OptimalAdapter).

Tolerance (fp32; the reference runs with nvcc's FMA contraction on a GPU, the oracle with -ffp-contract=off on a CPU,
so agreement is to round-off, several cases are bit-exact): populations max|df| <= 3e-7 after <= 3 steps,
max|drho| <= 2e-5 and max|du| <= 2e-5 after <= 100 steps (measured: <= 8.8e-6 and <= 1.1e-5).
"""
import os

import numpy as np
import pytest

import cases
from cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_matches_reference_cuda(case):
    g = np.load(os.path.join(GOLD, case.name + ".npz"))
    o = cases.make_oracle(case)
    rho0, u0 = case.init_fields()
    o.init(rho0, u0)
    # state right after LBM::init<Scenario>() (init.cuh:45-86)
    assert np.abs(o.populations() - g["f_t0"]).max() <= 2e-8
    assert np.abs(rho0 - g["rho_t0"]).max() <= 1e-7 and np.abs(u0 - g["u_t0"]).max() <= 2e-8
    done = 0
    for k in sorted(set(case.steps_f[1:]) | set(case.steps_m)):
        o.step(k - done)
        done = k
        if k in case.steps_f:
            df = np.abs(o.populations() - g[f"f_t{k}"]).max()
            assert df <= 3e-7, f"{case.name} t={k}: max|df|={df:.2e}"
        if k in case.steps_m:
            rho, u = o.macroscopics()
            dr, du = np.abs(rho - g[f"rho_t{k}"]).max(), np.abs(u - g[f"u_t{k}"]).max()
            assert dr <= 2e-5 and du <= 2e-5, f"{case.name} t={k}: max|drho|={dr:.2e} max|du|={du:.2e}"


def test_golden_fixture_inventory():
    """Every parity case has a committed fixture with the expected arrays."""
    for c in CASES:
        g = np.load(os.path.join(GOLD, c.name + ".npz"))
        for k in c.steps_f:
            assert g[f"f_t{k}"].shape == (c.ny, c.nx, 9)
        for k in (0,) + tuple(c.steps_m):
            assert g[f"rho_t{k}"].shape == (c.ny, c.nx) and g[f"u_t{k}"].shape == (c.ny, c.nx, 2)
        assert np.isfinite(g[f"rho_t{c.steps_m[-1]}"]).all()
