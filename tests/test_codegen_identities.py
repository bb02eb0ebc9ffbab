"""CPU: the folded moment transforms of the CUDA path are the reference's linear maps.

tools/codegen/d2q9_transforms.py proves with exact rational arithmetic (sympy) that the add/FMA chains of
cuda_lbm_b200/csrc/collide.cuh — transcribed there statement by statement — equal M f, M^-1 m, M f_eq, T(u) f and T(u)^-1 k of
the reference's definitions, and that BGK == MRT with S = omega.  Here the script runs, and its M / M^-1 are compared with the
literal tables the oracle restates from src/core/lbm_constants.cuh:33-55.
"""
import io
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "codegen"))
import d2q9_transforms as T  # noqa: E402


def _oracle_table(name):
    src = open(os.path.join(ROOT, "oracle", "lbm_oracle.c")).read()
    body = re.search(r"static const float %s\[Q \* Q\] = \{(.*?)\};" % name, src, re.S).group(1)
    vals = [eval(tok.replace("f", ""), {}) for tok in body.replace("\n", " ").split(",") if tok.strip()]
    return np.array(vals, np.float64).reshape(9, 9)


def test_all_identities_hold():
    assert T.verify() >= 100


def test_matrices_are_the_reference_tables():
    M = np.array(T.mrt_matrix().tolist(), np.float64)
    assert np.array_equal(M, _oracle_table("Mm"))                       # h_M, lbm_constants.cuh:33-43
    Minv = np.array(T.mrt_matrix().inv().tolist(), np.float64)
    assert np.abs(Minv - _oracle_table("Mi")).max() < 1e-15             # h_M_inv, lbm_constants.cuh:45-55


def test_emit_produces_c_text_for_every_transform():
    buf = io.StringIO()
    T.emit(buf)
    txt = buf.getvalue()
    for title in ("m = M f", "f = M^-1 m", "k = T(u) f", "f = T^-1(u) k", "f_eq"):
        assert title in txt
    assert txt.count(";") > 100 and "pow(" not in txt
