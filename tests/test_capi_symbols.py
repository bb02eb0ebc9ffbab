"""CPU: the C-ABI library loads without a GPU and exports every symbol include/lbm_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    assert "lbm_create" in syms and "lbm_step" in syms and "lbm_last_error" in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    from cuda_lbm_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(_capi.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/lbm_b200.h but not exported: {missing}"
    assert sorted(_capi.SYMBOLS) == header_symbols(), "cuda_lbm_b200/_capi.py SYMBOLS out of sync with the header"


def test_ctypes_struct_sizes_match_header_layout():
    from cuda_lbm_b200._capi import LbmConfig, LbmInfo
    # lbm_config: 5 int32 + float + 9 float + 3 float + 2 int32 + 3 int32 + 4 int32 = 27 x 4 bytes
    assert C.sizeof(LbmConfig) == 27 * 4
    # lbm_info_t: 10 int32 + int64 + double + int64
    assert C.sizeof(LbmInfo) == 10 * 4 + 3 * 8


def test_no_cpu_fallback_argument_errors_without_gpu():
    """Calls that do not need a device still follow the error convention (no exit(), message retrievable)."""
    from cuda_lbm_b200 import _capi
    L = _capi.lib()
    assert L.lbm_default_config(None) == _capi.LBM_ERR_INVALID
    assert b"NULL" in L.lbm_last_error()
    cfg = _capi.LbmConfig()
    assert L.lbm_default_config(C.byref(cfg)) == 0
    assert cfg.quirks == _capi.QK_REFERENCE and abs(cfg.viscosity - 1.0 / 6.0) < 1e-7
    cfg.nx = 2
    h = C.c_void_p()
    assert L.lbm_create(C.byref(cfg), C.byref(h)) == _capi.LBM_ERR_INVALID


def test_new_entry_points_follow_the_error_convention_without_gpu():
    """Argument checks come before any device work: NULL handles / pointers return LBM_ERR_INVALID with a message, never crash."""
    from cuda_lbm_b200 import _capi
    L = _capi.lib()
    two = (C.c_double * 2)()
    n = C.c_int64()
    for rc in (L.lbm_run_from_host(None, None, None, 1, None, None), L.lbm_checkpoint_write(None, b"/tmp/x"), L.lbm_checkpoint_read(None, None),
               L.lbm_checkpoint_bytes(None, C.byref(n)), L.lbm_velocity_error_sums(None, None, two), L.lbm_taylor_green_error_sums(None, 0.1, 0.1, 0.0, two),
               L.lbm_row_mean_velocity(None, None, None), L.lbm_ibm_exchange_floats(None, C.byref(n)), L.lbm_ibm_pack(None, None), L.lbm_ibm_unpack(None, None),
               L.lbm_set_body_velocities(None, 0, None), L.lbm_move_body(None, 0, None)):
        assert rc == _capi.LBM_ERR_INVALID
        assert L.lbm_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "cuda_lbm_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.lower() or fn in (), f"{fn} mentions the oracle"
