"""ctypes binding of the engine's C ABI (include/lbm_b200.h -> cuda_lbm_b200/liblbm_b200.so).

The shared library is the product; this module only declares its entry points.  There is no
CPU fallback: if the library is missing, import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblbm_b200.so")

LBM_OK, LBM_ERR_INVALID, LBM_ERR_CUDA, LBM_ERR_STATE = 0, -1, -2, -3
BGK, MRT, CM, CM_OPTIMAL = 0, 1, 2, 3
QK_D1_STALE_F0, QK_D2_MRT_ROWS, QK_D3_ZOUHE_RHO, QK_D7_IBM_CLIP, QK_D8_IBM_2X2, QK_D11_BB_RAW = 1, 2, 4, 8, 16, 32
QK_D9_IBM_ZERO_TARGET = 64
QK_REFERENCE, QK_FIXED = 127, 0
ADAPTER_EXACT, ADAPTER_LAGGED = 0, 1
PEER_DESC_BYTES = 128
FLUID, BOUNCE_BACK, ZOU_HE_TOP, ZOU_HE_LEFT = 0, 1, 2, 3
CYLINDER, ZG_OUTFLOW, PRESSURE_OUTLET, REGULARIZED_INLET_TOP = 6, 7, 8, 9
REGULARIZED_BOUNCE_BACK, REGULARIZED_BOUNCE_BACK_CORNER = 11, 12

# every symbol include/lbm_b200.h declares (tests/test_capi_symbols.py checks the header against this list)
SYMBOLS = [
    "lbm_default_config", "lbm_create", "lbm_destroy", "lbm_set_stream", "lbm_set_flags", "lbm_set_force_field", "lbm_set_force_field_device",
    "lbm_set_body_force", "lbm_add_body", "lbm_init_fields", "lbm_init_fields_local", "lbm_init_fields_device", "lbm_init_taylor_green",
    "lbm_set_populations", "lbm_get_populations", "lbm_step", "lbm_step_with_macroscopics", "lbm_sync",
    "lbm_get_macroscopics", "lbm_get_macroscopics_device", "lbm_reserve_macroscopics", "lbm_total_mass", "lbm_moment_avg", "lbm_adapter_prepass",
    "lbm_set_moment_sums", "lbm_get_moment_sums", "lbm_info", "lbm_next_step_needs_halo", "lbm_halo_pack_pre",
    "lbm_halo_unpack_pre", "lbm_halo_pack_post", "lbm_halo_unpack_post", "lbm_peer_export", "lbm_peer_attach", "lbm_peer_attach_all", "lbm_peer_detach", "lbm_set_lookahead",
    "lbm_host_alloc", "lbm_host_free",
    "lbm_velocity_error_sums", "lbm_taylor_green_error_sums", "lbm_row_mean_velocity", "lbm_sample_velocity",
    "lbm_checkpoint_bytes", "lbm_checkpoint_write", "lbm_checkpoint_read",
    "lbm_ibm_exchange_floats", "lbm_ibm_pack", "lbm_ibm_unpack", "lbm_set_body_velocities", "lbm_move_body",
    "lbm_run_from_host", "lbm_recover_macroscopics", "lbm_adapter_sums_pending",
    "lbm_last_error",
]


class LbmConfig(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("periodic_x", C.c_int32), ("periodic_y", C.c_int32),
                ("collision", C.c_int32), ("viscosity", C.c_float), ("S", C.c_float * 9), ("u_max", C.c_float),
                ("force_x", C.c_float), ("force_y", C.c_float), ("quirks", C.c_int32), ("adapter_mode", C.c_int32),
                ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("ibm_mailbox_nodes", C.c_int32), ("reserved", C.c_int32 * 3)]


class LbmInfo(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("y0", C.c_int32), ("ny_local", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("timestep", C.c_int32), ("num_markers", C.c_int32), ("num_ibm_nodes", C.c_int32),
                ("num_neighbour_bc_nodes", C.c_int32), ("device_bytes", C.c_int64), ("bytes_per_cell", C.c_double),
                ("kernel_launches", C.c_int64)]


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lbm_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load liblbm_b200.so.  Raises if it has not been built (python __graft_entry__.py / make -C cuda_lbm_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    dp = C.POINTER(C.c_double)
    sig = {
        "lbm_default_config": [C.POINTER(LbmConfig)],
        "lbm_create": [C.POINTER(LbmConfig), C.POINTER(vp)],
        "lbm_destroy": [vp],
        "lbm_set_stream": [vp, vp],
        "lbm_set_flags": [vp, ip],
        "lbm_set_force_field": [vp, fp],
        "lbm_set_force_field_device": [vp, vp],
        "lbm_set_body_force": [vp, C.c_float, C.c_float],
        "lbm_add_body": [vp, fp, C.c_int32],
        "lbm_init_fields": [vp, fp, fp],
        "lbm_init_fields_local": [vp, vp, vp],
        "lbm_init_fields_device": [vp, vp, vp],
        "lbm_init_taylor_green": [vp, C.c_float, C.c_float],
        "lbm_set_populations": [vp, fp, fp],
        "lbm_get_populations": [vp, fp],
        "lbm_step": [vp, C.c_int32],
        "lbm_step_with_macroscopics": [vp, C.c_int32],
        "lbm_sync": [vp],
        "lbm_get_macroscopics": [vp, vp, vp],
        "lbm_get_macroscopics_device": [vp, C.POINTER(vp), C.POINTER(vp)],
        "lbm_reserve_macroscopics": [vp],
        "lbm_total_mass": [vp, dp],
        "lbm_moment_avg": [vp, fp],
        "lbm_adapter_prepass": [vp],
        "lbm_set_moment_sums": [vp, dp],
        "lbm_get_moment_sums": [vp, dp],
        "lbm_info": [vp, C.POINTER(LbmInfo)],
        "lbm_next_step_needs_halo": [vp],
        "lbm_halo_pack_pre": [vp, C.c_int, vp],
        "lbm_halo_unpack_pre": [vp, C.c_int, vp],
        "lbm_halo_pack_post": [vp, C.c_int, vp],
        "lbm_halo_unpack_post": [vp, C.c_int, vp],
        "lbm_peer_export": [vp, vp],
        "lbm_peer_attach": [vp, C.c_int, vp],
        "lbm_peer_attach_all": [vp, vp, C.c_int32],
        "lbm_peer_detach": [vp],
        "lbm_set_lookahead": [vp, C.c_int32],
        "lbm_host_alloc": [C.POINTER(vp), C.c_int64],
        "lbm_host_free": [vp],
        "lbm_velocity_error_sums": [vp, vp, dp],
        "lbm_taylor_green_error_sums": [vp, C.c_float, C.c_float, C.c_float, dp],
        "lbm_row_mean_velocity": [vp, dp, dp],
        "lbm_sample_velocity": [vp, C.POINTER(C.c_int64), C.c_int32, fp],
        "lbm_checkpoint_bytes": [vp, C.POINTER(C.c_int64)],
        "lbm_checkpoint_write": [vp, C.c_char_p],
        "lbm_checkpoint_read": [vp, C.c_char_p],
        "lbm_ibm_exchange_floats": [vp, C.POINTER(C.c_int64)],
        "lbm_ibm_pack": [vp, vp],
        "lbm_ibm_unpack": [vp, vp],
        "lbm_set_body_velocities": [vp, C.c_int32, fp],
        "lbm_move_body": [vp, C.c_int32, fp],
        "lbm_run_from_host": [vp, vp, vp, C.c_int32, vp, vp],
        "lbm_recover_macroscopics": [vp],
        "lbm_adapter_sums_pending": [vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.lbm_last_error.restype = C.c_char_p
    L.lbm_last_error.argtypes = []
    _lib = L
    return L


def check(rc):
    if rc != LBM_OK:
        raise LbmError(rc, lib().lbm_last_error().decode())
