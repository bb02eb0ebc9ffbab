"""cuda_lbm_b200 — B200-native D2Q9 lattice-Boltzmann hot path behind the reference's solver interface.

csrc/        CUDA kernels (sm_100a) + the C ABI declared in include/lbm_b200.h -> liblbm_b200.so
_capi.py     ctypes declarations of that ABI (no CPU fallback: import fails if the library is missing)
solver.py    Engine: one handle of the C ABI with numpy in / out (tests, bench.py, the multi-process launcher)
slab.py      y-slab decomposition over the GPUs of one box (torch.distributed / NCCL halo exchange)
"""
from ._capi import (ADAPTER_EXACT, ADAPTER_LAGGED, BGK, CM, CM_OPTIMAL, MRT, QK_FIXED, QK_REFERENCE, LbmError,  # noqa: F401
                    lib)
from .solver import Engine, default_S  # noqa: F401
