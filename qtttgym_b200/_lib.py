"""ctypes binding of libqttt_b200.so (include/qttt_b200.h).  There is no CPU fallback: if the
CUDA library is missing or cannot be loaded every entry point fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

ACT_INDEX, ACT_PAIR = 0, 1
ST_OK, ST_ILLEGAL, ST_FINISHED, ST_RESET = 0, 1, 2, 4
STEP_FRESH, STEP_AUTORESET, STEP_AUTORESET_NEXT = 1, 2, 4
ABI_VERSION = 2

_lib = None

_vp, _i64, _u64, _i32, _u32, _int = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32, C.c_int

_SIGNATURES = {
    "qttt_abi_version": ([], _int),
    "qttt_strerror": ([_int], C.c_char_p),
    "qttt_reset": ([_vp, _vp, _i64, _vp], _int),
    "qttt_step": ([_vp, _vp, _int, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_reset_step": ([_vp, _vp, _int, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_reset_all": ([_vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_ex": ([_vp, _vp, _int, _vp, _u64, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_packed": ([_vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_packed_obs": ([_vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_packed_mapped": ([_vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_packed12_mapped": ([_vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_packed_host_obs12": ([_vp, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(C.c_void_p), _int], _int),
    "qttt_step_packed12_host": ([_vp, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(C.c_void_p), _int], _int),
    "qttt_step_packed_host": ([_vp, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(C.c_void_p), _int], _int),
    "qttt_step_packed_host_obs": ([_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, C.POINTER(C.c_void_p), _int], _int),
    "qttt_step_random": ([_vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_random_ex": ([_vp, _u64, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_observe": ([_vp] * 11 + [_i64, _vp], _int),
    "qttt_features": ([_vp, _vp, _i64, _vp], _int),
    "qttt_env1": ([_vp, _int, _int, _int, _int, _u64, _u64, _vp, _u32, _vp], _int),
    "qttt_qeval1": ([_vp, _int, _vp, _u32, _vp], _int),
    "qttt_get_mask": ([_vp, _vp, _i64, _vp], _int),
    "qttt_step_features": ([_vp, _vp, _int, _vp, _u64, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_step_obs": ([_vp, _vp, _int, _vp, _u64, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_pack": ([_vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_qeval_both": ([_vp] * 10 + [_i64, _vp], _int),
    "qttt_rollout": ([_vp, _i64, _i32, _u64, _vp, _vp, _vp, _vp], _int),
    "qttt_mcts_node_bytes": ([], _int),
    "qttt_mcts_init": ([_vp, _i64, _vp, _vp, _i64, _vp], _int),
    "qttt_mcts_run": ([_vp, _i64, _vp, _i32, _i32, C.c_double, _u64, _u64, _i64, _vp], _int),
    "qttt_mcts_stats": ([_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_mcts_sync": ([_vp, _i64, _vp, _vp, _vp, _i64, _vp], _int),
    "qttt_sweep": ([_i64, _i64, _u64, _vp, _vp], _int),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class QtttLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise QtttLibraryError(
            f"{LIB} is missing: build it with `python -m qtttgym_b200.build` "
            "(nvcc, sm_100a).  qtttgym_b200 has no CPU fallback.")
    try:
        handle = C.CDLL(LIB)
    except OSError as e:  # e.g. libcudart not found
        raise QtttLibraryError(f"cannot load {LIB}: {e}") from e
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the .so does not export the ABI
        fn.argtypes = argtypes
        fn.restype = restype
    if handle.qttt_abi_version() != ABI_VERSION:
        raise QtttLibraryError("libqttt_b200.so ABI version mismatch; rebuild it")
    _lib = handle
    return _lib


#: number of kernel launches issued through this binding (one per C call unless the caller
#: says otherwise); bench.py reads it to report how many of our kernels ran in a timed region
LAUNCHES = 0


def check(rc: int, launches: int = 1) -> None:
    global LAUNCHES
    if rc != 0:
        raise RuntimeError(lib().qttt_strerror(rc).decode())
    LAUNCHES += launches


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
