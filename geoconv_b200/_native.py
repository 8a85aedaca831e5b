"""ctypes binding of libgeoconv_b200.so - the C ABI declared in include/geoconv_b200.h.

There is no fallback: if the library is missing and cannot be built, importing the compute path
raises.  Tensors cross the boundary as raw device pointers (``tensor.data_ptr()``) plus sizes and
the current CUDA stream handle; PyTorch only owns the memory.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgeoconv_b200.so")

ABI_VERSION = 7
MAX_ROT = 32

PREC_FP32, PREC_TF32, PREC_TF32X3, PREC_F16X3 = 0, 1, 2, 3
CSR_RING_GROUPED = 1
ACT_IDS = {"relu": 0, "elu": 1, "leaky_relu": 2, "selu": 3, "sigmoid": 4, "tanh": 5}

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_int = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/geoconv_b200.h
SIGNATURES = {
    "gcb_abi_version": (_int, []),
    "gcb_last_error": (C.c_char_p, []),
    "gcb_launch_count": (_i64, []),
    "gcb_yield_sms": (None, [_int]),
    "gcb_profile_enable": (None, [_int]),
    "gcb_profile_read": (_int, [_int, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "gcb_tc_supported": (_int, [_i32, _i32, _i32, _i32, _i32]),
    "gcb_tc_half_supported": (_int, [_i32, _i32, _i32, _i32, _i32]),
    "gcb_prep_bary": (_int, [_p, _int, _i64, _i32, _p, _p, _p, _p]),
    "gcb_csr_workspace_bytes": (_sz, [_i64, _i32]),
    "gcb_csr_build": (_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "gcb_fold_prior": (_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _int, _p]),
    "gcb_conv_fwd_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "gcb_conv_fwd": (_int, [_p, _p, _p, _p, _p, _p, C.POINTER(_i32), _i32, _i32, _i32, _i32, _i32, _i32,
                            _i32, _int, _int, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    "gcb_fwd_records_bytes": (_sz, [_i32, _i32, _i32]),
    "gcb_fwd_pack_records": (_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "gcb_fwd_tables_bytes": (_sz, []),
    "gcb_fwd_build_tables": (_int, [C.POINTER(_i32), _i32, _i32, _i32, _i32, _i32, _i32, _int, _p, _p]),
    "gcb_conv_amp_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "gcb_conv_amp_bwd": (_int, [_p, _p, _p, C.POINTER(_i32), _i32, _int, _p, _p, _p, _p, _p, _p, _p, _int, _p, _p,
                                _i32, _i32, _i32, _i32, _i32, _i32, _int, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "gcb_conv_amp_bwd_can_split": (_int, [_i32, _i32, _i32, _i32, _int, _int]),
    "gcb_amp_fwd": (_int, [_p, _i32, _i32, _i32, _p, _p, _p]),
    "gcb_amp_bwd": (_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "gcb_bary_from_gpc": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _p]),
    "gcb_gpc_smem_per_system": (_i64, [_i32]),
    "gcb_gpc_wavefront": (_int, [_p, _p, _p, _p, _p, _i32, _p, _i32, C.c_double, C.c_double, _i32, _p, _p, _p, _p,
                                 _p, _p, _p, _p, _p]),
    "gcb_bn_workspace_bytes": (_sz, [_i32, _i32]),
    "gcb_bn_stats": (_int, [_p, _i32, _i32, C.c_float, _p, _p, _p, _p, C.c_float, _p, _sz, _p]),
    "gcb_bn_dropout_fwd": (_int, [_p, _i32, _i32, _p, _p, _p, _p, C.c_float, C.c_uint64, C.c_uint64, _p, _p]),
    "gcb_bn_dropout_bwd": (_int, [_p, _p, _i32, _i32, _p, _p, _p, _int, C.c_float, C.c_uint64, C.c_uint64, _p, _p, _p,
                                  _p, _sz, _p]),
    "gcb_apool_fwd": (_int, [_p, _i32, _i32, _i32, _int, _p, _p, _p]),
    "gcb_apool_bwd": (_int, [_p, _p, _i32, _i32, _i32, _int, _p, _p]),
    "gcb_act_grad": (_int, [_p, _p, _i64, _int, _p, _p]),
    "gcb_select_rot": (_int, [_p, C.POINTER(_i32), _i32, _i32, _i32, _p, _p]),
    "gcb_conv_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "gcb_conv_bwd_workspace_selfcheck": (_int, [_i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "gcb_conv_bwd": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _int, _i32, _i32, _i32, _i32, _i32, _i32, _int,
                            _int, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "gcb_conv_bwd_dense_supported": (_int, [_i32, _i32, _i32, _i32, _int]),
    "gcb_conv_bwd_dense": (_int, [_p, _p, _p, _p, _p, _p, C.POINTER(_i32), _i32, _p, _p, _int, _i32, _i32, _i32, _i32, _i32,
                                  _i32, _int, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "gcb_pack_rows": (_int, [_p, _p, _i32, _i32, _p, _p]),
    "gcb_peer_allreduce_bytes": (_sz, [_i64, _i32]),
    "gcb_peer_allreduce": (_int, [_p, _i64, _p, _i64, _p, _i64, _p, _i32, _i32, _i32, _p]),
    "gcb_peer_allreduce_staged": (_int, [_i64, _p, _i32, _i32, _i32, _p]),
}

_lib = None
_lock = threading.Lock()


class GeoconvNativeError(RuntimeError):
    """A C-ABI call returned a non-zero GCB_E* code."""


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if the .so is absent and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
                from . import build as _build
                _build.build()
            else:
                raise ImportError(
                    f"geoconv_b200: {LIB_PATH} is missing and nvcc is not available to build it. "
                    "Run `python -m geoconv_b200.build` (there is no CPU or PyTorch fallback).")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        got = lib.gcb_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"geoconv_b200: ABI version mismatch (library {got}, binding {ABI_VERSION}); rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gcb_last_error()
        raise GeoconvNativeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def rot_array(offsets):
    arr = (_i32 * len(offsets))(*[int(o) for o in offsets])
    return arr
