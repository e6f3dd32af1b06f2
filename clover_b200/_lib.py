"""ctypes loader for the product library ``libclover_b200.so`` (the C ABI of include/clover_b200.h).

There is no fallback of any kind: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libclover_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_SIZE, ERR_UNSUPPORTED = range(5)
DOT_AUTO, DOT_EXACT, DOT_FAST = 0, 1, 2
THRESHOLD_AUTO, THRESHOLD_EXACT, THRESHOLD_FAST = 0, 1, 2

_u64, _vp, _int = C.c_uint64, C.c_void_p, C.c_int

# name -> (restype, argtypes); mirrors include/clover_b200.h one to one
_SIGNATURES = {
    "clover_version": (_int, []),
    "clover_last_error": (C.c_char_p, []),
    "clover_device_count": (_int, []),
    "clover_set_device": (_int, [_int]),
    "clover_size_pad": (_u64, [_u64]),
    "clover_dot_exact_limit": (_u64, []),
    "clover_threshold_exact_limit": (_u64, []),
    "clover_kernel_launches": (_int, []),
    "clover_malloc": (_int, [C.POINTER(_vp), C.c_size_t]),
    "clover_free": (_int, [_vp]),
    "clover_malloc_host": (_int, [C.POINTER(_vp), C.c_size_t]),
    "clover_free_host": (_int, [_vp]),
    "clover_memset": (_int, [_vp, _int, C.c_size_t, _vp]),
    "clover_copy_h2d": (_int, [_vp, _vp, C.c_size_t, _vp]),
    "clover_copy_d2h": (_int, [_vp, _vp, C.c_size_t, _vp]),
    "clover_copy_d2d": (_int, [_vp, _vp, C.c_size_t, _vp]),
    "clover_stream_sync": (_int, [_vp]),
    "clover_ipc_export": (_int, [_vp, _vp]),
    "clover_ipc_import": (_int, [_vp, C.POINTER(_vp)]),
    "clover_ipc_close": (_int, [_vp]),
    "clover_prng_init": (_int, [_u64, _u64, _vp]),
    "clover_prng_next": (_int, [_vp, _vp]),
    "clover_prng_skip": (_int, [_vp, _u64]),
    "clover_v32_set_random_floats": (_int, [_vp, _u64, C.c_float, C.c_float, _vp, _vp]),
    "clover_v32_set_random_integers": (_int, [_vp, _u64, C.c_float, C.c_float, _vp, _vp]),
    "clover_v4_quantize": (_int, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "clover_v4_restore": (_int, [_vp, _vp, _u64, _vp, _vp]),
    "clover_v4_dot": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _int, _vp]),
    "clover_v4_scale_and_add": (_int, [_vp, _vp, _vp, _vp, C.c_float, _u64, _vp, _vp, _vp, _vp]),
    "clover_v8_scale_and_add": (_int, [_vp, _vp, _vp, _vp, C.c_float, _u64, _vp, _vp, _vp, _vp]),
    "clover_v4_threshold": (_int, [_vp, _vp, _u64, _u64, _int, _vp]),
    "clover_v8_threshold": (_int, [_vp, _vp, _u64, _u64, _int, _vp]),
    "clover_v8_quantize": (_int, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "clover_v8_restore": (_int, [_vp, _vp, _u64, _vp, _vp]),
    "clover_v8_dot": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _int, _vp]),
    "clover_m4_quantize": (_int, [_vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "clover_m4_mvm": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_mvm_v8": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_mvm_f32": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "clover_m4_restore": (_int, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "clover_m8_restore": (_int, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "clover_m8_mvm_f32": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "clover_m4_mvm_shard": (_int, [_vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_mvm_shard_fused": (_int, [_vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, C.c_uint32, _vp, _vp]),
    "clover_m4_mvm_shard_stamped": (_int, [_vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, C.c_uint32, _vp, _vp]),
    "clover_m4_shard_stamped_unpack": (_int, [_vp, _u64, _u64, _u64, C.c_uint32, _vp, _vp, _vp]),
    "clover_v4_requantize_mvm": (_int, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "clover_m4_gemm": (_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp]),
    "clover_m4_expand_e4m3": (_int, [_vp, _u64, _u64, _vp, _vp]),
    "clover_m4_gemm_expanded": (_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp]),
    "clover_m4_gemm_simt": (_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp, _u64, _vp]),
    "clover_m8_quantize": (_int, [_vp, _u64, _u64, _vp, _vp, _vp, _vp]),
    "clover_m8_mvm": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_transpose": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "clover_m8_transpose": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "clover_host_v4_quantize": (_int, [_vp, _u64, _vp, _vp, _vp]),
    "clover_host_v4_dot": (_int, [_vp, _vp, _vp, _vp, _u64, _vp, _int]),
    "clover_host_m4_mvm": (_int, [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_sharded_create": (_int, [C.POINTER(_vp), _u64, _u64, _int, _vp]),
    "clover_m4_sharded_destroy": (_int, [_vp]),
    "clover_m4_sharded_world": (_int, [_vp]),
    "clover_m4_sharded_shard": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "clover_m4_sharded_load_host": (_int, [_vp, _vp, _vp]),
    "clover_m4_sharded_mvm_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
}

ABI_SYMBOLS = tuple(_SIGNATURES)


class CloverError(RuntimeError):
    """A C-ABI call returned a non-zero clover_status."""

    def __init__(self, fn: str, code: int, message: str):
        super().__init__(f"{fn} failed with status {code}: {message}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a into clover_b200/libclover_b200.so."""
    res = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libclover_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded shared library (raises if it has not been built - there is no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `make -C clover_b200/csrc` (or __graft_entry__.build()). "
                "clover_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(fn: str, status: int) -> None:
    if status != OK:
        raise CloverError(fn, status, lib().clover_last_error().decode("utf-8", "replace"))


def call(fn: str, *args) -> None:
    check(fn, getattr(lib(), fn)(*args))
