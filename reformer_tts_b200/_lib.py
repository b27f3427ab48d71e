"""ctypes binding of libreformer_b200.so (the C ABI declared in include/rtts_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
Build it with ``python reformer_tts_b200/csrc/build.py`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p
from pathlib import Path

# RTTS_LIB: experiment / trace builds of the same sources (tools/ only; the product and the tests load the in-tree library)
LIB_PATH = Path(os.environ.get("RTTS_LIB") or Path(__file__).resolve().parent / "libreformer_b200.so")


class LSHSpecStruct(ctypes.Structure):
    """Mirror of ``rtts_lsh_spec``."""
    _fields_ = [("score_scale", c_float), ("key_norm", c_int32), ("mask_value", c_float), ("self_value", c_float),
                ("mask_mode", c_int32), ("causal", c_int32)]


_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float
_SPEC = POINTER(LSHSpecStruct)

# name -> argtypes, in header order
SIGNATURES = {
    "rtts_lsh_hash": [_P, _L, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "rtts_lsh_sumsq": [_P, _L, _P, _I, _I, _I, _I, _P],
    "rtts_lsh_sort": [_P, _P, _P, _I, _I, _I, _I, _P],
    "rtts_lsh_attn_fwd": [_P, _P, _L, _P, _P, _P, _SPEC, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "rtts_lsh_hash_tc": [_P, _L, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "rtts_lsh_hash_tc_supported": [_I, _I, _I, _I],
    "rtts_lsh_hash_tc_workspace_bytes": [_I, _I, _I],
    "rtts_lsh_merge_fwd": [_P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P],
    "rtts_lsh_delta": [_P, _P, _L, _P, _I, _I, _I, _I, _P],
    "rtts_lsh_attn_bwd": [_P, _P, _L, _P, _P, _P, _SPEC, _P, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "rtts_lsh_grad_reduce": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P],
    "rtts_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _F, _P],
    "rtts_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "rtts_layernorm_bwd_acc": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "rtts_gemm_bf16": [_P, _L, _I, _P, _L, _I, _P, _L, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P],
    "rtts_xattn_fwd": [_P, _L, _P, _P, _L, _P, _F, _F, _P, _P, _L, _P, _I, _I, _I, _I, _I, _P],
    "rtts_xattn_bwd": [_P, _L, _P, _P, _L, _P, _F, _F, _P, _P, _L, _P, _P, _P, _L, _P, _P, _L, _I, _I, _I, _I, _I, _P],
    "rtts_colsum_bf16": [_P, _L, _P, _I, _I, _P],
    "rtts_cast_bf16_colsum": [_P, _P, _P, _I, _I, _P],
    "rtts_cast_bf16_colsum_dropout": [_P, _P, _F, _P, _P, _I, _I, _P],
    "rtts_gemm_bf16_dropout": [_P, _L, _I, _P, _L, _I, _P, _L, _P, _P, _L, _P, _P, _L, _F, _I, _I, _I, _I, _I, _P],
}

RESTYPES = {"rtts_lsh_hash_tc_workspace_bytes": c_int64}      # everything else returns an int status / flag

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension is not built and there is no fallback path. "
                               "Run `python reformer_tts_b200/csrc/build.py`.")
        lib = ctypes.CDLL(str(LIB_PATH))
        lib.rtts_last_error.restype = c_char_p
        lib.rtts_last_error.argtypes = []
        lib.rtts_abi_version.restype = c_int
        lib.rtts_abi_version.argtypes = []
        lib.rtts_build_id.restype = c_char_p
        lib.rtts_build_id.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
            fn.restype = RESTYPES.get(name, c_int)
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.rtts_last_error().decode()}")


def build_id() -> str:
    """Hash of the sources the loaded binary was built from (csrc/build.py ``source_hash``)."""
    return load().rtts_build_id().decode()
