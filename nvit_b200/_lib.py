"""ctypes binding of libnvit_b200.so — the C ABI declared in include/nvit_b200.h.

This is the binding a maintainer of the reference would add.  There is NO fallback: if the shared
library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NVIT_LIB_PATH") or os.path.join(_HERE, "libnvit_b200.so")

P, I64, I32, F32 = c_void_p, c_int64, c_int, c_float

# name -> argument types (all return int except the three library calls)
SIGNATURES = {
    "nvit_gemm_bf16": [P, P, P, P, I64, I64, I64, I64, I64, I64, I64, I32, I32, I32, I32, I32, P, P, F32, P, I64, I64, P],
    "nvit_gemm_force_cta_group": [I32],
    "nvit_gemm_raster_group": [I32],
    "nvit_gemm_swiglu_cta_group": [I32],
    "nvit_attention_bwd_variant": [I32],
    "nvit_attention_fwd_variant": [I32],
    "nvit_set_sm_budget": [I32],
    "nvit_set_pdl": [I32],
    "nvit_tmap_cache_stats": [P, P],
    "nvit_residual_bwd_staged": [I32],
    "nvit_cast_f32_to_bf16": [P, P, I64, P],
    "nvit_sumsq_f32": [P, I64, P, P],
    "nvit_sumsq_f32_det": [P, I64, P, P, I64, P],
    "nvit_colsum_bf16": [P, I64, I64, I64, P, P],
    "nvit_pos_bias_grad": [P, I64, I64, I64, P, P, P],
    "nvit_residual_fwd": [P, P, P, F32, P, P, P, P, I64, I64, P],
    "nvit_residual_bwd": [P, P, P, P, F32, P, P, P, I32, P, P, P, P, I64, I64, P],
    "nvit_add_rmsnorm_fwd": [P, P, P, F32, P, P, I64, I64, P],
    "nvit_add_rmsnorm_bwd": [P, P, P, P, F32, P, I32, P, P, I64, I64, P],
    "nvit_add_skipnorm_fwd": [P, P, P, P, P, P, I64, I64, P],
    "nvit_add_skipnorm_bwd": [P, P, P, P, P, P, P, P, P, I64, I64, P],
    "nvit_swiglu_fwd": [P, P, F32, P, I64, I64, P],
    "nvit_swiglu_bwd": [P, P, P, F32, P, P, I64, I64, P],
    "nvit_attention_fwd": [P, P, P, I64, I64, I64, P, F32, F32, P, I64, P, I64, I64, I64, I64, P, P, I64, I64, P],
    "nvit_attention_bwd": [P, P, P, I64, I64, I64, P, F32, F32, P, P, I64, P, P, P, P, I64, I64, I64, P, I64, I64, I64, I64, P, P, I64, I64, P],
    "nvit_gemm_qknorm": [P, P, P, I64, I64, I64, I64, I64, I64, P, P, F32, I64, I64, P, I64, P],
    "nvit_gemm_gate_bwd": [P, P, P, P, F32, P, I64, I64, I64, I64, I64, I64, I64, P],
    "nvit_rowdot_div": [P, P, P, P, I64, I64, P],
    "nvit_split_bf16": [P, P, P, I64, P],
    "nvit_som_prepare": [P, I64, I64, P, P, P, P, P],
    "nvit_som_select": [P, P, P, I64, I64, I64, P, P, P, P, P, P, P],
    "nvit_som_pool": [P, I64, I64, P, P],
    "nvit_som_update": [P, P, P, I64, I64, I64, I64, P, F32, P],
    "nvit_som_pair_losses": [P, P, P, P, I64, I64, P, P, P, P, P, P, P],
    "nvit_som_smoothness": [P, P, I64, I64, I64, P, P, P, P],
    "nvit_tanh_mse_bwd": [P, P, I64, F32, P, P, P],
    "nvit_im2col_u8": [P, P, I64, I64, I64, I64, I64, I64, F32, F32, P],
    "nvit_augment_u8": [P, P, P, P, I64, I64, I64, P],
    "nvit_im2col_bf16": [P, P, I64, I64, I64, I64, I64, I64, P],
    "nvit_pool_ln_fwd": [P, P, P, F32, P, P, P, I64, I64, I64, P],
    "nvit_pool_ln_bwd": [P, P, P, P, P, P, P, I64, I64, I64, P],
    "nvit_head_scale_bwd": [P, P, P, F32, P, P, I64, I64, I64, P],
    "nvit_cross_entropy": [P, P, P, P, F32, I64, I64, P],
    "nvit_tanh_mse": [P, P, I64, F32, P, P],
    "nvit_adamw_flat": [P, P, P, P, I64, I64, F32, F32, F32, F32, F32, I64, P, F32, P, P],
    "nvit_weight_norm_multi": [P, I64, I64, P],
    "nvit_adamw_norm_fused": [P, P, P, P, P, P, I64, I64, F32, F32, F32, F32, F32, I64, P, F32, P, P, I32, P],
    "nvit_head_scale_fwd": [P, P, F32, P, I64, I64, P],
}
# measurement-only entry points: present only in -DNVIT_BENCH_HOOKS builds (include/nvit_b200_tuning.h, section 2)
HOOK_SIGNATURES = {"nvit_gemm_debug": [I32], "nvit_attention_debug": [P]}
LIBRARY_CALLS = {"nvit_last_error": ([], c_char_p), "nvit_version": ([], c_int), "nvit_sm_count": ([], c_int)}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built: there is no CPU or PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m nvit_b200.build` (nvit_b200 has no fallback path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, argtypes in HOOK_SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
    for name, (argtypes, restype) in LIBRARY_CALLS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    mode = os.environ.get("NVIT_SWIGLU_CTA_GROUP")    # benchmarking hook: CTA-group mode of the gate GEMM
    if mode in ("1", "2"):
        lib.nvit_gemm_swiglu_cta_group(int(mode))
    mode = os.environ.get("NVIT_GATEB_GROUPS")        # epilogue groups (2 or 4) of the fused gate-backward GEMM
    if mode in ("2", "4"):
        lib.nvit_gemm_swiglu_cta_group(20 + int(mode))
    mode = os.environ.get("NVIT_GATEB_CTA_GROUP")     # ... and of the fused gate-backward GEMM
    if mode in ("1", "2"):
        lib.nvit_gemm_swiglu_cta_group(10 + int(mode))
    mode = os.environ.get("NVIT_RESIDUAL_STAGED")     # form of the residual backward kernel (see include/nvit_b200.h)
    if mode in ("0", "1", "2"):
        lib.nvit_residual_bwd_staged(int(mode))
    mode = os.environ.get("NVIT_GEMM_RASTER_GROUP")   # tile order of the GEMMs: band width along n (0 = n fastest)
    if mode is not None and mode.lstrip("-").isdigit():
        lib.nvit_gemm_raster_group(int(mode))
    mode = os.environ.get("NVIT_PDL")                 # programmatic dependent launch of every kernel (see include/nvit_b200.h)
    if mode in ("0", "1"):
        lib.nvit_set_pdl(int(mode))
    mode = os.environ.get("NVIT_ATTN_BWD_VARIANT")    # 1 = round-1 single-role kernel, 2 = warp-specialised, 3 = 2 + epilogue warpgroup
    if mode in ("1", "2", "3"):
        lib.nvit_attention_bwd_variant(int(mode))
    mode = os.environ.get("NVIT_ATTN_FWD_VARIANT")    # 1 = one head per CTA (round 1), 2 = persistent warp-specialised kernel
    if mode in ("1", "2"):
        lib.nvit_attention_fwd_variant(int(mode))
    mode = os.environ.get("NVIT_GEMM_CTA_GROUP")      # benchmarking hook: pin cta_group::1 or ::2 GEMM tiles
    if mode in ("1", "2"):
        lib.nvit_gemm_force_cta_group(int(mode))
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().nvit_last_error()
    return msg.decode() if msg else ""


# Measurement aid (bench.py's per-kernel-group roofline table): when PROBE is a list, every entry-point call is bracketed by
# CUDA events on the current stream and (name, args, start, end) is appended.  None (the default) costs one comparison.
PROBE = None


# Data-parallel overlap (train.GradReducer): after a bucket's all-reduce has been put on the side stream, the next `n` entry-point
# calls size their persistent grids for `budget` SMs, so that NCCL's CTAs find free SMs beside them; then all SMs again.
_WINDOW = [0]


def sm_budget_window(n_calls: int, budget: int) -> None:
    load().nvit_set_sm_budget(int(budget))
    _WINDOW[0] = int(n_calls)


def call(name: str, *args) -> None:
    """Call an entry point; a non-zero status becomes a RuntimeError carrying nvit_last_error()."""
    if _WINDOW[0] > 0:
        _WINDOW[0] -= 1
        if _WINDOW[0] == 0:
            load().nvit_set_sm_budget(0)
    if PROBE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(load(), name)(*args)
        e1.record()
        PROBE.append((name, args, e0, e1))
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}: {last_error()}")


def tmap_cache_stats() -> tuple[int, int]:
    """(maps encoded by the driver, maps served from the table) since the library was loaded."""
    enc, hit = c_int64(0), c_int64(0)
    load().nvit_tmap_cache_stats(ctypes.byref(enc), ctypes.byref(hit))
    return int(enc.value), int(hit.value)
