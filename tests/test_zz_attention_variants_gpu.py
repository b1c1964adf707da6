"""The attention-backward kernels (include/nvit_b200_tuning.h: nvit_attention_bwd_variant) against the fp32 reference and
against each other: 1 = the single-role kernel of round 1 (one head per CTA), 2 = the persistent warp-specialised kernel
(8 compute warps + one MMA warp, the products of the next (kv tile, q tile) item in flight under the passes of the current
one, the next head's tiles loading meanwhile), 3 = 2 with the dV / dK / dQ epilogues on a warpgroup of their own (one thread
per accumulator row; falls back to 2 for raw q / k with sqk and for T > 208).  The file sorts
last on purpose: a fault in a kernel variant must not hide the rest of the suite behind `-x`."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from nvit_b200 import _lib, ops  # noqa: E402

_spec = importlib.util.spec_from_file_location("kernels_gpu_cases", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_kernels_gpu.py"))
K = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(K)

DEFAULT_VARIANT = int(os.environ.get("NVIT_ATTN_BWD_VARIANT", "3"))


@pytest.fixture(params=[1, 2, 3], ids=["single_role", "persistent", "persistent_epilogue_wg"])
def variant(request):
    _lib.call("nvit_attention_bwd_variant", request.param)
    yield request.param
    _lib.call("nvit_attention_bwd_variant", DEFAULT_VARIANT)


@pytest.mark.parametrize("B,H,T", [(2, 1, 64), (3, 3, 64), (2, 2, 196), (1, 12, 196), (2, 2, 16), (1, 2, 256), (2, 1, 130), (1, 1, 48),
                                   (2, 1, 128), (1, 2, 144), (40, 12, 196)])
@pytest.mark.parametrize("normed", [True, False])
def test_attention_backward_variants_match_reference(variant, B, H, T, normed):
    K.test_attention_fwd_bwd(B, H, T, normed)


@pytest.mark.parametrize("B,H,T", [(2, 2, 196), (3, 1, 64), (1, 2, 256)])
def test_attention_backward_variants_with_prenormalised_qk(variant, B, H, T):
    K.test_attention_with_prenormalised_qk(B, H, T)


@pytest.mark.parametrize("B,H,T", [(4, 12, 196), (2, 3, 64), (3, 2, 256)])
def test_attention_backward_variants_agree(B, H, T):
    """Same tiles, same arithmetic; only the fp32 accumulation order over q tiles differs: outputs agree to bf16 rounding."""
    C, M = H * 64, B * T
    qkv = K.randn(M, 3 * C, seed=70, scale=0.5, dtype=torch.bfloat16)
    sqk = (1.0 + 0.2 * K.randn(C, seed=71)).mul(0.03)
    gb = K.randn(M, C, seed=72, scale=0.1, dtype=torch.bfloat16)
    out = torch.zeros(M, C, device=K.DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=K.DEV)
    ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.03, 8.0, out, lse, B, H, T)
    res = {}
    try:
        for v in (1, 2, 3):
            _lib.call("nvit_attention_bwd_variant", v)
            for rep in range(3):       # repeated launches: no state may leak from one head / launch to the next
                d = torch.zeros(M, 3 * C, device=K.DEV, dtype=torch.bfloat16)
                ds = torch.zeros(C, device=K.DEV)
                ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.03, 8.0, out, gb, lse,
                                  d[:, :C], d[:, C:2 * C], d[:, 2 * C:], ds, B, H, T)
                torch.cuda.synchronize()
                if rep == 0:
                    res[v] = (d, ds)
                else:
                    assert torch.equal(d, res[v][0]), (v, rep)
    finally:
        _lib.call("nvit_attention_bwd_variant", DEFAULT_VARIANT)
    for v in (2, 3):
        assert K.rel(res[v][0], res[1][0]) <= 5e-3, (v, K.rel(res[v][0], res[1][0]))
        assert K.rel(res[v][1], res[1][1]) <= 5e-3


DEFAULT_FWD_VARIANT = int(os.environ.get("NVIT_ATTN_FWD_VARIANT", "2"))


@pytest.fixture(params=[1, 2], ids=["fwd_one_head_per_cta", "fwd_persistent"])
def fwd_variant(request):
    _lib.call("nvit_attention_fwd_variant", request.param)
    yield request.param
    _lib.call("nvit_attention_fwd_variant", DEFAULT_FWD_VARIANT)


@pytest.mark.parametrize("B,H,T", [(2, 1, 64), (3, 3, 64), (2, 2, 196), (1, 12, 196), (2, 2, 16), (1, 2, 256), (2, 1, 130), (1, 1, 48),
                                   (2, 1, 128), (1, 2, 144), (40, 12, 196)])
@pytest.mark.parametrize("normed", [True, False])
def test_attention_forward_variants_match_reference(fwd_variant, B, H, T, normed):
    K.test_attention_fwd_bwd(B, H, T, normed)


@pytest.mark.parametrize("B,H,T", [(2, 2, 196), (3, 1, 64), (1, 2, 256), (30, 12, 196)])
def test_attention_forward_variants_with_prenormalised_qk(fwd_variant, B, H, T):
    K.test_attention_with_prenormalised_qk(B, H, T)


@pytest.mark.parametrize("B,H,T", [(40, 12, 196), (2, 3, 64), (3, 2, 256), (5, 1, 100), (150, 2, 33)])
def test_attention_variants_agree_on_prenormalised_qk(B, H, T):
    """The engine's form of the call (q / k normalised by the projection GEMM, 1/||x|| as a side input): both forward kernels and
    all three backward kernels on the same inputs, three launches each - a launch must leave no state behind (the persistent
    kernels claim heads from a device-wide counter that has to be back at zero), results must repeat bit for bit where the
    kernel has a fixed summation order and agree with the other variants to bf16 rounding."""
    C, M = H * 64, B * T
    qkv = K.randn(M, 3 * C, seed=80, scale=0.5, dtype=torch.bfloat16)
    sqk = (1.0 + 0.2 * K.randn(C, seed=81)).mul(0.03)
    gb = K.randn(M, C, seed=82, scale=0.1, dtype=torch.bfloat16)
    heads = qkv[:, :2 * C].float().view(M, 2 * H, 64)
    nrm = heads.norm(dim=-1, keepdim=True)
    qkv[:, :2 * C] = (heads / nrm * (sqk / 0.03).repeat(2).view(1, 2 * H, 64)).reshape(M, 2 * C).to(torch.bfloat16)
    inv = (1.0 / nrm[..., 0]).contiguous()
    kw = dict(inv_q=inv[:, :H], inv_k=inv[:, H:])
    fwd, bwd = {}, {}
    try:
        for v in (1, 2):
            _lib.call("nvit_attention_fwd_variant", v)
            for rep in range(3):
                out = torch.zeros(M, C, device=K.DEV, dtype=torch.bfloat16)
                lse = torch.zeros(B, H, T, device=K.DEV)
                ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.03, 8.0, out, lse, B, H, T, **kw)
                torch.cuda.synchronize()
                if rep == 0:
                    fwd[v] = (out, lse)
                else:
                    assert torch.equal(out, fwd[v][0]) and torch.equal(lse, fwd[v][1]), ("forward", v, rep)
        assert K.rel(fwd[2][0], fwd[1][0]) <= 5e-3 and K.rel(fwd[2][1], fwd[1][1]) <= 1e-5
        out, lse = fwd[1]
        for v in (1, 2, 3):
            _lib.call("nvit_attention_bwd_variant", v)
            for rep in range(3):
                d = torch.zeros(M, 3 * C, device=K.DEV, dtype=torch.bfloat16)
                ds = torch.zeros(C, device=K.DEV)
                ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.03, 8.0, out, gb, lse,
                                  d[:, :C], d[:, C:2 * C], d[:, 2 * C:], ds, B, H, T, **kw)
                torch.cuda.synchronize()
                if rep == 0:
                    bwd[v] = (d, ds)
                else:
                    assert torch.equal(d, bwd[v][0]), ("backward", v, rep)
    finally:
        _lib.call("nvit_attention_fwd_variant", DEFAULT_FWD_VARIANT)
        _lib.call("nvit_attention_bwd_variant", DEFAULT_VARIANT)
    for v in (2, 3):
        assert K.rel(bwd[v][0], bwd[1][0]) <= 5e-3, (v, K.rel(bwd[v][0], bwd[1][0]))
        assert K.rel(bwd[v][1], bwd[1][1]) <= 5e-3, (v, K.rel(bwd[v][1], bwd[1][1]))
