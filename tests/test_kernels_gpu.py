"""Kernel-level parity on a B200 (-m gpu): every C-ABI entry point against the fp32 oracle arithmetic
(oracle/nvit_oracle.py restates nvit/model.py; cited there) on seeded inputs.

Tolerances (stated per north_star: bf16 kernels vs an fp32 reference): GEMM/attention outputs rel-L2 <= 1e-2,
gradients rel-L2 <= 3e-2; fp32 streaming kernels <= 1e-5 relative.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from nvit_b200 import ops, _lib  # noqa: E402
from oracle import nvit_oracle as O  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a = a.float()
    b = b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.fixture(params=[1, 2], ids=["cta_group1", "cta_group2"])
def cta_group(request):
    """Run every GEMM test with single-CTA tiles and with CTA-pair (cta_group::2) tiles."""
    _lib.call("nvit_gemm_force_cta_group", request.param)
    yield request.param
    _lib.call("nvit_gemm_force_cta_group", 0)


@pytest.fixture(params=[0, 2, 5, 8, -1], ids=["n_fastest", "bands2", "bands5", "bands8", "auto"])
def raster(request):
    """Tile order of the persistent GEMM grids (nvit_gemm_raster_group): n fastest, bands of 2 / 5 / 8 tiles along n (with
    the narrower last band whenever the tile count is not a multiple), or the automatic choice.  Results must not depend on it."""
    _lib.call("nvit_gemm_raster_group", request.param)
    yield request.param
    _lib.call("nvit_gemm_raster_group", int(os.environ.get("NVIT_GEMM_RASTER_GROUP", "0")))


GEMM_SHAPES = [
    (128, 128, 64), (256, 256, 128), (384, 768, 768), (200, 192, 192), (50, 1000, 768), (1024, 3072, 768),
    (256, 768, 3072), (333, 130, 72),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(M, N, K, a_mn, b_mn, cta_group, raster):
    # leading dims must be multiples of 8 elements for TMA: pad the storage, use views
    def padded(rows, cols, seed):
        ld = (cols + 7) // 8 * 8
        buf = randn(rows, ld, seed=seed, scale=0.5, dtype=torch.bfloat16)
        return buf, buf[:, :cols]

    abuf, a_view = padded(K, M, 1) if a_mn else padded(M, K, 1)
    bbuf, b_view = padded(K, N, 2) if b_mn else padded(N, K, 2)
    A = a_view.t() if a_mn else a_view     # logical [M,K]
    Bm = b_view.t() if b_mn else b_view    # logical [N,K]
    ref = A.float() @ Bm.float().t()
    c32 = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(abuf, bbuf, c32, M=M, N=N, K=K, lda=abuf.stride(0), ldb=bbuf.stride(0), ldc=N, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    assert rel(c32, ref) < 2e-3, (M, N, K, a_mn, b_mn, rel(c32, ref))
    ldc = (N + 7) // 8 * 8
    c16 = torch.zeros(M, ldc, device=DEV, dtype=torch.bfloat16)
    ops.gemm(abuf, bbuf, c16, M=M, N=N, K=K, lda=abuf.stride(0), ldb=bbuf.stride(0), ldc=ldc, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    assert rel(c16[:, :N], ref) < 6e-3


def test_gemm_epilogue_bias_scale_rowadd_and_bf16_copy(cta_group):
    M, N, K, T = 392, 768, 192, 196
    a = randn(M, K, seed=3, scale=0.3, dtype=torch.bfloat16)
    w = randn(N, K, seed=4, scale=0.3, dtype=torch.bfloat16)
    bias, cs, ra = randn(N, seed=5), randn(N, seed=6), randn(T, N, seed=7)
    ref = ((a.float() @ w.float().t()) + bias) * (cs * 0.5) + ra.repeat(M // T, 1)
    c = torch.empty(M, N, device=DEV)
    c2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.linear_fwd(a, w, c, bias=bias, colscale=cs, colscale_mul=0.5, rowadd=ra, rowadd_period=T, c2=c2)
    torch.cuda.synchronize()
    assert rel(c, ref) < 2e-3
    assert rel(c2, ref) < 6e-3


@pytest.mark.parametrize("splits", [0, 1, 4, 13])
def test_gemm_wgrad_splitk_and_accumulate(splits, cta_group, raster):
    M, N, K = 6272, 768, 192   # dW[N,K] = dY[M,N]^T X[M,K]
    dy = randn(M, N, seed=8, scale=0.1, dtype=torch.bfloat16)
    x = randn(M, K, seed=9, scale=0.1, dtype=torch.bfloat16)
    ref = dy.float().t() @ x.float()
    dw = torch.full((N, K), float("nan"), device=DEV)
    ops.linear_wgrad(dy, x, dw, splits=splits)
    torch.cuda.synchronize()
    assert rel(dw, ref) < 2e-3
    ops.linear_wgrad(dy, x, dw, splits=splits, accumulate=True)
    torch.cuda.synchronize()
    assert rel(dw, 2 * ref) < 2e-3


def test_gemm_dgrad_accumulates_into_fp32(cta_group):
    M, N, K = 520, 384, 192
    dy = randn(M, N, seed=10, scale=0.2, dtype=torch.bfloat16)
    w = randn(N, K, seed=11, scale=0.2, dtype=torch.bfloat16)
    base = randn(M, K, seed=12)
    dx = base.clone()
    ops.linear_dgrad(dy, w, dx, accumulate=True)
    torch.cuda.synchronize()
    assert rel(dx, base + dy.float() @ w.float()) < 2e-3


@pytest.mark.parametrize("M,C", [(300, 192), (1000, 768)])
@pytest.mark.parametrize("with_suv", [True, False])
def test_gemm_swiglu_epilogue(M, C, with_suv, cta_group, raster):
    Fh = 4 * C
    x = randn(M, C, seed=13, scale=1.0 / math.sqrt(C), dtype=torch.bfloat16)
    w = randn(2 * Fh, C, seed=14, scale=1.0, dtype=torch.bfloat16)
    suv = (1.0 + 0.1 * randn(2 * Fh, seed=15)) if with_suv else None
    mul = 0.7
    out = torch.empty(M, Fh, device=DEV, dtype=torch.bfloat16)
    raw = torch.empty(M, 2 * Fh, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, w, out, M=M, N=Fh, K=C, lda=C, ldb=C, ldc=Fh, colscale=suv, colscale_mul=mul, c2=raw, ldc2=2 * Fh, swiglu_half=Fh)
    torch.cuda.synchronize()
    uv = (x.float() @ w.float().t())
    assert rel(raw, uv) < 6e-3
    uvq = uv.bfloat16().float()
    if with_suv:
        uvq = uvq * (suv * mul)
    ref = uvq[:, :Fh] * F.silu(uvq[:, Fh:])
    assert rel(out, ref) < 1e-2


# ---------------------------------------------------------------------------------------------- residual
RESIDUAL_STAGED_DEFAULT = "2"     # the library's default: automatic choice of the form of nvit_residual_bwd


@pytest.fixture
def residual_form(request):
    """nvit_residual_bwd in both forms: rows in registers (0) and rows staged in shared memory by bulk copies (1)."""
    _lib.call("nvit_residual_bwd_staged", request.param)
    yield request.param
    _lib.call("nvit_residual_bwd_staged", int(os.environ.get("NVIT_RESIDUAL_STAGED", RESIDUAL_STAGED_DEFAULT)))


# (20011, 768) and (9001, 1024): enough rows per warp of the persistent grid for the stage ring to wrap several times
@pytest.mark.parametrize("M,C", [(64, 64), (333, 128), (777, 192), (2048, 768), (512, 1024), (20011, 768), (9001, 1024)])
@pytest.mark.parametrize("skip", [False, True])
@pytest.mark.parametrize("acc", [True, False])
@pytest.mark.parametrize("residual_form", [0, 1], indirect=True)
def test_residual_fwd_bwd(M, C, skip, acc, residual_form):
    cfg = O.named_config("micro", n_embd=C, base_scale=C ** -0.5)
    h = randn(M, C, seed=20).requires_grad_(True)
    xb = randn(M, C, seed=21, scale=0.3, dtype=torch.bfloat16)
    x = xb.float().requires_grad_(True)
    alpha = (cfg.base_scale * (1.0 + 0.3 * randn(C, seed=22))).requires_grad_(True)
    alpha_mul = 0.05 / cfg.base_scale
    h0 = randn(M, C, seed=23).requires_grad_(True) if skip else None
    sk = torch.tensor([0.9], device=DEV, requires_grad=True) if skip else None
    ref = O._nvit_residual(cfg, alpha, h, x)
    if skip:
        ref = O.justnorm(ref * sk + h0)
    g = randn(M, C, seed=24)
    ref.backward(g)

    out32 = torch.empty(M, C, device=DEV)
    out16 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.residual_fwd(h.detach(), xb, alpha.detach(), alpha_mul, out32, out16, h0=None if h0 is None else h0.detach(),
                     skip=None if sk is None else sk.detach())
    torch.cuda.synchronize()
    assert rel(out32, ref) < 1e-5
    assert rel(out16, ref) < 4e-3
    dh = torch.full((M, C), 0.25, device=DEV)
    dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    dh0 = torch.empty(M, C, device=DEV) if skip else None
    dalpha = torch.zeros(C, device=DEV)
    dskip = torch.zeros(1, device=DEV) if skip else None
    ops.residual_bwd(g, h.detach(), xb, alpha.detach(), alpha_mul, dh, dx, dalpha, dh_accumulate=acc,
                     h0=None if h0 is None else h0.detach(), skip=None if sk is None else sk.detach(), dh0=dh0, dskip=dskip)
    torch.cuda.synchronize()
    assert rel(dh - (0.25 if acc else 0.0), h.grad) < 1e-4
    assert rel(dx, x.grad) < 6e-3
    assert rel(dalpha, alpha.grad) < 1e-3
    if skip:
        assert rel(dh0, h0.grad) < 1e-4
        assert rel(dskip, sk.grad) < 1e-3


# ---------------------------------------------------------------------------------------------- original-ViT row kernels
@pytest.mark.parametrize("M,C", [(65, 64), (500, 192), (1024, 768), (300, 1024)])
@pytest.mark.parametrize("with_x", [True, False])
def test_add_rmsnorm_fwd_bwd(M, C, with_x):
    h = randn(M, C, seed=25).requires_grad_(True)
    xb = randn(M, C, seed=26, scale=0.5, dtype=torch.bfloat16) if with_x else None
    x = xb.float().requires_grad_(True) if with_x else None
    w = (1 + 0.2 * randn(C, seed=27)).requires_grad_(True)
    t = h + x if with_x else h
    ref = O.rmsnorm(t, w)
    gy = randn(M, C, seed=28)
    ref.backward(gy)
    y32 = torch.empty(M, C, device=DEV)
    y16 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.add_rmsnorm_fwd(h.detach(), xb, w.detach(), 1e-6, y32, y16)
    dh = torch.full((M, C), 0.5, device=DEV)
    dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16) if with_x else None
    dw = torch.zeros(C, device=DEV)
    ops.add_rmsnorm_bwd(gy, h.detach(), xb, w.detach(), 1e-6, dh, dx, dw, dh_accumulate=True)
    torch.cuda.synchronize()
    assert rel(y32, ref) < 1e-5 and rel(y16, ref) < 4e-3
    assert rel(dh - 0.5, h.grad) < 1e-4
    assert rel(dw, w.grad) < 1e-3
    if with_x:
        assert rel(dx, x.grad) < 6e-3


@pytest.mark.parametrize("M,C", [(65, 64), (500, 192), (1024, 768)])
def test_add_skipnorm_fwd_bwd(M, C):
    h = randn(M, C, seed=35).requires_grad_(True)
    xb = randn(M, C, seed=36, scale=0.5, dtype=torch.bfloat16)
    x = xb.float().requires_grad_(True)
    h0 = randn(M, C, seed=37).requires_grad_(True)
    sk = torch.tensor([0.8], device=DEV, requires_grad=True)
    ref = O.justnorm((h + x) * sk + h0)
    g = randn(M, C, seed=38)
    ref.backward(g)
    o32 = torch.empty(M, C, device=DEV)
    o16 = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.add_skipnorm_fwd(h.detach(), xb, h0.detach(), sk.detach(), o32, o16)
    dh, dh0 = torch.empty(M, C, device=DEV), torch.empty(M, C, device=DEV)
    dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    dsk = torch.zeros(1, device=DEV)
    ops.add_skipnorm_bwd(g, h.detach(), xb, h0.detach(), sk.detach(), dh, dx, dh0, dsk)
    torch.cuda.synchronize()
    assert rel(o32, ref) < 1e-5 and rel(o16, ref) < 4e-3
    assert rel(dh, h.grad) < 1e-4 and rel(dh0, h0.grad) < 1e-4 and rel(dx, x.grad) < 6e-3
    assert rel(dsk, sk.grad) < 1e-3


# ---------------------------------------------------------------------------------------------- swiglu (unfused)
@pytest.mark.parametrize("with_suv", [True, False])
def test_swiglu_fwd_bwd(with_suv):
    M, Fh = 515, 768
    uvb = randn(M, 2 * Fh, seed=30, dtype=torch.bfloat16)
    uv = uvb.float().requires_grad_(True)
    suv = (1.0 + 0.2 * randn(2 * Fh, seed=31)).requires_grad_(True) if with_suv else None
    mul = 1.7
    s = uv * (suv * mul) if with_suv else uv
    ref = s[:, :Fh] * F.silu(s[:, Fh:])
    g = randn(M, Fh, seed=32, dtype=torch.bfloat16)
    ref.backward(g.float())
    x = torch.empty(M, Fh, device=DEV, dtype=torch.bfloat16)
    ops.swiglu_fwd(uvb, None if suv is None else suv.detach(), mul, x)
    duv = torch.empty(M, 2 * Fh, device=DEV, dtype=torch.bfloat16)
    dsuv = torch.zeros(2 * Fh, device=DEV) if with_suv else None
    ops.swiglu_bwd(g, uvb, None if suv is None else suv.detach(), mul, duv, dsuv)
    torch.cuda.synchronize()
    assert rel(x, ref) < 5e-3
    assert rel(duv, uv.grad) < 5e-3
    if with_suv:
        assert rel(dsuv, suv.grad) < 1e-3


# ---------------------------------------------------------------------------------------------- gate backward in the dgrad GEMM
@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("M,Fh,K,with_suv", [(515, 768, 192, True), (1000, 3072, 768, True), (300, 64, 64, False), (4096, 768, 768, False),
                                             (129, 256, 136, True)])
def test_gemm_gate_backward(M, Fh, K, with_suv, cta_group, raster):
    """nvit_gemm_gate_bwd: d(uv_raw) = gate backward of dx = dY W against the saved raw u|v, dx never leaving the SM."""
    from nvit_b200 import _lib
    _lib.call("nvit_gemm_force_cta_group", cta_group)
    try:
        dy = randn(M, K, seed=40, dtype=torch.bfloat16) * 0.5
        w = (randn(K, Fh, seed=41) * (1.0 / K ** 0.5)).to(torch.bfloat16)
        uvb = randn(M, 2 * Fh, seed=42, dtype=torch.bfloat16)
        uv = uvb.float().requires_grad_(True)
        suv = (1.0 + 0.2 * randn(2 * Fh, seed=43)).requires_grad_(True) if with_suv else None
        mul = 1.3
        s = uv * (suv * mul) if with_suv else uv
        x = s[:, :Fh] * F.silu(s[:, Fh:])
        dx = dy.float() @ w.float()
        x.backward(dx)
        duv = torch.full((M, 2 * Fh), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.gemm_gate_bwd(dy, w, uvb, None if suv is None else suv.detach(), mul, duv)
        torch.cuda.synchronize()
        assert torch.isfinite(duv.float()).all()
        assert rel(duv, uv.grad) < 6e-3, rel(duv, uv.grad)
        # the two-kernel composition it replaces (dx rounded to bf16 in between) agrees to bf16 rounding
        dxb = torch.empty(M, Fh, device=DEV, dtype=torch.bfloat16)
        ops.linear_dgrad(dy, w, dxb)
        duv2 = torch.empty_like(duv)
        dsuv = torch.zeros(2 * Fh, device=DEV) if with_suv else None
        ops.swiglu_bwd(dxb, uvb, None if suv is None else suv.detach(), mul, duv2, dsuv)
        assert rel(duv, duv2) < 8e-3
    finally:
        _lib.call("nvit_gemm_force_cta_group", 0)


def test_rowdot_div_gives_the_suv_gradient():
    """dL/dsuv[c] = W[c,:] . dW[c,:] / suv[c] when uv = (suv mul) * (h W^T): checked against autograd through the whole gate."""
    M, C, Fh = 700, 192, 384
    h = randn(M, C, seed=50)
    W = (randn(2 * Fh, C, seed=51) * C ** -0.5).requires_grad_(True)
    suv = (1.0 + 0.3 * randn(2 * Fh, seed=52)).requires_grad_(True)
    mul = C ** 0.5
    uvs = (h @ W.t()) * (suv * mul)
    x = uvs[:, :Fh] * F.silu(uvs[:, Fh:])
    (x * randn(M, Fh, seed=53)).sum().backward()
    out = torch.full((2 * Fh,), float("nan"), device=DEV)
    ops.rowdot_div(W.detach().contiguous(), W.grad.contiguous(), suv.detach(), out)
    assert rel(out, suv.grad) < 1e-4
    suv0 = suv.detach().clone()
    suv0[3] = 0.0
    ops.rowdot_div(W.detach().contiguous(), W.grad.contiguous(), suv0, out)
    assert float(out[3]) == 0.0 and torch.isfinite(out).all()


# ---------------------------------------------------------------------------------------------- q/k normalisation in the GEMM
@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("M,C,K,norm_cols,with_bias", [(515, 192, 192, 128, False), (1000, 768, 768, 512, True), (300, 64, 64, 64, False),
                                                       (196, 2304, 768, 1536, False)])
def test_gemm_qknorm(M, C, K, norm_cols, with_bias, cta_group, raster):
    """nvit_gemm_qknorm: projection + per-head unit norm + sqk scale in the epilogue, 1/||x|| as a side output."""
    from nvit_b200 import _lib
    _lib.call("nvit_gemm_force_cta_group", cta_group)
    try:
        N = C
        period = 64 * max(1, norm_cols // 128)
        x = randn(M, K, seed=60, dtype=torch.bfloat16)
        w = (randn(N, K, seed=61) * K ** -0.5).to(torch.bfloat16)
        bias = 0.1 * randn(N, seed=62) if with_bias else None
        scale = 1.0 + 0.3 * randn(period, seed=63)
        mul = 1.7
        y = x.float() @ w.float().t()
        if bias is not None:
            y = y + bias
        heads = y[:, :norm_cols].reshape(M, norm_cols // 64, 64)
        nrm = heads.norm(dim=-1, keepdim=True)
        s_full = (scale * mul).repeat(norm_cols // period).view(1, norm_cols // 64, 64)
        ref = torch.cat([(heads / nrm * s_full).reshape(M, norm_cols), y[:, norm_cols:]], dim=1)
        out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        inv = torch.full((M, norm_cols // 64 + 3), float("nan"), device=DEV)
        ops.gemm_qknorm(x, w, out, scale, mul, period, norm_cols, inv, bias=bias)
        torch.cuda.synchronize()
        assert rel(out, ref) < 6e-3, rel(out, ref)
        assert rel(inv[:, :norm_cols // 64], 1.0 / nrm[..., 0]) < 2e-3
        assert torch.isnan(inv[:, norm_cols // 64:]).all()
    finally:
        _lib.call("nvit_gemm_force_cta_group", 0)


# ---------------------------------------------------------------------------------------------- attention
def attention_reference(q, k, v, sqk, sqk_mul, scale, B, H, T):
    D = 64
    def heads(x):
        return x.view(B, T, H, D).transpose(1, 2)
    qh, kh, vh = heads(q), heads(k), heads(v)
    if sqk is not None:
        s = (sqk * sqk_mul).view(1, H, 1, D)
        qh = s * O.justnorm(qh)
        kh = s * O.justnorm(kh)
    att = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1) @ vh
    return att.transpose(1, 2).reshape(B * T, H * D)


@pytest.mark.parametrize("B,H,T", [(2, 1, 64), (3, 3, 64), (2, 2, 196), (1, 12, 196), (2, 2, 16), (1, 2, 256), (2, 1, 130)])
@pytest.mark.parametrize("normed", [True, False])
def test_attention_fwd_bwd(B, H, T, normed):
    C = H * 64
    M = B * T
    qkvb = randn(M, 3 * C, seed=40, scale=0.5, dtype=torch.bfloat16)
    qkv = qkvb.float().requires_grad_(True)
    sqk = (1.0 + 0.2 * randn(C, seed=41)).mul(0.03).requires_grad_(True) if normed else None
    sqk_mul = 1.0 / 0.03
    scale = 8.0 if normed else 0.125
    ref = attention_reference(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, sqk_mul, scale, B, H, T)
    gb = randn(M, C, seed=42, scale=0.1, dtype=torch.bfloat16)
    ref.backward(gb.float())

    out = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=DEV)
    sq = None if sqk is None else sqk.detach()
    ops.attention_fwd(qkvb[:, :C], qkvb[:, C:2 * C], qkvb[:, 2 * C:], sq, sqk_mul, scale, out, lse, B, H, T)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-2, rel(out, ref)
    dqkv = torch.zeros(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    dsqk = torch.zeros(C, device=DEV) if normed else None
    ops.attention_bwd(qkvb[:, :C], qkvb[:, C:2 * C], qkvb[:, 2 * C:], sq, sqk_mul, scale, out, gb, lse,
                      dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T)
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        assert rel(dqkv[:, sl], qkv.grad[:, sl]) < 3e-2, (name, rel(dqkv[:, sl], qkv.grad[:, sl]))
    if normed:
        assert rel(dsqk, sqk.grad) < 3e-2


@pytest.mark.parametrize("B,H,T", [(2, 2, 196), (3, 1, 64), (1, 2, 256)])
def test_attention_with_prenormalised_qk(B, H, T):
    """inv_q / inv_k: q, k arrive as s * x / ||x|| (what nvit_gemm_qknorm writes) with 1/||x|| on the side; outputs and the
    gradients w.r.t. the RAW q, k must match the in-kernel normalisation path and the reference."""
    C = H * 64
    M = B * T
    qkvb = randn(M, 3 * C, seed=44, scale=0.5, dtype=torch.bfloat16)
    qkv = qkvb.float().requires_grad_(True)
    sqk = (1.0 + 0.2 * randn(C, seed=45)).mul(0.03).requires_grad_(True)
    sqk_mul, scale = 1.0 / 0.03, 8.0
    ref = attention_reference(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, sqk_mul, scale, B, H, T)
    gb = randn(M, C, seed=46, scale=0.1, dtype=torch.bfloat16)
    ref.backward(gb.float())
    with torch.no_grad():
        heads = qkvb[:, :2 * C].float().view(M, 2 * H, 64)
        nrm = heads.norm(dim=-1, keepdim=True)
        s = (sqk.detach() * sqk_mul).view(1, H, 64).repeat(1, 2, 1)
        pre = qkvb.clone()
        pre[:, :2 * C] = (heads / nrm * s).reshape(M, 2 * C).to(torch.bfloat16)
        inv = (1.0 / nrm[..., 0]).contiguous()                       # [M, 2H]: q heads, then k heads
    out = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=DEV)
    ops.attention_fwd(pre[:, :C], pre[:, C:2 * C], pre[:, 2 * C:], sqk.detach(), sqk_mul, scale, out, lse, B, H, T,
                      inv_q=inv[:, :H], inv_k=inv[:, H:])
    assert rel(out, ref) < 1e-2, rel(out, ref)
    dqkv = torch.zeros(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    dsqk = torch.zeros(C, device=DEV)
    ops.attention_bwd(pre[:, :C], pre[:, C:2 * C], pre[:, 2 * C:], sqk.detach(), sqk_mul, scale, out, gb, lse,
                      dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T, inv_q=inv[:, :H], inv_k=inv[:, H:])
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        assert rel(dqkv[:, sl], qkv.grad[:, sl]) < 3e-2, (name, rel(dqkv[:, sl], qkv.grad[:, sl]))
    assert rel(dsqk, sqk.grad) < 3e-2


# ---------------------------------------------------------------------------------------------- misc kernels
@pytest.mark.parametrize("P,S,B", [(4, 32, 3), (16, 224, 2), (8, 64, 2)])
def test_im2col_local_and_global(P, S, B):
    cfg = O.named_config("micro", image_size=S, local_patch_size=P, global_patch_size=2 * P)
    img = randn(B, 3, S, S, seed=50)
    g = S // P
    loc = torch.empty(B * g * g, 3 * P * P, device=DEV, dtype=torch.bfloat16)
    ops.im2col(img, loc, P, P, 0)
    assert rel(loc, O.reconstruction_target(cfg, img).reshape(B * g * g, -1)) < 4e-3
    G = 2 * P
    glo = torch.empty(B * g * g, 3 * G * G, device=DEV, dtype=torch.bfloat16)
    ops.im2col(img, glo, G, P, (G - P) // 2)
    pad = (G - P) // 2
    ref = F.unfold(F.pad(img, (pad,) * 4, mode="reflect"), kernel_size=G, stride=P).transpose(1, 2).reshape(B * g * g, -1)
    torch.cuda.synchronize()
    assert rel(glo, ref) < 4e-3
    # as a conv operand: im2col @ W^T == conv2d
    w = randn(64, 3, G, G, seed=51, scale=0.05)
    conv = F.conv2d(F.pad(img, (pad,) * 4, mode="reflect"), w, stride=P).flatten(2).transpose(1, 2).reshape(B * g * g, 64)
    assert rel(glo.float() @ w.reshape(64, -1).t(), conv) < 6e-3


def test_pool_ln_head_fwd_bwd():
    B, T, C = 5, 196, 768
    h = randn(B, T, C, seed=60).requires_grad_(True)
    gamma = (1 + 0.1 * randn(C, seed=61)).requires_grad_(True)
    beta = (0.1 * randn(C, seed=62)).requires_grad_(True)
    ref = F.layer_norm(h.mean(1), (C,), gamma, beta, eps=1e-5)
    gy = randn(B, C, seed=63, dtype=torch.bfloat16)
    ref.backward(gy.float())
    y = torch.empty(B, C, device=DEV, dtype=torch.bfloat16)
    xhat = torch.empty(B, C, device=DEV)
    rstd = torch.empty(B, device=DEV)
    ops.pool_ln_fwd(h.detach(), gamma.detach(), beta.detach(), 1e-5, y, xhat, rstd, B, T, C)
    dh = torch.empty(B, T, C, device=DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.pool_ln_bwd(gy, gamma.detach(), xhat, rstd, dh, dg, db, B, T, C)
    torch.cuda.synchronize()
    assert rel(y, ref) < 4e-3
    assert rel(dh, h.grad) < 1e-4
    assert rel(dg, gamma.grad) < 1e-4 and rel(db, beta.grad) < 1e-4


def test_cross_entropy_and_head_scale():
    B, N = 37, 1000
    logits = randn(B, N, seed=70, scale=2.0).requires_grad_(True)
    y = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(1)).to(DEV)
    ref = F.cross_entropy(logits, y)
    ref.backward()
    loss = torch.zeros(1, device=DEV)
    dl = torch.empty(B, N, device=DEV)
    ops.cross_entropy(logits.detach(), y, loss, dl, 1.0)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    assert rel(dl, logits.grad) < 1e-4
    raw = randn(B, N, seed=71).requires_grad_(True)
    sz = (1 + 0.1 * randn(N, seed=72)).requires_grad_(True)
    (raw * (sz * 1.5)).backward(dl)
    draw = torch.zeros(B, 1008, device=DEV, dtype=torch.bfloat16)
    dsz = torch.zeros(N, device=DEV)
    ops.head_scale_bwd(dl, raw.detach(), sz.detach(), 1.5, draw, dsz)
    torch.cuda.synchronize()
    assert rel(draw[:, :N], raw.grad) < 4e-3
    assert rel(dsz, sz.grad) < 1e-4


def test_small_reductions_and_cast():
    x = randn(1000003, seed=80)
    out = torch.zeros(1, device=DEV)
    ops.sumsq(x, out)
    xb = torch.empty(1000003, device=DEV, dtype=torch.bfloat16)
    ops.cast_bf16(x, xb)
    m = randn(999, 520, seed=81, dtype=torch.bfloat16)
    cs = torch.zeros(520, device=DEV)
    ops.colsum(m, cs)
    dx = randn(6, 49, 200, seed=82)
    dpos = torch.zeros(49, 200, device=DEV)
    dbias = torch.zeros(200, device=DEV)
    ops.pos_bias_grad(dx, 6, 49, 200, dpos, dbias)
    pred = randn(4097, seed=83, dtype=torch.bfloat16)
    tgt = randn(4097, seed=84, dtype=torch.bfloat16)
    mse = torch.zeros(1, device=DEV)
    ops.tanh_mse(pred, tgt, mse)
    torch.cuda.synchronize()
    assert abs(float(out) - float((x.double() ** 2).sum())) < 1e-4 * float(out)
    assert torch.equal(xb, x.bfloat16())
    assert rel(cs, m.float().sum(0)) < 1e-5
    assert rel(dpos, dx.sum(0)) < 1e-5 and rel(dbias, dx.sum((0, 1))) < 1e-5
    assert abs(float(mse) - float(F.mse_loss(torch.tanh(pred.float()), tgt.float()))) < 1e-4 * float(mse)


def test_adamw_flat_matches_torch_with_clip():
    n, n_decay = 100003, 60000
    p0 = randn(n, seed=90)
    pa = p0[:n_decay].clone().requires_grad_(True)
    pb = p0[n_decay:].clone().requires_grad_(True)
    opt = torch.optim.AdamW([{"params": [pa], "weight_decay": 0.1}, {"params": [pb], "weight_decay": 0.0}], lr=1e-3, betas=(0.9, 0.95))
    p = p0.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = randn(n, seed=90 + step, scale=0.01 * step)
        pa.grad, pb.grad = g[:n_decay].clone(), g[n_decay:].clone()
        torch.nn.utils.clip_grad_norm_([pa, pb], 1.0)
        opt.step()
        gn = torch.zeros(1, device=DEV)
        ops.sumsq(g, gn)
        ops.adamw_flat(p, g, m, v, n_decay, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, gn, 1.0)
    torch.cuda.synchronize()
    ref = torch.cat([pa.detach(), pb.detach()])
    assert float((p - ref).abs().max()) < 2e-6


def test_weight_norm_multi_rows_and_columns():
    shapes = [(768, 768, 1), (768, 768, 0), (6144, 768, 1), (768, 3072, 0), (100, 36, 1), (36, 100, 0), (50, 37, 1), (37, 50, 0),
              (64, 2048, 1), (1024, 4096, 0)]
    ws = [randn(r, c, seed=100 + i) for i, (r, c, _) in enumerate(shapes)]
    w16 = [torch.empty(r, c, device=DEV, dtype=torch.bfloat16) for (r, c, _) in shapes]
    rows, first = [], 0
    for w, h16, (r, c, axis) in zip(ws, w16, shapes):
        rows.append([w.data_ptr(), h16.data_ptr(), r, c, axis, first])
        first += (r + 7) // 8 if axis == 1 else (c + 127) // 128
    table = torch.tensor(rows, dtype=torch.int64, device=DEV)
    refs = [w / w.norm(p=2, dim=axis, keepdim=True) for w, (_, _, axis) in zip(ws, shapes)]
    ops.weight_norm_multi(table, len(shapes), first)
    torch.cuda.synchronize()
    for w, h16, ref, (_, _, axis) in zip(ws, w16, refs, shapes):
        assert rel(w, ref) < 1e-6
        assert float((w.norm(dim=axis) - 1).abs().max()) < 1e-5      # north_star: unit norm to 1e-3
        assert rel(h16, ref) < 4e-3
    # idempotence
    before = [w.clone() for w in ws]
    ops.weight_norm_multi(table, len(shapes), first)
    torch.cuda.synchronize()
    for w, b in zip(ws, before):
        assert float((w - b).abs().max()) < 1e-6


def test_sumsq_deterministic_form():
    """nvit_sumsq_f32_det: the value of nvit_sumsq_f32 (and torch), accumulated into out, bit-identical from run to run and
    self-cleaning (the ticket word is zero again after every launch)."""
    for n in (1, 1023, 1 << 20, 30_000_001):
        x = randn(n, seed=90)
        ref = float((x.double() ** 2).sum())
        ws = torch.zeros(1024, device=DEV)
        outs = []
        for rep in range(4):
            out = torch.full((1,), 2.0, device=DEV)
            ops.sumsq_det(x, out, ws)
            torch.cuda.synchronize()
            outs.append(out.clone())
            assert float(ws[0]) == 0.0
        assert abs(float(outs[0]) - 2.0 - ref) <= 1e-5 * ref + 1e-6, (n, float(outs[0]), ref)
        assert all(torch.equal(o, outs[0]) for o in outs), n
        # a small workspace limits the grid, not the result
        out2 = torch.zeros(1, device=DEV)
        ops.sumsq_det(x, out2, torch.zeros(9, device=DEV))
        assert abs(float(out2) - ref) <= 1e-5 * ref + 1e-6
