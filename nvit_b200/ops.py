"""Thin tensor-level wrappers over the C ABI (include/nvit_b200.h).

PyTorch is used for device memory and streams only: every function here takes CUDA tensors, passes their raw
pointers plus the current CUDA stream to libnvit_b200.so and returns nothing (outputs are caller-allocated).
No function has a PyTorch/CPU fallback; non-CUDA tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _p(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("nvit_b200 ops need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")


def gemm(a, b, c, *, M, N, K, lda, ldb, ldc, a_mn=False, b_mn=False, accumulate=False, splits=1, bias=None,
         colscale=None, colscale_mul=1.0, rowadd=None, rowadd_period=0, c2=None, ldc2=0, swiglu_half=0):
    """c[M,N] (+)= a @ b^T on tcgen05 (see nvit_gemm_bf16).  c.dtype selects fp32 or bf16 output."""
    out_f32 = 1 if c.dtype == F32 else 0
    _lib.call("nvit_gemm_bf16", _p(a), _p(b), _p(c), _p(c2), M, N, K, lda, ldb, ldc, ldc2, int(a_mn), int(b_mn), out_f32,
              int(accumulate), splits, _p(bias), _p(colscale), float(colscale_mul), _p(rowadd), rowadd_period, swiglu_half,
              _stream())


def linear_fwd(x, w, out, *, bias=None, colscale=None, colscale_mul=1.0, rowadd=None, rowadd_period=0, c2=None):
    """out[M,N] = x[M,K] @ w[N,K]^T (+ epilogue).  x, w bf16 row-major (views with a row pitch are fine)."""
    M, K = x.shape
    N = w.shape[0]
    gemm(x, w, out, M=M, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldc=out.stride(0), bias=bias, colscale=colscale,
         colscale_mul=colscale_mul, rowadd=rowadd, rowadd_period=rowadd_period, c2=c2, ldc2=(c2.stride(0) if c2 is not None else 0))


def linear_dgrad(dy, w, dx, *, accumulate=False):
    """dx[M,K] (+)= dy[M,N] @ w[N,K]  — B operand is w as it lies (MN-major)."""
    M, N = dy.shape
    K = w.shape[1]
    gemm(dy, w, dx, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldc=dx.stride(0), b_mn=True, accumulate=accumulate)


def linear_wgrad(dy, x, dw, *, splits=1, accumulate=False):
    """dw[N,K] (+)= dy[M,N]^T @ x[M,K] — both operands read as they lie (MN-major), fp32 output."""
    M, N = dy.shape
    K = x.shape[1]
    gemm(dy, x, dw, M=N, N=K, K=M, lda=dy.stride(0), ldb=x.stride(0), ldc=dw.stride(0), a_mn=True, b_mn=True, splits=splits,
         accumulate=accumulate)


def gemm_qknorm(x, w, out, scale, scale_mul, scale_period, norm_cols, inv, *, bias=None):
    """out[M,N] = x @ w^T (+bias) with columns [0, norm_cols) normalised per 64-column head and scaled (nvit_gemm_qknorm);
    inv[M, norm_cols / 64] receives 1/||x||."""
    M, K = x.shape
    N = w.shape[0]
    _lib.call("nvit_gemm_qknorm", _p(x), _p(w), _p(out), M, N, K, x.stride(0), w.stride(0), out.stride(0), _p(bias), _p(scale),
              float(scale_mul), scale_period, norm_cols, _p(inv), inv.stride(0), _stream())


def gemm_gate_bwd(dy, w, uv, suv, suv_mul, duv):
    """duv[M, 2F] = gate backward of (dy[M,K] @ w[K,F]) against the saved raw u|v (see nvit_gemm_gate_bwd)."""
    M, K = dy.shape
    F = w.shape[1]
    _lib.call("nvit_gemm_gate_bwd", _p(dy), _p(w), _p(uv), _p(suv), float(suv_mul), _p(duv), M, F, K, dy.stride(0), w.stride(0),
              uv.stride(0), duv.stride(0), _stream())


def rowdot_div(w, dw, div, out):
    rows = w.shape[0]
    _lib.call("nvit_rowdot_div", _p(w), _p(dw), _p(div), _p(out), rows, w.numel() // rows, _stream())


def cast_bf16(src, dst):
    _lib.call("nvit_cast_f32_to_bf16", _p(src), _p(dst), src.numel(), _stream())


def sumsq(x, out):
    _lib.call("nvit_sumsq_f32", _p(x), x.numel(), _p(out), _stream())


def sumsq_det(x, out, workspace):
    """out[0] += sum(x^2), reproducible to the bit (nvit_sumsq_f32_det); workspace: fp32 scratch, [0] zero before first use."""
    _lib.call("nvit_sumsq_f32_det", _p(x), x.numel(), _p(out), _p(workspace), workspace.numel(), _stream())


def colsum(x, out):
    _lib.call("nvit_colsum_bf16", _p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), _stream())


def pos_bias_grad(dx, B, T, C, dpos, dbias):
    _lib.call("nvit_pos_bias_grad", _p(dx), B, T, C, _p(dpos), _p(dbias), _stream())


def residual_fwd(h, x, alpha, alpha_mul, out32, out16, h0=None, skip=None):
    M, C = h.shape
    _lib.call("nvit_residual_fwd", _p(h), _p(x), _p(alpha), float(alpha_mul), _p(h0), _p(skip), _p(out32), _p(out16), M, C, _stream())


def residual_bwd(g, h, x, alpha, alpha_mul, dh, dx, dalpha, *, dh_accumulate=False, h0=None, skip=None, dh0=None, dskip=None):
    M, C = h.shape
    _lib.call("nvit_residual_bwd", _p(g), _p(h), _p(x), _p(alpha), float(alpha_mul), _p(h0), _p(skip), _p(dh), int(dh_accumulate),
              _p(dx), _p(dh0), _p(dalpha), _p(dskip), M, C, _stream())


def add_rmsnorm_fwd(h, x, w, eps, y32, y16):
    M, C = h.shape
    _lib.call("nvit_add_rmsnorm_fwd", _p(h), _p(x), _p(w), float(eps), _p(y32), _p(y16), M, C, _stream())


def add_rmsnorm_bwd(dy, h, x, w, eps, dh, dx, dw, *, dh_accumulate=False):
    M, C = h.shape
    _lib.call("nvit_add_rmsnorm_bwd", _p(dy), _p(h), _p(x), _p(w), float(eps), _p(dh), int(dh_accumulate), _p(dx), _p(dw), M, C, _stream())


def add_skipnorm_fwd(h, x, h0, skip, out32, out16):
    M, C = h.shape
    _lib.call("nvit_add_skipnorm_fwd", _p(h), _p(x), _p(h0), _p(skip), _p(out32), _p(out16), M, C, _stream())


def add_skipnorm_bwd(g, h, x, h0, skip, dh, dx, dh0, dskip):
    M, C = h.shape
    _lib.call("nvit_add_skipnorm_bwd", _p(g), _p(h), _p(x), _p(h0), _p(skip), _p(dh), _p(dx), _p(dh0), _p(dskip), M, C, _stream())


def swiglu_fwd(uv, suv, suv_mul, x):
    M, F2 = uv.shape
    _lib.call("nvit_swiglu_fwd", _p(uv), _p(suv), float(suv_mul), _p(x), M, F2 // 2, _stream())


def swiglu_bwd(dx, uv, suv, suv_mul, duv, dsuv):
    M, F2 = uv.shape
    _lib.call("nvit_swiglu_bwd", _p(dx), _p(uv), _p(suv), float(suv_mul), _p(duv), _p(dsuv), M, F2 // 2, _stream())


def attention_fwd(q, k, v, sqk, sqk_mul, scale, out, lse, B, H, T, D=64, inv_q=None, inv_k=None):
    """inv_q / inv_k ([M, *] fp32 views, head h at column h): q / k are already normalised (gemm_qknorm)."""
    _lib.call("nvit_attention_fwd", _p(q), _p(k), _p(v), q.stride(0), k.stride(0), v.stride(0), _p(sqk), float(sqk_mul), float(scale),
              _p(out), out.stride(0), _p(lse), B, H, T, D, _p(inv_q), _p(inv_k), 0 if inv_q is None else inv_q.stride(0),
              0 if inv_k is None else inv_k.stride(0), _stream())


def attention_bwd(q, k, v, sqk, sqk_mul, scale, out, dout, lse, dq, dk, dv, dsqk, B, H, T, D=64, inv_q=None, inv_k=None):
    assert out.stride(0) == dout.stride(0)
    _lib.call("nvit_attention_bwd", _p(q), _p(k), _p(v), q.stride(0), k.stride(0), v.stride(0), _p(sqk), float(sqk_mul), float(scale),
              _p(out), _p(dout), out.stride(0), _p(lse), _p(dq), _p(dk), _p(dv), dq.stride(0), dk.stride(0), dv.stride(0), _p(dsqk),
              B, H, T, D, _p(inv_q), _p(inv_k), 0 if inv_q is None else inv_q.stride(0), 0 if inv_k is None else inv_k.stride(0),
              _stream())


def im2col(img, out, ksize, stride, pad):
    B, ch, S, _ = img.shape
    _lib.call("nvit_im2col_bf16", _p(img), _p(out), B, ch, S, ksize, stride, pad, _stream())


def im2col_u8(img_u8, out, ksize, stride, pad, scale, shift):
    """img_u8: [B, S, S, ch] uint8 (HWC); value = pixel * scale + shift (ToTensor + Normalize folded in)."""
    B, S, _, ch = img_u8.shape
    if img_u8.dtype != torch.uint8 or not img_u8.is_contiguous():
        raise TypeError("im2col_u8: expected a contiguous uint8 [B, S, S, ch] tensor")
    _lib.call("nvit_im2col_u8", _p(img_u8), _p(out), B, ch, S, ksize, stride, pad, float(scale), float(shift), _stream())


def pool_ln_fwd(h, gamma, beta, eps, y, xhat, rstd, B, T, C):
    _lib.call("nvit_pool_ln_fwd", _p(h), _p(gamma), _p(beta), float(eps), _p(y), _p(xhat), _p(rstd), B, T, C, _stream())


def pool_ln_bwd(dy, gamma, xhat, rstd, dh, dgamma, dbeta, B, T, C):
    _lib.call("nvit_pool_ln_bwd", _p(dy), _p(gamma), _p(xhat), _p(rstd), _p(dh), _p(dgamma), _p(dbeta), B, T, C, _stream())


def head_scale_bwd(dlogits, raw, sz, sz_mul, draw, dsz):
    B, N = dlogits.shape
    _lib.call("nvit_head_scale_bwd", _p(dlogits), _p(raw), _p(sz), float(sz_mul), _p(draw), _p(dsz), B, N, draw.stride(0), _stream())


def augment_u8(img_u8, out, ops_i32, params_f32):
    """Per-image two-operation AutoAugment on a uint8 [B, S, S, 3] batch (nvit_b200/augment.py draws `ops` [B, 2] int32 and
    `params` [B, 2, 8] float32; train.py:1081-1092).  `out` is a second uint8 tensor of the same shape."""
    if img_u8.dim() != 4 or img_u8.dtype != torch.uint8 or not img_u8.is_contiguous():
        raise TypeError("augment_u8: expected a contiguous uint8 [B, S, S, 3] tensor")
    B, S, S2, ch = img_u8.shape
    if S != S2:
        raise ValueError("augment_u8: images must be square")
    if out.shape != img_u8.shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != img_u8.device:
        raise TypeError("augment_u8: out must be a contiguous uint8 tensor of the input's shape on the same device")
    if out.data_ptr() == img_u8.data_ptr():
        raise ValueError("augment_u8: out must not alias the input")
    _chk(ops_i32, torch.int32, "augment_u8: ops")
    _chk(params_f32, F32, "augment_u8: params")
    if ops_i32.device != img_u8.device or params_f32.device != img_u8.device:
        raise ValueError("augment_u8: ops and params must live on the batch's device")
    if tuple(ops_i32.shape) != (B, 2) or tuple(params_f32.shape) != (B, 2, 8):
        raise ValueError(f"augment_u8: ops must be [{B}, 2] and params [{B}, 2, 8]")
    _lib.call("nvit_augment_u8", _p(img_u8), _p(out), _p(ops_i32), _p(params_f32), B, S, ch, _stream())


def cross_entropy(logits, target, loss, dlogits, gscale=1.0):
    """Mean softmax cross-entropy (train.py:906).  Targets are int64 class indices; a target outside [0, N) is treated like
    F.cross_entropy's ignore_index: no loss, zero gradient row (the mean still divides by B)."""
    B, N = logits.shape
    _chk(logits, F32, "cross_entropy: logits")
    _chk(target, torch.int64, "cross_entropy: target")
    if target.shape != (B,):
        raise ValueError(f"cross_entropy: target must have shape ({B},), got {tuple(target.shape)}")
    if dlogits is not None:
        _chk(dlogits, F32, "cross_entropy: dlogits")
        if dlogits.shape != logits.shape:
            raise ValueError("cross_entropy: dlogits must match logits")
    _lib.call("nvit_cross_entropy", _p(logits), _p(target), _p(loss), _p(dlogits), float(gscale), B, N, _stream())


def tanh_mse(pred, target, out):
    n = pred.numel()
    _lib.call("nvit_tanh_mse", _p(pred), _p(target), n, 1.0 / n, _p(out), _stream())


def tanh_mse_bwd(pred, target, weight_dev, dpred):
    n = pred.numel()
    _lib.call("nvit_tanh_mse_bwd", _p(pred), _p(target), n, 1.0 / n, _p(weight_dev), _p(dpred), _stream())


# ---- Kohonen maps (BASELINE config 5)
def split_bf16(x, hi, lo):
    _lib.call("nvit_split_bf16", _p(x), _p(hi), _p(lo), x.numel(), _stream())


def som_prepare(nodes, node_sq, hi, lo, snapshot):
    G, C = nodes.shape
    _lib.call("nvit_som_prepare", _p(nodes), G, C, _p(node_sq), _p(hi), _p(lo), _p(snapshot), _stream())


def som_select(dots, node_sq, nodes, idx, idx64, onehot, counts, repr32, repr16):
    M, G = dots.shape
    _lib.call("nvit_som_select", _p(dots), _p(node_sq), _p(nodes), M, G, nodes.shape[1], _p(idx), _p(idx64), _p(onehot), _p(counts),
              _p(repr32), _p(repr16), _stream())


def som_pool(x, rows, run, out):
    _lib.call("nvit_som_pool", _p(x), rows, run, _p(out), _stream())


def som_update(nodes, pooled, bmu, steps, grid_rows, grid_cols, coef_dev, sigma):
    _lib.call("nvit_som_update", _p(nodes), _p(pooled), _p(bmu), steps, grid_rows, grid_cols, nodes.shape[1], _p(coef_dev), float(sigma),
              _stream())


def som_pair_losses(repr_l, repr_g, x_l, x_g, sums, weights=None, d_repr_l=None, d_repr_g=None, d_x_l=None, d_x_g=None):
    M, C = repr_l.shape
    _lib.call("nvit_som_pair_losses", _p(repr_l), _p(repr_g), _p(x_l), _p(x_g), M, C, _p(sums), _p(weights), _p(d_repr_l), _p(d_repr_g),
              _p(d_x_l), _p(d_x_g), _stream())


def som_smoothness(nodes, counts, side, M, loss, weight=None, gnodes=None):
    _lib.call("nvit_som_smoothness", _p(nodes), _p(counts), side, nodes.shape[1], M, _p(loss), _p(weight), _p(gnodes), _stream())


def adamw_flat(p, g, m, v, n_decay, lr, beta1, beta2, eps, weight_decay, step, gnorm_sq=None, max_norm=0.0, dev_lr_step=None):
    _lib.call("nvit_adamw_flat", _p(p), _p(g), _p(m), _p(v), p.numel(), n_decay, float(lr), float(beta1), float(beta2), float(eps),
              float(weight_decay), step, _p(gnorm_sq), float(max_norm), _p(dev_lr_step), _stream())


def weight_norm_multi(table, n_tensors, total_units):
    _lib.call("nvit_weight_norm_multi", _p(table), n_tensors, total_units, _stream())


def adamw_norm_fused(p, g, m, v, w16, table, n_segments, total_units, lr, beta1, beta2, eps, weight_decay, step, counter,
                     gnorm_sq=None, max_norm=0.0, dev_lr_step=None, zero_grad=True):
    """clip + AdamW + weight normalisation + bf16 operand emit + zero_grad in one pass (nvit_adamw_norm_fused)."""
    _lib.call("nvit_adamw_norm_fused", _p(p), _p(g), _p(m), _p(v), _p(w16), _p(table), n_segments, total_units, float(lr), float(beta1),
              float(beta2), float(eps), float(weight_decay), step, _p(gnorm_sq), float(max_norm), _p(dev_lr_step), _p(counter),
              int(zero_grad), _stream())


def head_scale_fwd(raw, sz, sz_mul, logits):
    B, N = raw.shape
    _lib.call("nvit_head_scale_fwd", _p(raw), _p(sz), float(sz_mul), _p(logits), B, N, _stream())
