"""CPU oracle of nvit_augment_u8 (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this; the product path never does).

Restates, in numpy on uint8 HWC images, the eleven operations the kernel applies (nvit_b200/csrc/augment.cu), taking the
same encoded (code, 8 float32 parameters) the host sampler emits.  The reference applies kornia's AutoAugment in its
DataLoader workers (/root/reference/nvit/train.py:1081-1092, 262-273); kornia is a third-party dependency that is absent
from this image (the reference's pyproject pins none of its arithmetic), so the pixel arithmetic follows the published
uint8 definitions of the same AutoAugment operations as torchvision 0.26 implements them
(torchvision/transforms/_functional_tensor.py: _blend, rgb_to_grayscale, adjust_*, posterize, solarize, autocontrast,
equalize, _blurred_degenerate_image, affine / rotate with nearest interpolation).

Pinned in tests/test_augment_cpu.py against torchvision run live in this image AND against the committed fixtures
tests/golden/augment_golden.npz (made by tests/golden/make_augment_golden.py from torchvision): bit-exact for the ten
pixel operations; the affine operations agree except where a source coordinate falls within float rounding of a pixel
boundary (the test states the rate).
"""
from __future__ import annotations

import numpy as np

IDENTITY, AFFINE, BRIGHTNESS, COLOR, CONTRAST, SHARPNESS, POSTERIZE, SOLARIZE, AUTOCONTRAST, EQUALIZE, INVERT = range(11)
f32 = np.float32


def _blend(a: np.ndarray, b, r: np.float32, r1: np.float32) -> np.ndarray:
    """trunc(clamp(r * a + r1 * b, 0, 255)), every product and the sum rounded to float32 once (_functional_tensor._blend)."""
    v = (f32(r) * a.astype(f32)).astype(f32) + (f32(r1) * np.asarray(b, dtype=f32)).astype(f32)
    return np.clip(v.astype(f32), f32(0), f32(255)).astype(np.uint8)


def _gray(img: np.ndarray) -> np.ndarray:
    """trunc(0.2989 r + 0.587 g + 0.114 b), left to right in float32 (rgb_to_grayscale)."""
    x = img.astype(f32)
    v = ((f32(0.2989) * x[..., 0]).astype(f32) + (f32(0.587) * x[..., 1]).astype(f32)).astype(f32)
    v = (v + (f32(0.114) * x[..., 2]).astype(f32)).astype(f32)
    return v.astype(np.uint8)


def _affine(img: np.ndarray, p: np.ndarray) -> np.ndarray:
    S = img.shape[0]
    m00, m01, ox, m10, m11, oy = (f32(v) for v in p[:6])
    c = f32(0.5) * f32(S - 1)
    ys, xs = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    dx, dy = xs.astype(f32) - c, ys.astype(f32) - c
    sx = (((m00 * dx).astype(f32) + (m01 * dy).astype(f32)).astype(f32) + ox).astype(f32)
    sy = (((m10 * dx).astype(f32) + (m11 * dy).astype(f32)).astype(f32) + oy).astype(f32)
    ix = np.rint(sx.astype(np.float64)).astype(np.int64)      # ties to even, like cvt.rni
    iy = np.rint(sy.astype(np.float64)).astype(np.int64)
    ok = (ix >= 0) & (ix < S) & (iy >= 0) & (iy < S) & np.isfinite(sx) & np.isfinite(sy)
    out = np.zeros_like(img)
    out[ok] = img[iy[ok], ix[ok]]
    return out


def _sharp_degenerate(img: np.ndarray) -> np.ndarray:
    S = img.shape[0]
    x = img.astype(np.int64)
    out = img.copy()
    if S <= 2:
        return out
    acc = 5 * x[1:-1, 1:-1]
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy or dx:
                acc = acc + x[1 + dy:S - 1 + dy, 1 + dx:S - 1 + dx]
    out[1:-1, 1:-1] = np.rint((acc.astype(f32) / f32(13)).astype(f32)).astype(np.uint8)
    return out


def _autocontrast(img: np.ndarray) -> np.ndarray:
    out = np.empty_like(img)
    for c in range(3):
        ch = img[..., c]
        mn, mx = int(ch.min()), int(ch.max())
        # torch evaluates `255.0 / t` as t.reciprocal() * 255: two float32 roundings
        scale, lo = (f32(f32(1) / f32(mx - mn)) * f32(255), f32(mn)) if mx > mn else (f32(1), f32(0))
        v = ((ch.astype(f32) - lo).astype(f32) * scale).astype(f32)
        out[..., c] = np.clip(v, f32(0), f32(255)).astype(np.uint8)
    return out


def _equalize(img: np.ndarray) -> np.ndarray:
    out = np.empty_like(img)
    for c in range(3):
        ch = img[..., c]
        hist = np.bincount(ch.reshape(-1), minlength=256).astype(np.int64)
        nz = hist[hist != 0]
        step = int(nz[:-1].sum()) // 255
        if step == 0:
            out[..., c] = ch
            continue
        lut = (np.cumsum(hist) + step // 2) // step
        lut = np.clip(np.concatenate([[0], lut[:-1]]), 0, 255).astype(np.uint8)
        out[..., c] = lut[ch]
    return out


def apply_op(img: np.ndarray, code: int, p: np.ndarray) -> np.ndarray:
    """One encoded operation on one uint8 [S, S, 3] image."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[0] == img.shape[1] and img.shape[2] == 3
    p = np.asarray(p, dtype=f32)
    if code == AFFINE:
        return _affine(img, p)
    if code == BRIGHTNESS:
        return _blend(img, f32(0), p[0], p[1])
    if code == COLOR:
        return _blend(img, _gray(img)[..., None], p[0], p[1])
    if code == CONTRAST:
        total = int(_gray(img).astype(np.int64).sum())
        mean = f32(f32(total) / f32(img.shape[0] * img.shape[1]))
        return _blend(img, mean, p[0], p[1])
    if code == SHARPNESS:
        if img.shape[0] <= 2:
            return img.copy()
        return _blend(img, _sharp_degenerate(img), p[0], p[1])
    if code == POSTERIZE:
        return img & np.uint8(int(p[0]))
    if code == SOLARIZE:
        return np.where(img.astype(f32) >= p[0], 255 - img, img).astype(np.uint8)
    if code == AUTOCONTRAST:
        return _autocontrast(img)
    if code == EQUALIZE:
        return _equalize(img)
    if code == INVERT:
        return (255 - img).astype(np.uint8)
    return img.copy()          # identity and unknown codes


def apply_plan(batch: np.ndarray, ops: np.ndarray, params: np.ndarray) -> np.ndarray:
    """The whole entry point: batch uint8 [B, S, S, 3], ops int32 [B, 2], params float32 [B, 2, 8]."""
    out = np.empty_like(batch)
    for b in range(batch.shape[0]):
        x = apply_op(batch[b], int(ops[b, 0]), params[b, 0])
        out[b] = apply_op(x, int(ops[b, 1]), params[b, 1])
    return out
