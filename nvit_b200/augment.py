"""Device-side AutoAugment for raw uint8 batches — the host half.

The reference augments per sample in its DataLoader workers: ``kornia.augmentation.auto.AutoAugment(dataset)`` behind
``Normalize(0.5, 0.5)`` (``get_transforms``, train.py:1081-1092, used at train.py:262-273).  Here a batch crosses PCIe as
uint8 ``[B, S, S, 3]``; this module draws, per image, one sub-policy of the dataset's AutoAugment policy (Cubuk et al.,
"AutoAugment: Learning Augmentation Strategies from Data", 2019 — 25 sub-policies of two operations, each with a
probability and one of ten magnitude bins; order and magnitude space as torchvision 0.26 tabulates them, which
tests/test_augment_cpu.py checks) and encodes the two operations as ``(code, 8 floats)`` for ``nvit_augment_u8``
(nvit_b200/csrc/augment.cu).  ``Normalize(0.5, 0.5)`` stays folded into the patch gather (``nvit_im2col_u8``).

    train_tf, val_tf = get_transforms("cifar10", seed=0)        # same call shape as the reference's Trainer.get_transforms
    for X_u8, y in DeviceLoader(batches, device, transform=train_tf): ...

The arithmetic lives in the kernel; nothing here touches pixels and there is no CPU path for them.
"""
from __future__ import annotations

import math

import numpy as np

# operation codes of nvit_augment_u8 (include/nvit_b200.h)
IDENTITY, AFFINE, BRIGHTNESS, COLOR, CONTRAST, SHARPNESS, POSTERIZE, SOLARIZE, AUTOCONTRAST, EQUALIZE, INVERT = range(11)
NPARAM = 8
NUM_BINS = 10

# "op:probability:magnitude bin" pairs; "-" = the operation takes no magnitude
_POLICY_TEXT = {
    "imagenet": """Posterize:.4:8 Rotate:.6:9|Solarize:.6:5 AutoContrast:.6:-|Equalize:.8:- Equalize:.6:-|Posterize:.6:7 Posterize:.6:6|
        Equalize:.4:- Solarize:.2:4|Equalize:.4:- Rotate:.8:8|Solarize:.6:3 Equalize:.6:-|Posterize:.8:5 Equalize:1:-|
        Rotate:.2:3 Solarize:.6:8|Equalize:.6:- Posterize:.4:6|Rotate:.8:8 Color:.4:0|Rotate:.4:9 Equalize:.6:-|
        Equalize:0:- Equalize:.8:-|Invert:.6:- Equalize:1:-|Color:.6:4 Contrast:1:8|Rotate:.8:8 Color:1:2|
        Color:.8:8 Solarize:.8:7|Sharpness:.4:7 Invert:.6:-|ShearX:.6:5 Equalize:1:-|Color:.4:0 Equalize:.6:-|
        Equalize:.4:- Solarize:.2:4|Solarize:.6:5 AutoContrast:.6:-|Invert:.6:- Equalize:1:-|Color:.6:4 Contrast:1:8|
        Equalize:.8:- Equalize:.6:-""",
    "cifar10": """Invert:.1:- Contrast:.2:6|Rotate:.7:2 TranslateX:.3:9|Sharpness:.8:1 Sharpness:.9:3|ShearY:.5:8 TranslateY:.7:9|
        AutoContrast:.5:- Equalize:.9:-|ShearY:.2:7 Posterize:.3:7|Color:.4:3 Brightness:.6:7|Sharpness:.3:9 Brightness:.7:9|
        Equalize:.6:- Equalize:.5:-|Contrast:.6:7 Sharpness:.6:5|Color:.7:7 TranslateX:.5:8|Equalize:.3:- AutoContrast:.4:-|
        TranslateY:.4:3 Sharpness:.2:6|Brightness:.9:6 Color:.2:8|Solarize:.5:2 Invert:0:-|Equalize:.2:- AutoContrast:.6:-|
        Equalize:.2:- Equalize:.6:-|Color:.9:9 Equalize:.6:-|AutoContrast:.8:- Solarize:.2:8|Brightness:.1:3 Color:.7:0|
        Solarize:.4:5 AutoContrast:.9:-|TranslateY:.9:9 TranslateY:.7:9|AutoContrast:.9:- Solarize:.8:3|Equalize:.8:- Invert:.1:-|
        TranslateY:.7:9 AutoContrast:.9:-""",
    "svhn": """ShearX:.9:4 Invert:.2:-|ShearY:.9:8 Invert:.7:-|Equalize:.6:- Solarize:.6:6|Invert:.9:- Equalize:.6:-|
        Equalize:.6:- Rotate:.9:3|ShearX:.9:4 AutoContrast:.8:-|ShearY:.9:8 Invert:.4:-|ShearY:.9:5 Solarize:.2:6|
        Invert:.9:- AutoContrast:.8:-|Equalize:.6:- Rotate:.9:3|ShearX:.9:4 Solarize:.3:3|ShearY:.8:8 Invert:.7:-|
        Equalize:.9:- TranslateY:.6:6|Invert:.9:- Equalize:.6:-|Contrast:.3:3 Rotate:.8:4|Invert:.8:- TranslateY:0:2|
        ShearY:.7:6 Solarize:.4:8|Invert:.6:- Rotate:.8:4|ShearY:.3:7 TranslateX:.9:3|ShearX:.1:6 Invert:.6:-|
        Solarize:.7:2 TranslateY:.6:7|ShearY:.8:4 Invert:.8:-|ShearX:.7:9 TranslateY:.8:3|ShearY:.8:5 AutoContrast:.7:-|
        ShearX:.7:2 Invert:.1:-""",
}
_SIGNED = {"ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "Brightness", "Color", "Contrast", "Sharpness"}
# kornia's AutoAugment accepts the same three policy names; the reference passes settings.data.dataset (train.py:1083)
_ALIASES = {"cifar": "cifar10", "cifar-10": "cifar10", "cifar100": "cifar10", "imagenet1k": "imagenet", "imagenet-1k": "imagenet"}


def policies(dataset: str) -> list[tuple[tuple[str, float, int | None], tuple[str, float, int | None]]]:
    """The 25 sub-policies of a dataset's AutoAugment policy as ((op, p, bin), (op, p, bin)) pairs."""
    key = _ALIASES.get(dataset.lower(), dataset.lower())
    if key not in _POLICY_TEXT:
        raise ValueError(f"no AutoAugment policy for dataset {dataset!r} (have {sorted(_POLICY_TEXT)})")
    out = []
    for sub in _POLICY_TEXT[key].replace("\n", "").split("|"):
        pair = []
        for item in sub.split():
            name, p, mag = item.split(":")
            pair.append((name, float(p), None if mag == "-" else int(mag)))
        out.append((pair[0], pair[1]))
    return out


def _linspace32(start: float, end: float, steps: int = NUM_BINS) -> list[float]:
    """float32 linspace as torch.linspace evaluates it (float32 step; start + i step up to the middle, end - (n-1-i) step
    beyond it, each as ONE fused multiply-add), so the magnitudes are the very float32 values torchvision hands to its kernels."""
    start, end = np.float32(start), np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    half = steps // 2
    # float64 holds the float32 product exactly; rounding the float64 sum once to float32 is the fused result
    return [float(np.float32(np.float64(start) + np.float64(step) * i)) if i < half
            else float(np.float32(np.float64(end) - np.float64(step) * (steps - 1 - i))) for i in range(steps)]


def magnitude(op: str, bin_id: int | None, image_size: int) -> float:
    """Magnitude of bin `bin_id` (0..9) of operation `op` on a square image of side `image_size`."""
    if bin_id is None:
        return 0.0
    if op in ("ShearX", "ShearY"):
        return _linspace32(0.0, 0.3)[bin_id]
    if op in ("TranslateX", "TranslateY"):
        return _linspace32(0.0, 150.0 / 331.0 * image_size)[bin_id]
    if op == "Rotate":
        return _linspace32(0.0, 30.0)[bin_id]
    if op in ("Brightness", "Color", "Contrast", "Sharpness"):
        return _linspace32(0.0, 0.9)[bin_id]
    if op == "Posterize":
        return float(8 - int(np.round(np.float32(bin_id) / np.float32((NUM_BINS - 1) / 4))))
    if op == "Solarize":
        return _linspace32(255.0, 0.0)[bin_id]
    raise ValueError(f"operation {op!r} takes no magnitude")


def inverse_affine(S: int, angle_deg: float = 0.0, translate=(0.0, 0.0), shear_deg=(0.0, 0.0), center=(0.0, 0.0)) -> list[float]:
    """Output pixel -> source pixel map of `translate . (rotate/shear about center)`, coordinates relative to the image
    centre with y pointing down and positive angles turning clockwise on screen, as [m00, m01, ox, m10, m11, oy]:
    source = M (x - c, y - c) + o with c = (S - 1) / 2 already added into o."""
    rot, sx, sy = math.radians(angle_deg), math.radians(shear_deg[0]), math.radians(shear_deg[1])
    fwd = np.eye(3)
    fwd[0, 0] = math.cos(rot - sy) / math.cos(sy)
    fwd[0, 1] = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    fwd[1, 0] = math.sin(rot - sy) / math.cos(sy)
    fwd[1, 1] = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)

    def shift(tx, ty):
        m = np.eye(3)
        m[0, 2], m[1, 2] = tx, ty
        return m

    full = shift(translate[0], translate[1]) @ shift(center[0], center[1]) @ fwd @ shift(-center[0], -center[1])
    inv = np.linalg.inv(full)
    c = 0.5 * (S - 1)
    return [inv[0, 0], inv[0, 1], inv[0, 2] + c, inv[1, 0], inv[1, 1], inv[1, 2] + c]


def encode_op(op: str, mag: float, S: int) -> tuple[int, list[float]]:
    """(code, parameters) of one AutoAugment operation at signed magnitude `mag` for nvit_augment_u8."""
    p = [0.0] * NPARAM
    if op == "Identity":
        return IDENTITY, p
    if op == "ShearX":      # about the top-left corner, by atan(mag) (the policy's level is the matrix entry itself)
        p[:6] = inverse_affine(S, shear_deg=(math.degrees(math.atan(mag)), 0.0), center=(-0.5 * S, -0.5 * S))
        return AFFINE, p
    if op == "ShearY":
        p[:6] = inverse_affine(S, shear_deg=(0.0, math.degrees(math.atan(mag))), center=(-0.5 * S, -0.5 * S))
        return AFFINE, p
    if op == "TranslateX":
        p[:6] = inverse_affine(S, translate=(float(int(mag)), 0.0))
        return AFFINE, p
    if op == "TranslateY":
        p[:6] = inverse_affine(S, translate=(0.0, float(int(mag))))
        return AFFINE, p
    if op == "Rotate":      # positive magnitudes turn counter-clockwise
        p[:6] = inverse_affine(S, angle_deg=-mag)
        return AFFINE, p
    if op in ("Brightness", "Color", "Contrast", "Sharpness"):
        ratio = 1.0 + mag
        p[0], p[1] = ratio, 1.0 - ratio          # both rounded to float32 on their own, as a float32 kernel sees two scalars
        return {"Brightness": BRIGHTNESS, "Color": COLOR, "Contrast": CONTRAST, "Sharpness": SHARPNESS}[op], p
    if op == "Posterize":
        bits = int(mag)
        if not 0 <= bits <= 8:
            raise ValueError(f"Posterize keeps 0..8 bits, got {bits}")
        p[0] = float(256 - (1 << (8 - bits)))     # the byte mask
        return POSTERIZE, p
    if op == "Solarize":
        p[0] = mag
        return SOLARIZE, p
    if op == "AutoContrast":
        return AUTOCONTRAST, p
    if op == "Equalize":
        return EQUALIZE, p
    if op == "Invert":
        return INVERT, p
    raise ValueError(f"unknown AutoAugment operation {op!r}")


class AutoAugment:
    """Per-image AutoAugment on a uint8 HWC device batch: ``aug(X_u8) -> X_u8`` (a new tensor).

    Randomness: one ``numpy`` PCG64 stream seeded with ``seed`` (+ the data-parallel rank, so replicas augment
    differently); per image a sub-policy index, two uniforms against the operations' probabilities and two sign bits."""

    def __init__(self, dataset: str = "imagenet", seed: int = 0, rank: int = 0):
        self.dataset = dataset
        self.policies = policies(dataset)
        self.rng = np.random.Generator(np.random.PCG64([int(seed), int(rank)]))
        self._table_cache: dict = {}
        self._rings: dict = {}
        self.last_plan = None

    def _encoded(self, S: int):
        """codes int32 [25, 2, 2] and params float32 [25, 2, 2, 8] indexed (sub-policy, slot, sign bit) for image side S."""
        t = self._table_cache.get(S)
        if t is None:
            n = len(self.policies)
            codes = np.zeros((n, 2, 2), np.int32)
            params = np.zeros((n, 2, 2, NPARAM), np.float32)
            for i, sub in enumerate(self.policies):
                for j, (name, _, bin_id) in enumerate(sub):
                    m = magnitude(name, bin_id, S) if bin_id is not None else 0.0
                    for sign in (0, 1):
                        codes[i, j, sign], params[i, j, sign] = encode_op(name, -m if (name in _SIGNED and sign == 0) else m, S)
            t = (codes, params, np.array([[sub[0][1], sub[1][1]] for sub in self.policies], np.float64))
            self._table_cache[S] = t
        return t

    def plan(self, B: int, S: int) -> tuple[np.ndarray, np.ndarray]:
        """Draw the operations of one batch: ops int32 [B, 2], params float32 [B, 2, 8] (identity where an operation's
        probability did not fire)."""
        codes, table, prob = self._encoded(S)
        sub = self.rng.integers(0, len(self.policies), size=B)
        probs = self.rng.random((B, 2))
        signs = self.rng.integers(0, 2, size=(B, 2))
        fire = probs <= prob[sub]                                   # [B, 2]
        slot = np.arange(2)[None, :]
        ops = np.where(fire, codes[sub[:, None], slot, signs], 0).astype(np.int32)
        params = np.where(fire[..., None], table[sub[:, None], slot, signs], np.float32(0)).astype(np.float32)
        return np.ascontiguousarray(ops), np.ascontiguousarray(params)

    def _staging(self, B: int, device):
        """A small ring of pinned host buffers for the plans, so that the copy to the device is asynchronous; a slot is reused
        only after the copy out of it has completed."""
        import torch
        key = (B, str(device))
        ring = self._rings.get(key)
        if ring is None:
            ring = {"i": 0, "slots": [{"ops": torch.empty(B, 2, dtype=torch.int32).pin_memory(),
                                       "params": torch.empty(B, 2, NPARAM, dtype=torch.float32).pin_memory(),
                                       "done": None} for _ in range(4)]}
            self._rings[key] = ring
        slot = ring["slots"][ring["i"] % len(ring["slots"])]
        ring["i"] += 1
        if slot["done"] is not None:
            slot["done"].synchronize()
        return slot

    def __call__(self, x_u8, out=None):
        """out: optional uint8 tensor of the same shape to write into (e.g. Trainer.input_buffers, which a captured step reads)."""
        import torch
        from . import ops as _ops
        if x_u8.dim() != 4 or x_u8.dtype != torch.uint8 or x_u8.shape[1] != x_u8.shape[2] or x_u8.shape[3] != 3:
            raise ValueError(f"AutoAugment expects a uint8 [B, S, S, 3] (HWC) batch, got {tuple(x_u8.shape)} {x_u8.dtype}")
        if not x_u8.is_cuda:
            raise RuntimeError("nvit_b200.AutoAugment runs on the device only (there is no CPU path): move the batch to the GPU, "
                               "e.g. through DeviceLoader(..., transform=aug)")
        B, S = int(x_u8.shape[0]), int(x_u8.shape[1])
        ops_h, params_h = self.plan(B, S)
        self.last_plan = (ops_h, params_h)
        slot = self._staging(B, x_u8.device)
        slot["ops"].copy_(torch.from_numpy(ops_h))
        slot["params"].copy_(torch.from_numpy(params_h))
        ops_d = slot["ops"].to(x_u8.device, non_blocking=True)
        params_d = slot["params"].to(x_u8.device, non_blocking=True)
        slot["done"] = torch.cuda.Event()
        slot["done"].record(torch.cuda.current_stream(x_u8.device))
        if out is None:
            out = torch.empty_like(x_u8)
        _ops.augment_u8(x_u8, out, ops_d, params_d)
        return out


class Identity:
    """The validation transform: Normalize(0.5, 0.5) only (train.py:1088-1090), which nvit_im2col_u8 already applies."""

    def __call__(self, x_u8):
        return x_u8


def get_transforms(dataset: str, seed: int = 0, rank: int = 0):
    """(train_transform, val_transform) for uint8 device batches — the shape of Trainer.get_transforms (train.py:1081-1092)."""
    return AutoAugment(dataset, seed=seed, rank=rank), Identity()
