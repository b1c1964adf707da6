"""Writes tests/golden/augment_golden.npz: outputs of torchvision 0.26's uint8 AutoAugment operations
(torchvision.transforms.autoaugment._apply_op, nearest interpolation, zero fill) on three seeded 24 x 24 images, for every
operation at several signed magnitudes.  Run in the build container:  python tests/golden/make_augment_golden.py"""
import os
import sys

import numpy as np
import torch
from torchvision.transforms import InterpolationMode
from torchvision.transforms.autoaugment import _apply_op

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from nvit_b200 import augment as A  # noqa: E402  (magnitude bins only)
sys.path.insert(0, os.path.dirname(HERE))
from test_augment_cpu import ALL_OPS, images  # noqa: E402

S = 24
X = images(S, 2, seed=2024)
names, mags, idxs, outs = [], [], [], []
for op in ALL_OPS:
    bins = [None] if op in ("AutoContrast", "Equalize", "Invert") else [2, 6, 9]
    for b in bins:
        for sign in ((1, -1) if op in A._SIGNED else (1,)):
            mag = sign * A.magnitude(op, b, S) if b is not None else 0.0
            for i in range(len(X)):
                x = torch.from_numpy(X[i]).permute(2, 0, 1).contiguous()
                y = _apply_op(x, op, mag, interpolation=InterpolationMode.NEAREST, fill=None).permute(1, 2, 0).contiguous().numpy()
                names.append(op); mags.append(mag); idxs.append(i); outs.append(y)
np.savez_compressed(os.path.join(HERE, "augment_golden.npz"), images=X, op_names=np.array(names), magnitudes=np.array(mags, np.float64),
                    image_index=np.array(idxs, np.int64), outputs=np.stack(outs))
print(len(outs), "cases")
