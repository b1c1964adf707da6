"""Micro-benchmarks of two HBM-bound kernels at the bench shapes, A/B of their variants on one box:

  * nvit_weight_norm_multi with the column-normalised matrices first in the unit table vs plain block order;
  * nvit_residual_bwd with rows held in registers (0) vs rows staged in shared memory by bulk copies (1).

    python scripts/hbm_kernels_bench.py

Times are CUDA events over 20 (10) launches on operands far larger than L2; GB/s are ALGORITHMIC bytes / time.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nvit_b200 import ViT, ViTConfig, ops, _lib
from oracle import nvit_oracle as O   # config table only

DEV = "cuda"


def timed(fn, n):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3      # us


def weight_norm():
    cfg = ViTConfig(**O.named_config("b16").as_dict())
    torch.manual_seed(0)
    model = ViT(cfg).to(DEV)
    eng = model.engine
    eng.param_list()
    nbytes = 16 * cfg.n_embd ** 2 * cfg.n_layer * 8
    for order in (True, False, True, False):
        eng._build_norm_table(columns_first=order)
        us = timed(eng.normalize_matrices, 20)
        print(f"weight_norm b16 columns_first={order}: {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s ({nbytes / 1e6:.0f} MB)", flush=True)
    eng._build_norm_table()


def residual_bwd(M, C):
    g = torch.Generator().manual_seed(0)
    mk = lambda dt=torch.float32: torch.randn(M, C, generator=g).to(DEV).to(dt)
    gr, h, h0, x = mk(), mk(), mk(), mk(torch.bfloat16)
    alpha = torch.full((C,), C ** -0.5, device=DEV)
    skip = torch.tensor([0.9], device=DEV)
    dh, dh0 = torch.zeros(M, C, device=DEV), torch.empty(M, C, device=DEV)
    dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    dalpha, dskip = torch.zeros(C, device=DEV), torch.zeros(1, device=DEV)
    for sk in (False, True):
        for acc in (False, True):
            nbytes = M * C * ((4 + 4 + 2) + (4 + 2) + (8 if sk else 0) + (4 if acc else 0))
            res = {}
            for mode in (0, 1, 0, 1):
                _lib.call("nvit_residual_bwd_staged", mode)
                us = timed(lambda: ops.residual_bwd(gr, h, x, alpha, 0.05 * C ** 0.5, dh, dx, dalpha, dh_accumulate=acc,
                                                    h0=h0 if sk else None, skip=skip if sk else None,
                                                    dh0=dh0 if sk else None, dskip=dskip if sk else None), 10)
                res.setdefault(mode, []).append(us)
            line = "  ".join(f"staged={m}: " + " / ".join(f"{u:.1f}" for u in v) + f" us ({nbytes / min(v) / 1e3:.0f} GB/s)" for m, v in res.items())
            print(f"residual_bwd M={M} C={C} skip={sk} acc={acc} ({nbytes / 1e6:.0f} MB): {line}", flush=True)
    _lib.call("nvit_residual_bwd_staged", 2)


if __name__ == "__main__":
    weight_norm()
    residual_bwd(50176, 768)
    residual_bwd(50176, 1024)
