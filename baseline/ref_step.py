"""Drive the UNMODIFIED reference model (baseline/_ref/nvit/model.py) through the reference's own step contract.

`nvit/train.py` itself cannot be imported (kornia / dynaconf / wandb are absent, Trainer.__init__ needs CUDA + wandb
online), so the inner step of Trainer.train is restated here around the real model object - nothing of this repo's
engine, kernels or oracle is on this path:

    train.py:898-946   autocast forward -> cross-entropy -> (scaled) backward -> unscale_ + clip_grad_norm_(1.0)
                       -> optimizer.step (ViT.configure_optimizers' AdamW) -> zero_grad(set_to_none=True)
    train.py:461-480   normalize_matrices (only with use_nvit)
    train.py:135-136   GradScaler() whenever AMP with float16 / bfloat16 is on (the shipped default: bfloat16)

Used by `bench.py --impl reference` (CPU, fp32: what the reference runs on a CPU device, where its autocast context is a
nullcontext, train.py:254) and by scripts/ref_gpu_baseline.py (one B200, bf16 autocast + GradScaler, eager and
torch.compile: the secondary baselines of SURVEY.md 8d).
"""
from __future__ import annotations

import contextlib
import os
import time

import torch
import torch.nn.functional as F

from .stage_ref import import_reference


def build_reference_model(cfg_dict: dict, device, seed: int = 0):
    ref = import_reference()
    torch.manual_seed(seed)
    cfg = ref.ViTConfig(**cfg_dict)
    model = ref.ViT(cfg).to(device)
    model.train()
    return model


def normalize_matrices(model) -> None:
    """Trainer.normalize_matrices, train.py:461-480 (justnorm in fp32 over dim 1 / dim 0, written through .data.copy_)."""
    if not model.config.use_nvit:
        return

    def justnorm(x, idim):
        dtype = x.dtype
        x = x.float()
        return (x / x.norm(p=2, dim=idim, keepdim=True)).to(dtype=dtype)

    for block in model.transformer.h:
        for name, dim in (("query", 1), ("key", 1), ("value", 1), ("att_c_proj", 0), ("c_fc", 1), ("mlp_c_proj", 0)):
            w = getattr(block, name).weight
            w.data.copy_(justnorm(w.data, dim))


class ReferenceStepper:
    """One rank of the reference loop on `device` ("cpu" or a cuda device)."""

    def __init__(self, cfg_dict: dict, device="cpu", lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0, seed=0,
                 amp_dtype=None, compile_model: bool = False):
        self.device = torch.device(device)
        self.model = build_reference_model(cfg_dict, self.device, seed)
        self.fwd = torch.compile(self.model) if compile_model else self.model     # train.py:422 `torch.compile(self.model)`
        self.optimizer = self.model.configure_optimizers(weight_decay, lr, betas, self.device.type)
        self.grad_clip = grad_clip
        cuda = self.device.type == "cuda"
        self.ctx = (torch.autocast(device_type="cuda", dtype=amp_dtype) if (cuda and amp_dtype is not None)
                    else contextlib.nullcontext())
        self.scaler = torch.amp.GradScaler("cuda") if (cuda and amp_dtype in (torch.float16, torch.bfloat16)) else None

    def step(self, X, y):
        with self.ctx:
            logits, aux = self.fwd(X)
            loss = F.cross_entropy(logits, y)
        if self.scaler is not None:
            self.scaler.scale(loss).backward()
        else:
            loss.backward()
        if self.grad_clip:
            if self.scaler is not None:
                self.scaler.unscale_(self.optimizer)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip)
        if self.scaler is not None:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        normalize_matrices(self.model)
        return loss.detach()


def synthetic_batch(cfg_dict: dict, batch: int, rank: int = 0):
    """per rank r, Generator().manual_seed(1234 + r); X = randn, y = randint."""
    g = torch.Generator().manual_seed(1234 + rank)
    S, ch = cfg_dict["image_size"], cfg_dict.get("channels", 3)
    X = torch.randn(batch, ch, S, S, generator=g)
    y = torch.randint(0, cfg_dict["num_classes"], (batch,), generator=g)
    return X, y


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_cpu_steps(cfg_dict: dict, batch: int, steps: int, warmup: int, threads: int | None = None):
    """Seconds per step of the reference model + restated step on the host cores (fp32, eager)."""
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    st = ReferenceStepper(cfg_dict, "cpu")
    X, y = synthetic_batch(cfg_dict, batch)
    for _ in range(warmup):
        st.step(X, y)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st.step(X, y)
        times.append(time.perf_counter() - t0)
    return times, threads
