"""nvit_b200 — B200-native (sm_100a) implementation of the nViT training hot path of slobodaapl/nvit.

Public surface mirrors the reference's ``nvit.model`` (ViTConfig, ViT, Block, CrossAttentionBlock, RMSNorm, justnorm; ``nvit.kohonen.KohonenMap``)
and the train-step contract of ``nvit/train.py`` (Trainer.step / normalize_matrices).  All arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``libnvit_b200.so`` (include/nvit_b200.h); there is no CPU fallback.
"""
from .model import ViTConfig, ViT, Block, CrossAttentionBlock, RMSNorm, justnorm  # noqa: F401
from .kohonen import KohonenMap  # noqa: F401
from .train import Trainer, GradReducer, DeviceLoader  # noqa: F401
from . import ops  # noqa: F401
from . import augment  # noqa: F401
from .augment import AutoAugment, get_transforms  # noqa: F401

__all__ = ["ViTConfig", "ViT", "Block", "CrossAttentionBlock", "RMSNorm", "justnorm", "KohonenMap", "Trainer", "GradReducer", "DeviceLoader", "ops", "augment", "AutoAugment", "get_transforms"]
