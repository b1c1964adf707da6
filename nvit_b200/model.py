"""Drop-in mirror of the reference's ``nvit/model.py`` module API, executed by hand-written sm_100a kernels.

Same public names (``ViTConfig``, ``ViT``, ``Block``, ``CrossAttentionBlock``, ``RMSNorm``, ``justnorm``), same
constructor/forward signatures, same ``state_dict`` keys, shapes and initial distributions (modules are created in the
reference's order, so ``torch.manual_seed(s); ViT(cfg)`` draws the same random stream), same attributes the reference
trainer and debug tooling touch.  The modules here are *parameter containers*: the arithmetic of
``ViT.forward`` and its backward is run by :mod:`nvit_b200.engine` through the C ABI of ``libnvit_b200.so``.
There is no CPU or eager-PyTorch fallback: a CPU tensor or a missing library raises.

Reference: /root/reference/nvit/model.py (ViTConfig :13-40, justnorm :43-44, Block :47-169, RMSNorm :172-184,
CrossAttentionBlock :187-275, ViT :278-470).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import nn

from .kohonen import KohonenMap


@dataclass
class ViTConfig:
    """Field-for-field the reference's ViTConfig (nvit/model.py:13-40)."""
    image_size: int = 224
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 1024
    base_scale: float = 1.0 / (1024.0 ** 0.5)
    use_nvit: bool = False
    flash_attn: bool = False      # accepted; attention is always token attention
    sz_init_value: float = 1.00
    sz_init_scaling: float = 1.0
    dropout: float = 0.0          # never applied by the reference either
    bias: bool = False
    channels: int = 3
    num_classes: int = 1000
    local_patch_size: int = 8
    global_patch_size: int = 16
    kohonen_nodes: int = 512
    kohonen_alpha: float = 0.01
    use_kohonen: bool = False
    reconstruction_weight: float = 0.1
    map_balance_weight: float = 0.5
    kohonen_scheduler_enabled: bool = False
    kohonen_scheduler_warmup_steps: int = 1000
    kohonen_scheduler_decay_steps: int = 10000
    kohonen_scheduler_min_lr: float = 0.001
    local_quantization_weight: float = 0.1
    global_quantization_weight: float = 0.1


def justnorm(x: torch.Tensor) -> torch.Tensor:
    """x / ||x||_2 over the last dim, no epsilon (nvit/model.py:43-44).  API-surface utility, not on the hot path
    (inside the model the normalisations are fused into the residual and attention kernels)."""
    return x / x.norm(p=2, dim=-1, keepdim=True)


class RMSNorm(nn.Module):
    """Parameter container for the reference's RMSNorm (nvit/model.py:172-184)."""

    def __init__(self, embdim: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(embdim))
        self.eps = eps


class Block(nn.Module):
    """Parameters of one transformer block (nvit/model.py:47-82); arithmetic in engine.Engine._block_fwd/_block_bwd."""

    def __init__(self, config: ViTConfig) -> None:
        super().__init__()
        self.config = config
        C = config.n_embd
        self.key = nn.Linear(C, C, bias=config.bias)
        self.query = nn.Linear(C, C, bias=config.bias)
        self.value = nn.Linear(C, C, bias=config.bias)
        self.att_c_proj = nn.Linear(C, C, bias=config.bias)
        self.skip_param = nn.Parameter(torch.ones(1))
        self.c_fc = nn.Linear(C, 2 * 4 * C, bias=config.bias)
        self.silu = nn.SiLU()
        self.mlp_c_proj = nn.Linear(4 * C, C, bias=config.bias)
        # The reference creates these only when use_nvit (model.py:63-65) but uses them only when NOT use_nvit
        # (model.py:95-96, 145-146), so its original-ViT mode crashes.  Here they exist in both
        # modes: unused (and grad-less) under nViT exactly as in the reference, functional in the original-ViT mode.
        self.rmsnorm_att = RMSNorm(C)
        self.rmsnorm_mlp = RMSNorm(C)
        if config.use_nvit:
            f32 = torch.float32
            self.attn_alpha_init_value = torch.scalar_tensor(0.05, dtype=f32)
            self.attn_alpha_init_scaling = torch.scalar_tensor(config.base_scale, dtype=f32)
            self.attn_alpha = nn.Parameter(self.attn_alpha_init_scaling * torch.ones(C, dtype=f32))
            self.mlp_alpha_init_value = torch.scalar_tensor(0.05, dtype=f32)
            self.mlp_alpha_init_scaling = torch.scalar_tensor(config.base_scale, dtype=f32)
            self.mlp_alpha = nn.Parameter(self.mlp_alpha_init_scaling * torch.ones(C, dtype=f32))
            self.sqk_init_value = torch.scalar_tensor(1.0, dtype=f32)
            self.sqk_init_scaling = torch.scalar_tensor(config.base_scale, dtype=f32)
            self.sqk = nn.Parameter(self.sqk_init_scaling * torch.ones(C, dtype=f32))
            self.suv_init_value = torch.scalar_tensor(1.0, dtype=f32)
            self.suv_init_scaling = torch.scalar_tensor(1.0, dtype=f32)
            self.suv = nn.Parameter(self.suv_init_scaling * torch.ones(2 * 4 * C, dtype=f32))

    def justnorm(self, x: torch.Tensor) -> torch.Tensor:
        return justnorm(x)


class CrossAttentionBlock(nn.Module):
    """Parameters of the local/global cross-attention merge (nvit/model.py:187-217)."""

    def __init__(self, config: ViTConfig) -> None:
        super().__init__()
        self.config = config
        C = config.n_embd
        if not config.use_nvit:
            self.local_norm = RMSNorm(C)
            self.global_norm = RMSNorm(C)
        self.q_local = nn.Linear(C, C, bias=config.bias)
        self.k_global = nn.Linear(C, C, bias=config.bias)
        self.v_global = nn.Linear(C, C, bias=config.bias)
        self.proj = nn.Linear(C, 2 * C, bias=config.bias)
        self.silu = nn.SiLU()
        self.out_proj = nn.Linear(C, C, bias=config.bias)
        if config.use_nvit:
            f32 = torch.float32
            self.attn_alpha_init_value = torch.scalar_tensor(0.05, dtype=f32)
            self.attn_alpha_init_scaling = torch.scalar_tensor(config.base_scale, dtype=f32)
            self.attn_alpha = nn.Parameter(self.attn_alpha_init_scaling * torch.ones(C, dtype=f32))
            self.sqk_init_value = torch.scalar_tensor(1.0, dtype=f32)
            self.sqk_init_scaling = torch.scalar_tensor(config.base_scale, dtype=f32)
            self.sqk = nn.Parameter(self.sqk_init_scaling * torch.ones(C, dtype=f32))


class ViT(nn.Module):
    """The reference's ViT (nvit/model.py:278-470) with its forward/backward run by the sm_100a engine.

    ``forward(img[B, channels, S, S]) -> (logits[B, num_classes], aux_losses)`` with ``aux_losses["reconstruction"]``
    always present.  ``loss.backward()`` works through a single autograd node whose backward is the engine's
    hand-scheduled backward pass; :class:`nvit_b200.train.Trainer` drives the same engine without autograd.
    """

    def __init__(self, config: ViTConfig):
        super().__init__()
        self.config = config
        self.step = 0
        self.total_steps = 0
        if config.n_embd % config.n_head != 0 or config.n_embd // config.n_head != 64:
            raise ValueError("nvit_b200 attention kernels need head_dim = n_embd / n_head = 64")
        if (config.image_size // config.local_patch_size) ** 2 > 256:
            raise ValueError("nvit_b200 attention kernels hold one sequence per CTA: at most 256 tokens "
                             f"(image_size {config.image_size} / patch {config.local_patch_size} gives "
                             f"{(config.image_size // config.local_patch_size) ** 2})")
        if config.n_embd > 1024 or config.n_embd % 64 != 0:
            raise ValueError("nvit_b200 residual / norm kernels keep a token row in one warp's registers: n_embd must be a "
                             f"multiple of 64, at most 1024 (got {config.n_embd})")
        if config.use_kohonen and not config.use_nvit:
            raise NotImplementedError("Kohonen maps are built for the nViT branch (BASELINE config 5); use_nvit=False + use_kohonen is not")
        C, P, G = config.n_embd, config.local_patch_size, config.global_patch_size
        self.local_patch_embed = nn.Conv2d(config.channels, C, kernel_size=P, stride=P)
        self.global_patch_embed = nn.Sequential(
            nn.ReflectionPad2d((G - P) // 2),
            nn.Conv2d(config.channels, C, kernel_size=G, stride=P),
        )
        n_patches = (config.image_size // P) ** 2
        self.local_pos_embed = nn.Parameter(torch.zeros(1, n_patches, C))
        self.global_pos_embed = nn.Parameter(torch.zeros(1, n_patches, C))
        if config.use_kohonen:      # model.py:312-325
            a0 = config.kohonen_alpha if not config.kohonen_scheduler_enabled else config.kohonen_scheduler_min_lr
            self.local_kohonen = KohonenMap(C, config.kohonen_nodes // 2, a0)
            self.global_kohonen = KohonenMap(C, config.kohonen_nodes // 2, a0)
            self.map_balance = nn.Parameter(torch.tensor(config.map_balance_weight))
        self.cross_attention = CrossAttentionBlock(config)
        self.reconstruction_head = nn.Sequential(nn.Linear(C, P * P * config.channels), nn.Tanh())
        self.transformer = nn.ModuleDict({
            "drop": nn.Dropout(config.dropout),
            "h": nn.ModuleList([Block(config) for _ in range(config.n_layer)]),
        })
        self.mlp_head = nn.Sequential(nn.LayerNorm(C), nn.Linear(C, config.num_classes))
        if config.use_nvit:
            self.sz = nn.Parameter(config.sz_init_scaling * torch.ones(config.num_classes, dtype=torch.float32))
        self.apply(self._init_weights)
        for pn, p in self.named_parameters():
            if pn.endswith("c_proj.weight"):
                nn.init.normal_(p, mean=0.0, std=0.02 / math.sqrt(2 * config.n_layer))
        self._engine = None

    def _init_weights(self, module: nn.Module) -> None:
        if isinstance(module, nn.Linear):
            nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            nn.init.zeros_(module.bias)
            nn.init.ones_(module.weight)
        if self.config.use_nvit and isinstance(module, nn.Linear):
            nn.init.constant_(self.sz, self.config.sz_init_value)

    # ------------------------------------------------------------------ reference API surface
    @property
    def num_params(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def configure_optimizers(self, weight_decay: float, learning_rate: float, betas: tuple[float, float],
                             device_type: str) -> torch.optim.AdamW:
        """Same parameter groups as the reference (nvit/model.py:369-385).  nvit_b200.train.Trainer uses the flat
        fused AdamW kernel instead; this method keeps external training loops working."""
        param_dict = {pn: p for pn, p in self.named_parameters() if p.requires_grad}
        if self.config.use_nvit:
            groups = [
                {"params": [p for n, p in param_dict.items() if "sz" not in n and p.dim() >= 2], "weight_decay": weight_decay},
                {"params": [p for n, p in param_dict.items() if "sz" not in n and p.dim() < 2], "weight_decay": 0.0},
                {"params": [self.sz], "weight_decay": 0.0},
            ]
        else:
            groups = [
                {"params": [p for n, p in param_dict.items() if p.dim() >= 2], "weight_decay": weight_decay},
                {"params": [p for n, p in param_dict.items() if p.dim() < 2], "weight_decay": 0.0},
            ]
        return torch.optim.AdamW(groups, lr=learning_rate, betas=betas, fused=(device_type == "cuda"))

    def estimate_mfu(self, fwdbwd_per_iter: int, dt: float) -> tuple[float, float]:
        """Same estimate as nvit/model.py:387-401 (A100 312 TFLOPS constant kept for comparability)."""
        N = sum(p.numel() for p in self.parameters())
        cfg = self.config
        L, H, Q = cfg.n_layer, cfg.n_head, cfg.n_embd // cfg.n_head
        T = cfg.image_size // cfg.local_patch_size * cfg.image_size // cfg.local_patch_size
        flops_achieved = (6 * N + 12 * L * H * Q * T) * T * fwdbwd_per_iter / dt
        return flops_achieved / 312e12, flops_achieved

    def get_kohonen_lr(self, step: int) -> float:
        """Learning rate of the map update (nvit/model.py:563-581): constant, or linear warm-up then cosine decay."""
        cfg = self.config
        if not cfg.kohonen_scheduler_enabled:
            return cfg.kohonen_alpha
        w, d = cfg.kohonen_scheduler_warmup_steps, cfg.kohonen_scheduler_decay_steps
        lo, hi = cfg.kohonen_scheduler_min_lr, cfg.kohonen_alpha
        if step < w:
            return lo + (hi - lo) * (step / w)
        if step > d:
            return lo
        return lo + 0.5 * (1.0 + math.cos(math.pi * (step - w) / (d - w))) * (hi - lo)

    def combine_representations(self, local_repr: torch.Tensor, global_repr: torch.Tensor) -> torch.Tensor:
        """nvit/model.py:477-480 (API surface used by the reference's debug tooling; not on the training path)."""
        combined = local_repr * global_repr
        return combined / combined.norm(p=2, dim=-1, keepdim=True)

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    def _apply(self, fn, *args, **kwargs):
        # model.to(dev) / .cuda() / .float() re-create parameter storage: the engine's flat buffers must be rebuilt then - but
        # only then (a no-op .to() must not throw the flat buffers, and the Trainer's optimizer state keyed on them, away)
        before = [(p.data_ptr(), p.device, p.dtype) for p in self.parameters()]
        out = super()._apply(fn, *args, **kwargs)
        after = [(p.data_ptr(), p.device, p.dtype) for p in self.parameters()]
        if getattr(self, "_engine", None) is not None and before != after:
            self._engine.invalidate()
        return out

    def forward(self, img: torch.Tensor) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        if self.training:
            self.step += 1
        if not img.is_cuda:
            raise RuntimeError("nvit_b200.ViT runs on a CUDA device only (sm_100a kernels, no CPU fallback)")
        from .engine import NViTFunction
        eng = self.engine
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            out = NViTFunction.apply(eng, img, *eng.param_list())
        else:
            logits, recon = eng.forward(img, save=False)
            out = (logits, recon, *(eng.last_aux[k] for k in AUX_KEYS)) if self.config.use_kohonen else (logits, recon)
        aux = {}
        if self.config.use_kohonen:      # same keys, same order as nvit/model.py:437-442, 464
            aux.update(zip(AUX_KEYS, out[2:]))
        aux["reconstruction"] = out[1]
        return out[0], aux


AUX_KEYS = ("kohonen_consistency", "kohonen_smoothness", "local_quantization", "global_quantization")
