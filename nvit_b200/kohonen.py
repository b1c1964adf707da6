"""Drop-in mirror of the reference's ``nvit/kohonen.py`` module API (KohonenMap), on the sm_100a kernels.

Same constructor, parameters (``nodes``), buffers (``locations``, ``offsets``) and attributes (``m``, ``n``, ``grid_size``,
``input_dim``, ``alpha``, ``sigma``, ``periodic``) as /root/reference/nvit/kohonen.py:31-80, so ``state_dict`` interchanges.
Inside ``ViT.forward`` the maps are driven by :mod:`nvit_b200.engine`; ``forward`` / ``update_nodes`` here serve callers
that use a map on its own (the reference's debug tooling does, debug.py:129-138) and run the same kernels.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

F32 = torch.float32
BF16 = torch.bfloat16


class KohonenMap(nn.Module):
    def __init__(self, input_dim: int, num_nodes: int, alpha: float = 0.01, sigma: float | None = None,
                 periodic: bool = True) -> None:
        super().__init__()
        self.m = int(num_nodes ** 0.5)
        self.n = num_nodes // self.m
        self.grid_size = self.m * self.n
        self.input_dim = input_dim
        self.alpha = alpha
        self.periodic = periodic
        self.nodes = nn.Parameter(torch.randn(self.grid_size, input_dim))
        rows = torch.arange(self.m).repeat_interleave(self.n)
        cols = torch.arange(self.n).repeat(self.m)
        self.register_buffer("locations", torch.stack((rows, cols), dim=1).to(torch.long))
        self.sigma = (self.m * self.n) ** 0.5 / 2.0 if sigma is None else float(sigma)
        if periodic:
            m, n = self.m, self.n
            self.register_buffer("offsets", torch.tensor([[-m, -n], [m, n], [-m, 0], [m, 0], [0, -n], [0, n], [-m, n], [m, -n]]))

    def _require_cuda(self, x):
        if not x.is_cuda:
            raise RuntimeError("nvit_b200.KohonenMap runs on a CUDA device only (sm_100a kernels, no CPU fallback)")
        if not self.periodic:
            raise NotImplementedError("only the periodic (torus) topology the reference's ViT uses is built")

    @torch.no_grad()
    def _bmu(self, x2d: torch.Tensor):
        M, C = x2d.shape
        G, dev = self.grid_size, x2d.device
        nodes = self.nodes.data.float().contiguous()
        node_sq = torch.empty(G, device=dev, dtype=F32)
        hi, lo = torch.empty(G, C, device=dev, dtype=BF16), torch.empty(G, C, device=dev, dtype=BF16)
        snap = torch.empty(G, C, device=dev, dtype=F32)
        ops.som_prepare(nodes, node_sq, hi, lo, snap)
        xhi, xlo = torch.empty(M, C, device=dev, dtype=BF16), torch.empty(M, C, device=dev, dtype=BF16)
        ops.split_bf16(x2d, xhi, xlo)
        dots = torch.empty(M, G, device=dev, dtype=F32)
        for k, (xa, nb) in enumerate(((xhi, hi), (xhi, lo), (xlo, hi))):
            ops.gemm(xa, nb, dots, M=M, N=G, K=C, lda=C, ldb=C, ldc=G, accumulate=(k > 0))
        idx = torch.empty(M, device=dev, dtype=torch.int32)
        idx64 = torch.empty(M, device=dev, dtype=torch.int64)
        repr32, repr16 = torch.empty(M, C, device=dev, dtype=F32), torch.empty(M, C, device=dev, dtype=BF16)
        ops.som_select(dots, node_sq, snap, idx, idx64, None, None, repr32, repr16)
        return repr32, idx64

    def forward(self, x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Best-matching unit per input vector (kohonen.py:100-119).  Standalone calls carry no autograd edge to
        ``nodes``; inside ViT.forward the engine provides it."""
        self._require_cuda(x)
        if x.dim() == 1:
            x = x.unsqueeze(0)
        lead = x.shape[:-1]
        repr32, idx = self._bmu(x.detach().float().contiguous().view(-1, x.shape[-1]))
        return repr32.view(*lead, -1), idx.view(*lead)

    @torch.no_grad()
    def update_nodes(self, x: torch.Tensor, winning_indices: torch.Tensor, learning_rate: float) -> None:
        """Sequential neighbourhood update (kohonen.py:121-165), one kernel."""
        if not self.training:
            return
        self._require_cuda(x)
        C = self.nodes.shape[1]
        B = x.shape[0]
        per = x[0].numel()
        if per < C or per % C != 0:
            raise NotImplementedError("update_nodes: inputs smaller than the node width are not supported")
        flat = winning_indices.reshape(-1).to(torch.int32).contiguous()
        steps = min(flat.numel(), B)
        xf = x.detach().float().contiguous()
        if per > C:
            pooled = torch.empty(B, C, device=x.device, dtype=F32)
            ops.som_pool(xf, B * C, per // C, pooled)
        else:
            pooled = xf.view(B, C)
        coef = torch.tensor([float(learning_rate) * self.alpha], device=x.device, dtype=F32)
        nodes = self.nodes.data
        if nodes.dtype != F32 or not nodes.is_contiguous():
            raise TypeError("KohonenMap.nodes must be contiguous fp32")
        ops.som_update(nodes, pooled, flat, steps, self.m, self.n, coef, self.sigma)
