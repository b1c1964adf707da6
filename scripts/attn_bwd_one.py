"""One warm-up + a few launches of the attention backward at the bench shapes (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import ops, _lib

B, H, T = 256, 12, 196
C, M = H * 64, B * T
dev = "cuda"
qkv = (torch.randn(M, 3 * C, device=dev) * 0.5).to(torch.bfloat16)
sqk = torch.full((C,), 0.036, device=dev)
heads = qkv[:, :2 * C].float().view(M, 2 * H, 64)
nrm = heads.norm(dim=-1, keepdim=True)
qkv[:, :2 * C] = (heads / nrm).reshape(M, 2 * C).to(torch.bfloat16)
inv = (1.0 / nrm[..., 0]).contiguous()
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
do = (torch.randn(M, C, device=dev) * 0.1).to(torch.bfloat16)
dqkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
dsqk = torch.zeros(C, device=dev)
kw = dict(inv_q=inv[:, :H], inv_k=inv[:, H:])
_lib.call("nvit_attention_bwd_variant", int(os.environ.get("VARIANT", "2")))
ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, lse, B, H, T, **kw)
for _ in range(3):
    ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, do, lse,
                      dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T, **kw)
torch.cuda.synchronize()
print("done")
