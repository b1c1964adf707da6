"""CUDA-event timing of the attention kernels at the bench shapes (B = 256, H = 12, T = 196, pre-normalised q / k), both
backward variants (nvit_attention_bwd_variant)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import ops, _lib

B, H, T = int(os.environ.get("B", 256)), int(os.environ.get("H", 12)), int(os.environ.get("T", 196))
C = H * 64
M = B * T
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(M, 3 * C, device=dev, generator=g) * 0.5).to(torch.bfloat16)
sqk = torch.full((C,), 0.036, device=dev)
with torch.no_grad():
    heads = qkv[:, :2 * C].float().view(M, 2 * H, 64)
    nrm = heads.norm(dim=-1, keepdim=True)
    qkv[:, :2 * C] = (heads / nrm).reshape(M, 2 * C).to(torch.bfloat16)
    inv = (1.0 / nrm[..., 0]).contiguous()
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
do = (torch.randn(M, C, device=dev, generator=g) * 0.1).to(torch.bfloat16)
dqkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
dsqk = torch.zeros(C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
kw = dict(inv_q=inv[:, :H], inv_k=inv[:, H:])
fwd = lambda: ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, lse, B, H, T, **kw)
bwd = lambda: ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, do, lse,
                                dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T, **kw)


def timeit(f, n=20):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print('   ', ' '.join(f'{t:.0f}' for t in ts))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


fwd()
for v in (1, 2):
    _lib.call("nvit_attention_fwd_variant", v)
    med, best = timeit(fwd)
    print(f"attention fwd variant {v}: median {med:.1f} us, best {best:.1f} us   ({4.0 * B * H * T * T * 64 / 1e12:.4f} TFLOP)")
for v in (1, 2, 3):
    _lib.call("nvit_attention_bwd_variant", v)
    med, best = timeit(bwd)
    fl = 10.0 * B * H * T * T * 64
    print(f"attention bwd variant {v}: median {med:.1f} us, best {best:.1f} us  -> {fl / med / 1e6:.1f} TFLOP/s")
