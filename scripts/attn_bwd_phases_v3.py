"""Phase timing of the attention backward with the epilogue warpgroup (variant 3; clock64 marks, hooks build): compute thread 0,
the MMA warp and thread 0 of the epilogue warpgroup, second head of the first two CTAs."""
import os
import sys

os.environ.setdefault("NVIT_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nvit_b200", "libnvit_b200_hooks.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import ops, _lib

B, H, T = 256, 12, 196
C = H * 64
M = B * T
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
ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, lse, B, H, T, **kw)
bwd = lambda: ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, do, lse,
                                dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T, **kw)
_lib.call("nvit_attention_bwd_variant", 3)
for _ in range(3):
    bwd()
buf = torch.zeros(16384, dtype=torch.int64, device=dev)
_lib.call("nvit_attention_debug", buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
bwd()
e1.record()
torch.cuda.synchronize()
print(f"marked launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
_lib.call("nvit_attention_debug", None)
h = buf.cpu()
ct = h[512:512 + 296].view(148, 2)
sm = h[808:808 + 148].tolist()
dur = ((ct[:, 1] - ct[:, 0]).float() / 1e3).tolist()
full = [i for i in range(148) if i + 20 * 148 < B * H]          # CTAs that process 21 heads
order = sorted(full, key=lambda i: dur[i])
print("durations (us) of the 21-head CTAs, sorted:", " ".join(f"{dur[i]:.0f}" for i in order))
for tag, cta in (("fastest", order[0]), ("median", order[len(order) // 2]), ("slowest", order[-1])):
    row = h[1024 + cta * 96:1024 + cta * 96 + 96]
    t0 = int(row[0])
    print(f"--- {tag} CTA {cta} (SM {sm[cta]}, {dur[cta]:.0f} us); marks of its 4th head, cycles since the compute warps entered it")
    for name, off in (("compute", 0), ("MMA warp", 32), ("epilogue", 64)):
        print(f"  {name}:", {i: int(row[off + i]) - t0 for i in range(32) if int(row[off + i]) != 0 or (off == 0 and i == 0)})
