"""Phase marks of the persistent attention forward (hooks build): softmax warpgroups 0 / 1 and the MMA warp, 4th head of the
fastest / median / slowest CTA.  Marks: warpgroup: 0 top, 1 head known, 2 S ready, 3 P written, 4 O ready, 5 O rows staged;
MMA: 0 top, 1 Q/K landed, 2/3 S_0/S_1 issued, 4 V landed, 5/6 O_0/O_1 issued."""
import os
import sys

os.environ.setdefault("NVIT_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nvit_b200", "libnvit_b200_hooks.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import ops, _lib

B, H, T = int(os.environ.get("B", 256)), int(os.environ.get("H", 12)), int(os.environ.get("T", 196))
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
kw = dict(inv_q=inv[:, :H], inv_k=inv[:, H:])
fwd = lambda: ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, lse, B, H, T, **kw)
_lib.call("nvit_attention_fwd_variant", 2)
for _ in range(3):
    fwd()
buf = torch.zeros(16384, dtype=torch.int64, device=dev)
_lib.call("nvit_attention_debug", buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fwd()
e1.record()
torch.cuda.synchronize()
print(f"marked launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
_lib.call("nvit_attention_debug", None)
h = buf.cpu()
ct = h[512:512 + 296].view(148, 2)
dur = ((ct[:, 1] - ct[:, 0]).float() / 1e3).tolist()
order = sorted(range(148), key=lambda i: dur[i])
print("CTA durations (us), sorted:", " ".join(f"{dur[i]:.0f}" for i in order[::8]))
for tag, cta in (("fastest", order[0]), ("median", order[74]), ("slowest", order[-1])):
    row = h[1024 + cta * 96:1024 + cta * 96 + 96]
    t0 = int(row[0])
    print(f"--- {tag} CTA {cta} ({dur[cta]:.0f} us); cycles since warpgroup 0 reached the top of its 4th head")
    for name, off in (("warpgroup 0", 0), ("warpgroup 1", 32), ("MMA warp", 64)):
        print(f"  {name}:", {i: int(row[off + i]) - t0 for i in range(8) if int(row[off + i]) != 0 or (off == 0 and i == 0)})
