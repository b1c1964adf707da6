"""Per-CTA wall time of the persistent attention-backward kernels (hooks build: globaltimer at kernel entry / exit of every CTA,
%smid of the SM it ran on).  Usage: python scripts/attn_bwd_cta_times.py [variant ...]"""
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
for v in [int(a) for a in sys.argv[1:]] or [2, 3]:
    _lib.call("nvit_attention_bwd_variant", v)
    for _ in range(3):
        bwd()
    buf = torch.zeros(16384, dtype=torch.int64, device=dev)
    _lib.call("nvit_attention_debug", buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bwd()
    e1.record()
    torch.cuda.synchronize()
    _lib.call("nvit_attention_debug", None)
    h = buf.cpu()
    G = int(os.environ.get("NVIT_ATTN_GRID", 148))
    ct = h[512:512 + 2 * G].view(G, 2)
    sm = h[808:808 + G].tolist()
    t0 = int(ct[:, 0].min())
    dur = ((ct[:, 1] - ct[:, 0]).float() / 1e3).tolist()
    print(f"variant {v}: launch {e0.elapsed_time(e1) * 1e3:.1f} us; last exit {(int(ct[:, 1].max()) - t0) / 1e3:.1f} us after the first entry")
    print("  duration (us) by SM id:")
    by_sm = sorted(zip(sm, dur, range(G)))
    for i in range(0, G, 16):
        print("   ", " ".join(f"{s}:{d:.0f}" for s, d, _ in by_sm[i:i + 16]))
    print("  CTA -> SM:", " ".join(str(s) for s in sm[:32]), "...")
