"""Phase timing of the attention kernels (clock64 marks of thread 0 in the first 8 CTAs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import os as _os  # measurement hooks live in the -DNVIT_BENCH_HOOKS build (python -m nvit_b200.build --hooks)
_os.environ.setdefault("NVIT_LIB_PATH", _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nvit_b200", "libnvit_b200_hooks.so"))
from nvit_b200 import ops, _lib

B, H, T, C = 256, 12, 196, 768
M = B * T
dev = "cuda"
qkv = torch.randn(M, 3 * C, device=dev, dtype=torch.bfloat16) * 0.5
sqk = torch.full((C,), 0.036, device=dev)
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
do = torch.randn(M, C, device=dev, dtype=torch.bfloat16) * 0.1
dqkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
dsqk = torch.zeros(C, device=dev)
buf = torch.zeros(16384, dtype=torch.int64, device=dev)
inv = torch.rand(M, 2 * H, device=dev) + 0.5 if os.environ.get("PRENORM") == "1" else None    # timing only: values arbitrary
kw = {} if inv is None else dict(inv_q=inv[:, :H], inv_k=inv[:, H:])
fwd = lambda: ops.attention_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, lse, B, H, T, **kw)
bwd = lambda: ops.attention_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], sqk, 1 / 0.036, 8.0, out, do, lse,
                                dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], dsqk, B, H, T, **kw)
for name, f, marks in (("fwd", fwd, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11]), ("bwd", bwd, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 14, 15, 16, 17, 24, 25])):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    buf.zero_()
    _lib.call("nvit_attention_debug", buf.data_ptr())
    f()
    torch.cuda.synchronize()
    _lib.call("nvit_attention_debug", None)
    t = buf.view(8, 32).cpu()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1000:.0f} us per launch; cycles since mark 0, CTAs 0..3 (first wave):")
    for cta in range(4):
        row = t[cta]
        print("   cta", cta, " ".join(f"m{m}:{int(row[m] - row[0])}" for m in marks if row[m] > 0))
