"""Residual backward, staged form: warps per CTA (and with them the ring depth that fits 227 KB) at the bench shapes.
Hooks build only (NVIT_RES_W is read by libnvit_b200_hooks.so).  Prints us per launch and algorithmic GB/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NVIT_LIB_PATH", os.path.join(ROOT, "nvit_b200", "libnvit_b200_hooks.so"))
sys.path.insert(0, ROOT)
import torch
from nvit_b200 import ops, _lib

DEV = "cuda"
M, C = 50176, int(sys.argv[1]) if len(sys.argv) > 1 else 768
g = torch.Generator().manual_seed(0)
mk = lambda dt=torch.float32: torch.randn(M, C, generator=g).to(DEV).to(dt)
gr, h, h0, x = mk(), mk(), mk(), mk(torch.bfloat16)
alpha = torch.full((C,), C ** -0.5, device=DEV)
skip = torch.tensor([0.9], device=DEV)
dh, dh0 = torch.zeros(M, C, device=DEV), torch.empty(M, C, device=DEV)
dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
dalpha, dskip = torch.zeros(C, device=DEV), torch.zeros(1, device=DEV)
flush = torch.empty(64 << 20, device=DEV)


def timed(fn, n=10):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for sk, acc in ((True, False), (False, True), (False, False)):
    nbytes = M * C * ((4 + 4 + 2) + (4 + 2) + (8 if sk else 0) + (4 if acc else 0))
    fn = lambda: ops.residual_bwd(gr, h, x, alpha, 0.05 * C ** 0.5, dh, dx, dalpha, dh_accumulate=acc, h0=h0 if sk else None,
                                  skip=skip if sk else None, dh0=dh0 if sk else None, dskip=dskip if sk else None)
    out = []
    for rep in range(2):
        for mode, w in ((0, 8), (1, 8), (1, 6), (1, 4)):
            _lib.call("nvit_residual_bwd_staged", mode)
            os.environ["NVIT_RES_W"] = str(w)
            us = timed(fn)
            out.append(f"{'regs' if mode == 0 else 'W=' + str(w)} {us:.1f} us ({nbytes / us / 1e3:.0f} GB/s)")
    print(f"skip={sk} acc={acc} ({nbytes / 1e6:.0f} MB): " + "  ".join(out), flush=True)
