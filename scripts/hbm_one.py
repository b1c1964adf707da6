"""One launch of each HBM-bound kernel variant of the step at the bench shapes (for `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum -k regex:residual|weight_norm`): residual forward (plain, skip), residual
backward (the three variants of the step; register form, then the library's automatic choice), weight normalisation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nvit_b200 import ViT, ViTConfig, ops, _lib
from oracle import nvit_oracle as O   # config table only

DEV = "cuda"
M, C = 50176, int(sys.argv[1]) if len(sys.argv) > 1 else 768
g = torch.Generator().manual_seed(0)
mk = lambda dt=torch.float32: torch.randn(M, C, generator=g).to(DEV).to(dt)
gr, h, h0, x = mk(), mk(), mk(), mk(torch.bfloat16)
alpha = torch.full((C,), C ** -0.5, device=DEV)
skip = torch.tensor([0.9], device=DEV)
out32, out16 = torch.empty(M, C, device=DEV), torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
dh, dh0 = torch.zeros(M, C, device=DEV), torch.empty(M, C, device=DEV)
dx = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
dalpha, dskip = torch.zeros(C, device=DEV), torch.zeros(1, device=DEV)
mul = 0.05 * C ** 0.5
ops.residual_fwd(h, x, alpha, mul, out32, out16)
ops.residual_fwd(h, x, alpha, mul, out32, out16, h0=h0, skip=skip)
for mode in (0, 2):
    _lib.call("nvit_residual_bwd_staged", mode)
    ops.residual_bwd(gr, h, x, alpha, mul, dh, dx, dalpha, dh_accumulate=False)
    ops.residual_bwd(gr, h, x, alpha, mul, dh, dx, dalpha, dh_accumulate=True)
    ops.residual_bwd(gr, h, x, alpha, mul, dh, dx, dalpha, dh_accumulate=False, h0=h0, skip=skip, dh0=dh0, dskip=dskip)
torch.cuda.synchronize()
if C == 768:
    model = ViT(ViTConfig(**O.named_config("b16").as_dict())).to(DEV)
    model.engine.param_list()
    model.engine.normalize_matrices()
    torch.cuda.synchronize()
print("done")
