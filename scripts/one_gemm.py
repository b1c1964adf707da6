"""Launch one GEMM shape a few times (for ncu): python scripts/one_gemm.py <shape substring> [cta_group]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import _lib
import gemm_bench as gb

name = sys.argv[1]
cg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.call("nvit_gemm_force_cta_group", cg)
for nm, N, K, kind in gb.SHAPES:
    if name in nm:
        ms, tf = gb.run(nm, N, K, kind, iters=3)
        print(nm, f"{ms * 1000:.0f} us {tf:.0f} TFLOP/s")
