"""Tile order of the persistent GEMM grids (nvit_gemm_raster_group) on the step's shapes: n fastest (0) vs bands.

    python scripts/raster_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nvit_b200 import _lib
import gemm_bench

SHAPES = [s for s in gemm_bench.SHAPES if s[0] in ("c_fc swiglu", "mlp_c_proj dgrad+gate", "qkv fwd + qk norm", "mlp_c_proj dgrad")]
GROUPS = (0, 8, 0, 8, -1, 4, 6, 12)
print(f"{'shape':24s} " + " ".join(f"{('G=' + str(g)):>13s}" for g in GROUPS))
for name, N, K, kind in SHAPES:
    cells = []
    for g in GROUPS:
        _lib.call("nvit_gemm_raster_group", g)
        ms, tf = gemm_bench.run(name, N, K, kind)
        cells.append(f"{ms * 1000:6.0f}us {tf:5.0f}")
    print(f"{name:24s} " + " ".join(cells), flush=True)
_lib.call("nvit_gemm_raster_group", 0)
