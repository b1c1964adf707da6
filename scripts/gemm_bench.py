"""GEMM micro-benchmark: TFLOP/s of nvit_gemm_bf16 on the step's shapes, per CTA-group mode and debug mode.
Usage: python scripts/gemm_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import os as _os  # measurement hooks live in the -DNVIT_BENCH_HOOKS build (python -m nvit_b200.build --hooks)
_os.environ.setdefault("NVIT_LIB_PATH", _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nvit_b200", "libnvit_b200_hooks.so"))
from nvit_b200 import ops, _lib

M = 50176
SHAPES = [  # name, N, K, kind
    ("qkv fwd", 2304, 768, "fwd"), ("att_c_proj fwd", 768, 768, "fwd"), ("mlp_c_proj fwd", 768, 3072, "fwd"),
    ("c_fc swiglu", 3072, 768, "swiglu"), ("mlp_c_proj dgrad", 3072, 768, "dgrad"), ("c_fc dgrad acc", 768, 6144, "dgrad_acc"),
    ("qkv fwd + qk norm", 2304, 768, "qknorm"), ("c_fc wgrad", 6144, 768, "wgrad"), ("qkv wgrad", 2304, 768, "wgrad"), ("mlp_c_proj dgrad+gate", 3072, 768, "gate_bwd"),
]


def run(name, N, K, kind, iters=20):
    dev = "cuda"
    if kind == "wgrad":          # dW[N,K] = dY[M,N]^T X[M,K]
        dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        dw = torch.zeros(N, K, device=dev)
        f = lambda: ops.linear_wgrad(dy, x, dw, splits=0, accumulate=True)
        flops = 2.0 * M * N * K
    elif kind in ("dgrad", "dgrad_acc"):      # dX[M,N] = dY[M,K] W[K,N]
        dy = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        dx = torch.zeros(M, N, device=dev, dtype=torch.float32 if kind == "dgrad_acc" else torch.bfloat16)
        f = lambda: ops.linear_dgrad(dy, w, dx, accumulate=(kind == "dgrad_acc"))
        flops = 2.0 * M * N * K
    elif kind == "qknorm":
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        sc = torch.ones(K, device=dev)
        inv = torch.empty(M, 2 * K // 64, device=dev)
        f = lambda: ops.gemm_qknorm(x, w, out, sc, 1.0, K, 2 * K, inv)
        flops = 2.0 * M * N * K
    elif kind == "gate_bwd":     # d(uv)[M, 2N] = gate backward of dY[M,K] W[K,N]
        dy = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        uv = torch.randn(M, 2 * N, device=dev, dtype=torch.bfloat16)
        suv = torch.ones(2 * N, device=dev)
        duv = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
        f = lambda: ops.gemm_gate_bwd(dy, w, uv, suv, 1.0, duv)
        flops = 2.0 * M * N * K
    elif kind == "swiglu":
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(2 * N, K, device=dev, dtype=torch.bfloat16)
        suv = torch.ones(2 * N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        raw = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
        f = lambda: ops.gemm(x, w, out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, colscale=suv, colscale_mul=1.0, c2=raw, ldc2=2 * N, swiglu_half=N)
        flops = 2.0 * M * 2 * N * K
    else:
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        f = lambda: ops.linear_fwd(x, w, out)
        flops = 2.0 * M * N * K
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, flops / ms / 1e9


if __name__ == "__main__":
    modes = [int(x) for x in os.environ.get("DBG_MODES", "0,1,2").split(",")]
    cgs = [int(x) for x in os.environ.get("CG_MODES", "1,2").split(",")]
    print(f"{'shape':22s} " + " ".join(f"cg{cg}/dbg{d:<7d}" for cg in cgs for d in modes))
    only = os.environ.get("ONLY")
    for name, N, K, kind in SHAPES:
        if only and only not in name:
            continue
        cells = []
        for cg in cgs:
            for dbg in modes:
                _lib.call("nvit_gemm_force_cta_group", cg)
                _lib.call("nvit_gemm_debug", dbg)
                ms, tf = run(name, N, K, kind)
                cells.append(f"{ms*1000:6.0f}us {tf:5.0f}")
        print(f"{name:22s} " + " | ".join(cells))
    _lib.call("nvit_gemm_debug", 0)
    _lib.call("nvit_gemm_force_cta_group", 0)
