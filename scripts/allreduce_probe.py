"""NCCL all-reduce of the flat fp32 gradient buffer of nViT-B/16 (121.5 M floats = 486 MB) in isolation: eager and from a
captured CUDA graph, time per call (CUDA events, max over ranks) and bus bandwidth.  torchrun --nproc-per-node N."""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 121_520_348
x = torch.randn(n, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def report(tag, ms, nbytes):
    if rank == 0:
        bus = 2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9
        print(f"{tag}: {ms:.3f} ms per call, algbw {nbytes / (ms * 1e-3) / 1e9:.0f} GB/s, busbw {bus:.0f} GB/s", flush=True)


ms = timed(lambda: dist.all_reduce(x))
report(f"eager all_reduce fp32 x{n}", ms, 4 * n)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(x)
ms = timed(g.replay)
report("graph-captured all_reduce", ms, 4 * n)
for parts in (2, 4):
    chunks = x.chunk(parts)
    ms = timed(lambda: [dist.all_reduce(c) for c in chunks])
    report(f"eager, {parts} chunks back to back", ms, 4 * n)
xb = x.to(torch.bfloat16)
ms = timed(lambda: dist.all_reduce(xb))
report("eager all_reduce bf16 (for scale only: not used)", ms, 2 * n)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
