"""Same-box A/B of a process-wide on/off hook of the library on the training step: nViT-B/16, batch 256, one GPU.

    python scripts/step_ab.py [--hook nvit_set_pdl] [--steps 10] [--reps 3] [--config b16] [--batch 256] [--graph-only]

Hooks: nvit_set_pdl (programmatic dependent launch), nvit_residual_bwd_staged (form of the residual backward kernel).

Alternates the two modes inside one process (same clocks, same allocations): for each mode the step graph is captured
again and `steps` replays are timed with CUDA events; the eager (launch-from-Python) path is timed the same way.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nvit_b200 import ViT, ViTConfig, Trainer, _lib
from oracle import nvit_oracle as O   # config table only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--config", default="b16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hook", default="nvit_set_pdl")
    ap.add_argument("--graph-only", action="store_true")
    ap.add_argument("--modes", default="0,1", help="the two hook values to alternate")
    ap.add_argument("--restore", type=int, default=None, help="hook value to leave behind (default: the first mode)")
    args = ap.parse_args()
    modes = [int(m) for m in args.modes.split(",")]
    dev = torch.device("cuda", 0)
    cfg = ViTConfig(**O.named_config(args.config).as_dict())
    torch.manual_seed(0)
    model = ViT(cfg).to(dev).train()
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(args.batch, 3, cfg.image_size, cfg.image_size, generator=g).to(dev)
    y = torch.randint(0, cfg.num_classes, (args.batch,), generator=g).to(dev)

    def timed(tr, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            loss = tr.step(X, y)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n, float(loss)

    for graph in ((True,) if args.graph_only else (True, False)):
        tr = Trainer(model, learning_rate=1e-3, cuda_graph=graph)
        for _ in range(3):
            tr.step(X, y)
        for rep in range(args.reps):
            for mode in modes:
                _lib.call(args.hook, mode)
                tr._graph = None                 # capture again under the new launch attribute
                tr.step(X, y)
                tr.step(X, y)
                ms, loss = timed(tr, args.steps)
                print(f"{'graph' if graph else 'eager'} {args.hook}={mode} rep={rep}: {ms:.3f} ms/step  ({args.batch / ms * 1e3:.0f} images/s)  loss {loss:.4f}",
                      flush=True)
    _lib.call(args.hook, modes[0] if args.restore is None else args.restore)


if __name__ == "__main__":
    main()
