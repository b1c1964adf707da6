"""Secondary baselines of SURVEY.md 8d: the UNMODIFIED reference model (baseline/_ref/nvit/model.py) on one B200 under
bf16 autocast + GradScaler (train.py:135-136, 254), driven by the restated step of baseline/ref_step.py:
eager, and torch.compile after >= 10 warm-up steps (ViT.forward mutates self.step, so the first 8 steps recompile and
dynamo then falls back - SURVEY.md 2.3 #6).  Prints one JSON line per mode; nothing of this repo's engine is involved.

    python scripts/ref_gpu_baseline.py [--config b16] [--batch 256] [--steps 10] [--modes eager,compile]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from baseline import ref_step  # noqa: E402


def run(mode, cfg, batch, steps, warmup):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.cuda.reset_peak_memory_stats()
    st = ref_step.ReferenceStepper(cfg, dev, amp_dtype=torch.bfloat16, compile_model=(mode == "compile"))
    X, y = ref_step.synthetic_batch(cfg, batch)
    X, y = X.to(dev), y.to(dev)
    t0 = time.perf_counter()
    for _ in range(warmup):
        loss = st.step(X, y)
    torch.cuda.synchronize()
    warm_s = time.perf_counter() - t0
    sampler = bench.ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for _ in range(steps):
        loss = st.step(X, y)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "reference-gpu", "mode": mode, "metric": bench.METRIC, "value": batch / (ms / 1e3), "unit": bench.UNIT,
            "ms_per_step": ms, "batch": batch, "steps": steps, "warmup": warmup, "warmup_s": warm_s, "dtype": "bf16 autocast + GradScaler",
            "final_loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "clocks": clocks,
            "what": "unmodified reference nvit/model.py (baseline/_ref) + restated train.py step, 1 B200"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="b16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--modes", default="eager,compile")
    args = ap.parse_args()
    cfg = bench.config_dict(args.config)
    for mode in args.modes.split(","):
        batch = args.batch
        while batch >= 16:
            try:
                line = run(mode, cfg, batch, args.steps, 12 if mode == "compile" else 3)
                print(json.dumps(line), flush=True)
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                print(json.dumps({"impl": "reference-gpu", "mode": mode, "batch": batch, "error": "out of memory"}), flush=True)
                batch //= 2
            except Exception as e:      # e.g. torch.compile failing on the reference's forward
                print(json.dumps({"impl": "reference-gpu", "mode": mode, "batch": batch, "error": f"{type(e).__name__}: {str(e)[:300]}"}), flush=True)
                break
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
