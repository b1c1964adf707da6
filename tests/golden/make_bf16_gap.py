"""Calibration fixture: the REFERENCE'S OWN bf16-autocast vs fp32 gradient gap on the closed-formula cases.

    python tests/golden/make_bf16_gap.py        (build container only: imports /root/reference)

The formula fixtures are ill-conditioned under bf16 (per-sample gradients cancel in the batch sum), so "matches the
reference within bf16 tolerance" is stated per tensor relative to what the unmodified reference itself loses when run
under torch.autocast(bf16): tests/test_model_gpu.py accepts err <= max(3e-2*|g|, 3*gap[name], 1e-3*|g|_global).
Writes tests/golden/bf16_gap.npz with keys "<case>:<param>" -> L2 norm of (grad_bf16_autocast - grad_fp32) and
"<case>:__logits_rel__" -> rel-L2 of the bf16-autocast logits against the fp32 logits.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.append("/root/reference")
from oracle import nvit_oracle as O  # noqa: E402
import nvit.model as ref  # noqa: E402

CASES = {"micro": ("micro", dict(), 4), "micro_bias": ("micro", dict(bias=True), 5), "mini_bs32": ("mini", dict(base_scale=1.0 / 32.0), 3),
         "micro_orig": ("micro", dict(use_nvit=False), 4), "micro_bias4": ("micro", dict(bias=True), 4)}

if __name__ == "__main__":
    torch.set_num_threads(1)
    out = {}
    for tag, (name, over, batch) in CASES.items():
        cfg = O.named_config(name, **over)
        sd = O.formula_state_dict(cfg)
        X, y = O.formula_batch(cfg, batch)
        grads = {}
        logit = {}
        for mode in ("fp32", "bf16"):
            m = ref.ViT(ref.ViTConfig(**cfg.as_dict()))
            if not cfg.use_nvit:       # the reference's own original-ViT mode needs the norms attached
                for blk in m.transformer.h:
                    blk.rmsnorm_att = ref.RMSNorm(cfg.n_embd)
                    blk.rmsnorm_mlp = ref.RMSNorm(cfg.n_embd)
            m.load_state_dict(sd)
            m.train()
            if mode == "bf16":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    logits, _ = m(X)
                    loss = F.cross_entropy(logits, y)
            else:
                logits, _ = m(X)
                loss = F.cross_entropy(logits, y)
            loss.backward()
            logit[mode] = logits.detach().double()
            grads[mode] = {k: p.grad.detach().double() for k, p in m.named_parameters() if p.grad is not None}
        out[f"{tag}:__logits_rel__"] = np.float64(((logit["bf16"] - logit["fp32"]).norm() / logit["fp32"].norm()).item())
        for k, g in grads["fp32"].items():
            out[f"{tag}:{k}"] = np.float64((grads["bf16"][k] - g).norm().item())
    np.savez_compressed(os.path.join(HERE, "bf16_gap.npz"), **out)
    print("wrote", len(out), "entries")
