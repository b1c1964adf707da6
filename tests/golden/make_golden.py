"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Weights/inputs come from closed formulas in oracle.nvit_oracle (formula_state_dict /
formula_batch) so the fixtures only need to hold outputs.  The reference model
(/root/reference/nvit/model.py, untouched) is executed in fp32 on CPU: one forward+backward
with loss = CE + 0.1*reconstruction, then the restated optimizer tail for the step fixture.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.append("/root/reference")

from oracle import nvit_oracle as O  # noqa: E402
import nvit.model as ref  # noqa: E402

CASES = {
    "micro_nvit": ("micro", dict(), 4),
    "micro_nvit_bias": ("micro", dict(bias=True), 4),
    "mini_nvit_bs32": ("mini", dict(base_scale=1.0 / 32.0), 3),
    "micro_orig": ("micro", dict(use_nvit=False), 4),
}


def run(name, over, batch):
    cfg = O.named_config(name, **over)
    sd = O.formula_state_dict(cfg)
    m = ref.ViT(ref.ViTConfig(**cfg.as_dict()))
    if not cfg.use_nvit:
        for blk in m.transformer.h:
            blk.rmsnorm_att = ref.RMSNorm(cfg.n_embd)
            blk.rmsnorm_mlp = ref.RMSNorm(cfg.n_embd)
    m.load_state_dict(sd, strict=True)
    m.train()
    X, y = O.formula_batch(cfg, batch)
    logits, aux = m(X)
    ce = F.cross_entropy(logits, y)
    loss = ce + 0.1 * aux["reconstruction"]
    loss.backward()
    out = {"logits": logits.detach().numpy(), "ce": ce.detach().numpy(),
           "reconstruction": aux["reconstruction"].detach().numpy()}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().flatten()
        out["gnorm:" + k] = np.float64(g.double().norm().item())
        out["ghead:" + k] = g[:8].numpy().copy()
    # one full training step (clip 1.0, AdamW lr 1e-3 b(0.9,0.95) wd 0.1, normalize_matrices) with CE only
    m.zero_grad(set_to_none=True)
    opt = m.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu")
    logits, _ = m(X)
    F.cross_entropy(logits, y).backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    if cfg.use_nvit:
        with torch.no_grad():
            for blk in m.transformer.h:
                for nm, dim in O.NORMALIZED:
                    w = getattr(blk, nm).weight
                    w.copy_(w / w.norm(p=2, dim=dim, keepdim=True))
    m.eval()
    with torch.no_grad():
        logits2, _ = m(X)
    out["logits_after_step"] = logits2.numpy()
    for k in ("transformer.h.0.query.weight", "transformer.h.1.mlp_c_proj.weight", "transformer.h.0.suv", "sz",
              "cross_attention.proj.weight"):
        if k in dict(m.named_parameters()):
            out["whead:" + k] = dict(m.named_parameters())[k].detach().flatten()[:8].numpy().copy()
    return out


KOHONEN_CASES = {
    "micro_kohonen": ("micro", dict(use_kohonen=True, kohonen_nodes=32, kohonen_alpha=0.3), 4),
}
KOHONEN_WEIGHTS = dict(kohonen_consistency=0.1, kohonen_smoothness=0.1)     # settings.yaml:15-16


def run_kohonen(name, over, batch):
    """BASELINE config 5 shape of the forward: one training-mode forward+backward with the train.py:906-926 loss."""
    cfg = O.named_config(name, **over)
    sd = O.formula_state_dict(cfg)
    m = ref.ViT(ref.ViTConfig(**cfg.as_dict()))
    m.load_state_dict({**sd, **O.kohonen_buffers(cfg)}, strict=True)
    m.train()
    X, y = O.formula_batch(cfg, batch)
    idx = {}
    for tag in ("local", "global"):
        mod = getattr(m, tag + "_kohonen")
        mod.register_forward_hook(lambda _m, _i, out, tag=tag: idx.__setitem__(tag, out[1].detach().clone()))
    logits, aux = m(X)
    ce = F.cross_entropy(logits, y)
    loss = ce + 0.1 * aux["kohonen_consistency"] + 0.1 * aux["kohonen_smoothness"] \
        + cfg.local_quantization_weight * aux["local_quantization"] + cfg.global_quantization_weight * aux["global_quantization"] \
        + cfg.reconstruction_weight * aux["reconstruction"]
    loss.backward()
    out = {"logits": logits.detach().numpy(), "ce": ce.detach().numpy(), "loss": loss.detach().numpy()}
    for k, v in aux.items():
        out["aux:" + k] = v.detach().numpy()
    out["local_indices"], out["global_indices"] = idx["local"].numpy(), idx["global"].numpy()
    print("   distinct units:", len(np.unique(out["local_indices"])), len(np.unique(out["global_indices"])))
    out["local_nodes_after"] = m.local_kohonen.nodes.detach().numpy().copy()
    out["global_nodes_after"] = m.global_kohonen.nodes.detach().numpy().copy()
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().flatten()
        out["gnorm:" + k] = np.float64(g.double().norm().item())
        out["ghead:" + k] = g[:8].numpy().copy()
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(1)
    only = sys.argv[1:]
    for tag, (name, over, batch) in CASES.items():
        if only and tag not in only:
            continue
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **run(name, over, batch))
        print("wrote", tag)
    for tag, (name, over, batch) in KOHONEN_CASES.items():
        if only and tag not in only:
            continue
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **run_kohonen(name, over, batch))
        print("wrote", tag)
