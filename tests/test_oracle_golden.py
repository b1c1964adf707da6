"""Oracle vs committed golden vectors (outputs of the real reference on formula weights).

Runs anywhere (no /root/reference needed).  tests/golden/make_golden.py is the generator.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nvit_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "micro_nvit": ("micro", dict(), 4),
    "micro_nvit_bias": ("micro", dict(bias=True), 4),
    "mini_nvit_bs32": ("mini", dict(base_scale=1.0 / 32.0), 3),
    "micro_orig": ("micro", dict(use_nvit=False), 4),
}


def load_case(tag):
    name, over, batch = CASES[tag]
    cfg = O.named_config(name, **over)
    gold = dict(np.load(os.path.join(GOLDEN, tag + ".npz")))
    return cfg, batch, gold


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_forward_backward_matches_golden(tag):
    torch.set_num_threads(1)
    cfg, batch, gold = load_case(tag)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.formula_state_dict(cfg).items()}
    X, y = O.formula_batch(cfg, batch)
    logits, aux = O.vit_forward(sd, cfg, X)
    ce = F.cross_entropy(logits, y)
    (ce + 0.1 * aux["reconstruction"]).backward()
    np.testing.assert_allclose(logits.detach().numpy(), gold["logits"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(float(ce.detach()), float(gold["ce"]), rtol=1e-5)
    np.testing.assert_allclose(float(aux["reconstruction"]), float(gold["reconstruction"]), rtol=1e-5)
    n_checked = 0
    gmax = max(float(v) for kk, v in gold.items() if kk.startswith("gnorm:"))
    for k, p in sd.items():
        key = "gnorm:" + k
        if key not in gold:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        gn = float(gold[key])
        assert abs(float(p.grad.double().norm()) - gn) <= 2e-3 * gn + 1e-6 * gmax, k
        np.testing.assert_allclose(p.grad.flatten()[:8].numpy(), gold["ghead:" + k], rtol=5e-3, atol=2e-3 * gn / max(1.0, p.numel() ** 0.5) + 1e-6 * gmax, err_msg=k)
        n_checked += 1
    assert n_checked > 20


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_step_matches_golden(tag):
    torch.set_num_threads(1)
    cfg, batch, gold = load_case(tag)
    tr = O.OracleTrainer(O.formula_state_dict(cfg), cfg, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0)
    X, y = O.formula_batch(cfg, batch)
    tr.step(X, y)
    with torch.no_grad():
        logits2, _ = O.vit_forward(tr.sd, cfg, X)
    np.testing.assert_allclose(logits2.numpy(), gold["logits_after_step"], rtol=5e-4, atol=5e-5)
    for key, val in gold.items():
        if key.startswith("whead:"):
            np.testing.assert_allclose(tr.sd[key[6:]].detach().flatten()[:8].numpy(), val, rtol=1e-4, atol=1e-6, err_msg=key)
    if cfg.use_nvit:
        for i in range(cfg.n_layer):
            for nm, dim in O.NORMALIZED:
                w = tr.sd[f"transformer.h.{i}.{nm}.weight"].detach()
                assert float((w.norm(dim=dim) - 1).abs().max()) < 1e-3


KOHONEN_CASES = {"micro_kohonen": ("micro", dict(use_kohonen=True, kohonen_nodes=32, kohonen_alpha=0.3), 4)}


@pytest.mark.parametrize("tag", list(KOHONEN_CASES))
def test_oracle_kohonen_matches_golden(tag):
    """BASELINE config 5: units, in-forward map update, the five auxiliary losses and all gradients vs the reference."""
    torch.set_num_threads(1)
    name, over, batch = KOHONEN_CASES[tag]
    cfg = O.named_config(name, **over)
    gold = dict(np.load(os.path.join(GOLDEN, tag + ".npz")))
    sd = {k: v.clone().requires_grad_(True) for k, v in O.formula_state_dict(cfg).items()}
    X, y = O.formula_batch(cfg, batch)
    logits, aux = O.vit_forward(sd, cfg, X, training=True, step=1)
    ce = F.cross_entropy(logits, y)
    loss = ce + 0.1 * aux["kohonen_consistency"] + 0.1 * aux["kohonen_smoothness"] \
        + cfg.local_quantization_weight * aux["local_quantization"] + cfg.global_quantization_weight * aux["global_quantization"] \
        + cfg.reconstruction_weight * aux["reconstruction"]
    loss.backward()
    assert np.array_equal(aux["_local_indices"].numpy(), gold["local_indices"])
    assert np.array_equal(aux["_global_indices"].numpy(), gold["global_indices"])
    np.testing.assert_allclose(logits.detach().numpy(), gold["logits"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(float(loss.detach()), float(gold["loss"]), rtol=1e-5)
    for k in ("kohonen_consistency", "kohonen_smoothness", "local_quantization", "global_quantization", "reconstruction"):
        np.testing.assert_allclose(float(aux[k]), float(gold["aux:" + k]), rtol=2e-5, err_msg=k)
    np.testing.assert_allclose(sd["local_kohonen.nodes"].detach().numpy(), gold["local_nodes_after"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(sd["global_kohonen.nodes"].detach().numpy(), gold["global_nodes_after"], rtol=1e-5, atol=1e-7)
    gmax = max(float(v) for kk, v in gold.items() if kk.startswith("gnorm:"))
    n_checked = 0
    for k, p in sd.items():
        key = "gnorm:" + k
        if key not in gold:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        gn = float(gold[key])
        assert abs(float(p.grad.double().norm()) - gn) <= 2e-3 * gn + 1e-6 * gmax, k
        np.testing.assert_allclose(p.grad.flatten()[:8].numpy(), gold["ghead:" + k], rtol=5e-3, atol=2e-3 * gn / max(1.0, p.numel() ** 0.5) + 1e-6 * gmax, err_msg=k)
        n_checked += 1
    assert n_checked > 25 and "gnorm:local_kohonen.nodes" in gold and "gnorm:reconstruction_head.0.weight" in gold


def test_survey_golden_smoke_tiny(reference_model_module):
    """SURVEY.md section 4 smoke goldens (recipe: manual_seed(0) reference init, generator 1234 batch)."""
    ref = reference_model_module
    cfg = O.named_config("tiny")
    torch.manual_seed(0)
    m = ref.ViT(ref.ViTConfig(**cfg.as_dict()))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(64, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (64,), generator=g)
    logits, aux = O.vit_forward(sd, cfg, X)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    assert abs(float(loss) - 2.323694) < 5e-5
    assert abs(float(logits.abs().sum()) - 120.79670) < 2e-2
    assert abs(float(aux["reconstruction"]) - 1.003778) < 5e-5
    assert abs(float(sd["transformer.h.0.query.weight"].grad.norm()) - 3.138280e-2) < 1e-5
    assert abs(float(sd["transformer.h.0.attn_alpha"].grad.norm()) - 4.265926e-2) < 1e-5
    assert abs(float(sd["sz"].grad.norm()) - 3.216266e-2) < 1e-5
