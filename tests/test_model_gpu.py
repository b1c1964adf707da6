"""Model-, step- and API-level parity on a B200 (-m gpu).

The CUDA path (nvit_b200.ViT / Trainer, bf16 tensor-core GEMMs + fp32 residual stream) is compared with
  (1) the committed golden fixtures = outputs of the REAL reference on formula weights (tests/golden/*.npz), and
  (2) the fp32 oracle (oracle/nvit_oracle.py, pinned against the reference) run on the same device and inputs.
Stated tolerance (north_star: "within a stated bf16 tolerance, fp32 reference"; calibrated in SURVEY.md section 4 on the
reference's own bf16-autocast vs fp32 gap): logits rel-L2 <= 1e-2; every gradient tensor rel-L2 <= 3e-2 and cosine >=
0.999 (tiny-magnitude tensors: abs <= 1e-3 * global gradient norm); weight rows |norm - 1| <= 1e-3 after a step.

The closed-formula fixtures (micro / mini: C = 64 / 128, batch 3-5, sinusoidal weights and images) are deliberately
ill-conditioned: per-sample gradients nearly cancel in the batch sum.  On them the REFERENCE'S OWN bf16-autocast run
differs from its fp32 run by logits rel-L2 1.4e-2 (micro, bias) and per-tensor gradient rel-L2 of 0.3 ... 10 with an
absolute error up to 0.43 of the global gradient norm (measured in the build container with /root/reference under
torch.autocast("cpu", bf16); per-tensor values committed in tests/golden/bf16_gap.npz by make_bf16_gap.py).  For those
cases the stated tolerance is logits rel-L2 <= 2e-2 and, per gradient tensor, error <= max(3e-2 * |g|, 3 x the
reference's own bf16 gap on that tensor, 1e-3 * global gradient norm); the random-init cases (tiny, B/16) keep the
strict bar.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from nvit_b200 import ViT, ViTConfig, Trainer  # noqa: E402
from oracle import nvit_oracle as O  # noqa: E402

DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def build(cfg: O.OracleConfig, sd):
    m = ViT(ViTConfig(**cfg.as_dict()))
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).train()


def oracle_grads(cfg, sd, X, y):
    p = {k: v.detach().to(DEV).clone().requires_grad_(True) for k, v in sd.items()}
    logits, aux = O.vit_forward(p, cfg, X)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    return logits.detach(), aux["reconstruction"].detach(), loss.detach(), {k: v.grad for k, v in p.items()}


def check_grads(model, ref_grads, tol=3e-2, abs_frac=1e-3, gap=None):
    gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in ref_grads.values() if g is not None)))
    worst = ("", 0.0)
    n = 0
    for name, p in model.named_parameters():
        rg = ref_grads[name]
        if rg is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{name}: reference has no gradient"
            continue
        assert p.grad is not None, f"{name}: missing gradient"
        n += 1
        err = float((p.grad.double() - rg.double()).norm())
        if err <= abs_frac * gnorm and float(rg.norm()) < 3e-2 * gnorm:
            continue                       # tiny tensor: absolute criterion
        if gap is not None and err <= 3.0 * gap.get(name, 0.0):
            continue                       # ill-conditioned fixture: within 3x the reference's own bf16 gap
        r = rel(p.grad, rg)
        if r > worst[1]:
            worst = (name, r)
        assert r <= tol, f"{name}: grad rel-L2 {r:.4f} (|g|={float(rg.norm()):.3e}, global {gnorm:.3e})"
        assert cosine(p.grad, rg) >= 0.999, f"{name}: grad cosine {cosine(p.grad, rg):.5f}"
    assert n > 20
    return worst


CASES = {      # tag -> (config, overrides, batch, key in bf16_gap.npz)
    "micro_nvit": ("micro", dict(), 4, "micro"),
    "micro_nvit_bias": ("micro", dict(bias=True), 4, "micro_bias4"),
    "mini_nvit_bs32": ("mini", dict(base_scale=1.0 / 32.0), 3, "mini_bs32"),
    "micro_orig": ("micro", dict(use_nvit=False), 4, "micro_orig"),      # BASELINE config 4 branch (reference + attached RMSNorms)
}


@pytest.mark.parametrize("tag", list(CASES))
def test_forward_backward_matches_reference_golden(tag):
    """Golden vectors produced by the real reference (tests/golden/make_golden.py): loss = CE + 0.1 * recon there, so
    only logits / CE / reconstruction are compared here; gradients are compared against the oracle below."""
    name, over, batch, gap_key = CASES[tag]
    cfg = O.named_config(name, **over)
    gold = dict(np.load(os.path.join(GOLDEN, tag + ".npz")))
    # logits tolerance: 2e-2, or 1.5 x what the reference itself loses on this ill-conditioned fixture under bf16 autocast
    ref_gap = float(np.load(os.path.join(GOLDEN, "bf16_gap.npz"))[gap_key + ":__logits_rel__"])
    tol = max(2e-2, 1.5 * ref_gap)
    model = build(cfg, O.formula_state_dict(cfg))
    X, y = O.formula_batch(cfg, batch)
    logits, aux = model(X.to(DEV))
    ce = F.cross_entropy(logits, y.to(DEV))
    assert rel(logits.detach(), torch.from_numpy(gold["logits"]).to(DEV)) <= tol
    assert abs(float(ce.detach()) - float(gold["ce"])) <= max(1e-2, ref_gap) * abs(float(gold["ce"]))
    assert abs(float(aux["reconstruction"]) - float(gold["reconstruction"])) <= max(1e-2, ref_gap) * float(gold["reconstruction"])


@pytest.mark.parametrize("name,over,batch,seed", [
    ("micro", dict(), 4, "micro"), ("micro", dict(bias=True), 5, "micro_bias"), ("mini", dict(base_scale=1.0 / 32.0), 3, "mini_bs32"),
    ("tiny", dict(), 16, 0), ("tiny", dict(base_scale=1.0 / 32.0), 64, 1), ("b16", dict(), 2, 0),
    ("tiny", dict(use_nvit=False), 16, 2), ("tiny", dict(use_nvit=False, bias=True), 8, 3), ("tiny", dict(bias=True), 8, 4),
])
def test_logits_and_every_gradient_match_oracle(name, over, batch, seed):
    cfg = O.named_config(name, **over)
    formula = isinstance(seed, str)
    if formula:
        sd = O.formula_state_dict(cfg)
        X, y = O.formula_batch(cfg, batch)
    else:
        sd = O.init_state_dict(cfg, seed)
        g = torch.Generator().manual_seed(1234)
        X = torch.randn(batch, cfg.channels, cfg.image_size, cfg.image_size, generator=g)
        y = torch.randint(0, cfg.num_classes, (batch,), generator=g)
    X, y = X.to(DEV), y.to(DEV)
    ref_logits, ref_recon, ref_loss, ref_grads = oracle_grads(cfg, sd, X, y)
    model = build(cfg, sd)
    logits, aux = model(X)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    assert rel(logits.detach(), ref_logits) <= (2e-2 if formula else 1e-2), rel(logits.detach(), ref_logits)
    assert abs(float(aux["reconstruction"]) - float(ref_recon)) <= 1e-2 * float(ref_recon)
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss))
    gap = None
    if formula:
        allgap = dict(np.load(os.path.join(GOLDEN, "bf16_gap.npz")))
        gap = {k.split(":", 1)[1]: float(v) for k, v in allgap.items() if k.startswith(seed + ":")}
    check_grads(model, ref_grads, gap=gap)
    # parameters that never receive a gradient in the reference stay grad-less
    for n, p in model.named_parameters():
        if (cfg.use_nvit and ".rmsnorm_" in n) or n.startswith("reconstruction_head."):
            assert p.grad is None, n


def test_trainer_step_matches_oracle_step_and_unit_norm_rows():
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(32, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (32,), generator=g).to(DEV)
    ref = O.OracleTrainer({k: v.to(DEV) for k, v in sd.items()}, cfg, lr=1e-3)
    model = build(cfg, sd)
    tr = Trainer(model, learning_rate=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0)
    for it in range(3):
        ref_loss, _, _ = ref.step(X, y)
        loss = tr.step(X, y)
        assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss)), (it, float(loss), float(ref_loss))
    params = dict(model.named_parameters())
    for i in range(cfg.n_layer):
        for nm, dim in O.NORMALIZED:
            w = params[f"transformer.h.{i}.{nm}.weight"].detach()
            assert float((w.norm(dim=dim) - 1).abs().max()) <= 1e-3
            assert rel(w, ref.sd[f"transformer.h.{i}.{nm}.weight"].detach()) <= 2e-2
    # untouched (grad-less) parameters did not move, decayed ones did
    assert torch.equal(params["reconstruction_head.0.weight"].detach().cpu(), sd["reconstruction_head.0.weight"])
    model.eval()
    with torch.no_grad():
        l1, _ = model(X)
        l2, _ = O.vit_forward(ref.sd, cfg, X)
    assert rel(l1, l2) <= 2e-2


def test_trainer_path_equals_autograd_path():
    cfg = O.named_config("mini")
    sd = O.init_state_dict(cfg, 5)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(6, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, cfg.num_classes, (6,), generator=g).to(DEV)
    m1, m2 = build(cfg, sd), build(cfg, sd)
    logits, _ = m1(X)
    F.cross_entropy(logits, y).backward()
    tr = Trainer(m2, grad_clip=0.0)
    tr._ensure_state()
    tr.loss_buf.zero_()
    m2.engine.zero_grad()
    tr.micro_step(X, y)
    torch.cuda.synchronize()
    for n, p in m1.named_parameters():
        if p.grad is not None:
            assert rel(m2.engine.g(n), p.grad) <= 2e-3, n       # same kernels; only atomic summation order differs


def test_state_dict_roundtrip_and_device_move():
    cfg = O.named_config("micro")
    m = ViT(ViTConfig(**cfg.as_dict()))
    keys = list(m.state_dict().keys())
    assert set(keys) == set(O.param_shapes(cfg).keys())
    for k, shp in O.param_shapes(cfg).items():
        assert tuple(m.state_dict()[k].shape) == tuple(shp), k
    m = m.to(DEV)
    X, _ = O.formula_batch(cfg, 2)
    m.eval()
    with torch.no_grad():
        a, _ = m(X.to(DEV))
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        m2 = ViT(ViTConfig(**cfg.as_dict())).to(DEV).eval()
        m2.load_state_dict(sd)
        b, _ = m2(X.to(DEV))
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X)                                              # CPU tensor: no fallback


def test_trainer_optimizer_state_round_trips_through_the_torch_layout():
    """the flat fused-AdamW state exports to / imports from torch.optim.AdamW's state_dict (the layout of the
    reference's checkpoint["optimizer"], train.py:641): a resumed trainer continues exactly like the original."""
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(99)
    X = torch.randn(16, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    m1 = build(cfg, sd)
    t1 = Trainer(m1, learning_rate=2e-3)
    t1.step(X, y)
    t1.step(X, y)
    state = t1.optimizer_state_dict()
    assert len(state["state"]) > 20 and state["param_groups"][0]["lr"] == 2e-3
    some = next(iter(state["state"].values()))
    assert set(some) >= {"step", "exp_avg", "exp_avg_sq"} and float(some["step"]) == 2.0
    # resume in a fresh model/trainer from (model state_dict, optimizer state_dict)
    m2 = build(cfg, {k: v.detach().cpu() for k, v in m1.state_dict().items()})
    t2 = Trainer(m2, learning_rate=1e-3)
    t2.load_optimizer_state_dict(state)
    assert t2.opt_step == 2 and t2.lr == 2e-3
    l1 = float(t1.step(X, y))
    l2 = float(t2.step(X, y))
    assert abs(l1 - l2) <= 1e-6 * abs(l1)
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert rel(b.detach(), a.detach()) <= 1e-5, k
    # a torch AdamW can take the state over too (what the reference's load_checkpoint does)
    m1.configure_optimizers(0.1, 2e-3, (0.9, 0.95), "cuda").load_state_dict(t1.optimizer_state_dict())


def test_trainer_steps_in_the_original_vit_branch():
    """BASELINE config 4 through nvit_b200.Trainer (bench.py --variant orig): two steps against the oracle's step."""
    cfg = O.named_config("tiny", use_nvit=False)
    sd = O.init_state_dict(cfg, 7)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(16, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    ref = O.OracleTrainer({k: v.to(DEV) for k, v in sd.items()}, cfg, lr=1e-3)
    model = build(cfg, sd)
    tr = Trainer(model, learning_rate=1e-3)
    for it in range(2):
        ref_loss, _, _ = ref.step(X, y)
        loss = tr.step(X, y)
        assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss)), (it, float(loss), float(ref_loss))
    assert tr.last_aux == {}


def test_uint8_hwc_input_and_device_loader():
    """SURVEY.md 8f (input pipeline on device): raw uint8 HWC batches, ToTensor + Normalize(0.5, 0.5) folded into the im2col
    kernels (train.py:266-273, 1081-1092), staged through the double-buffered DeviceLoader."""
    from nvit_b200 import DeviceLoader
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 11)
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randint(0, 256, (8, 32, 32, 3), generator=g, dtype=torch.uint8), torch.randint(0, 10, (8,), generator=g))
               for _ in range(5)]
    m_u8, m_f = build(cfg, sd).eval(), build(cfg, sd).eval()
    seen = 0
    with torch.no_grad():
        for (Xd, yd), (Xh, yh) in zip(DeviceLoader(batches, DEV), batches):
            assert Xd.is_cuda and Xd.dtype == torch.uint8 and torch.equal(Xd.cpu(), Xh) and torch.equal(yd.cpu(), yh)
            lu, au = m_u8(Xd)
            Xf = ((Xh.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5          # torchvision ToTensor + Normalize(0.5, 0.5)
            lf, af = m_f(Xf.to(DEV))
            assert rel(lu, lf) <= 2e-3, rel(lu, lf)
            assert abs(float(au["reconstruction"]) - float(af["reconstruction"])) <= 2e-3 * float(af["reconstruction"])
            # and against the oracle on the normalised float image
            lo, _ = O.vit_forward({k: v.to(DEV) for k, v in sd.items()}, cfg, Xf.to(DEV))
            assert rel(lu, lo) <= 1e-2
            seen += 1
    assert seen == len(batches)
    # training step straight from uint8
    model = build(cfg, sd)
    tr = Trainer(model)
    Xd, yd = batches[0][0].to(DEV), batches[0][1].to(DEV)
    l0 = float(tr.step(Xd, yd))
    l1 = float(tr.step(Xd, yd))
    assert l1 < l0
    with pytest.raises(ValueError, match="HWC"):
        model(torch.zeros(2, 3, 32, 32, dtype=torch.uint8, device=DEV))
    # asynchronous consumption: training steps fed by the loader == the same steps fed by blocking copies
    many = [batches[i % len(batches)] for i in range(12)]
    ma, mb = build(cfg, sd), build(cfg, sd)
    ta, tb = Trainer(ma), Trainer(mb)
    for Xd, yd in DeviceLoader(many, DEV):
        la = ta.step(Xd, yd)
    for Xh, yh in many:
        lb = tb.step(Xh.to(DEV), yh.to(DEV))
    assert abs(float(la) - float(lb)) <= 2e-3 * abs(float(lb)), (float(la), float(lb))


def test_gradient_accumulation_matches_single_step():
    """train.py:898-933 runs `gradient_accumulation_steps` micro-steps on the SAME batch with loss / steps: the accumulated
    gradient equals the single-step one, so both trainers must take the same optimizer step (this also pins the
    overwrite semantics of dL/dsuv = rowdot(W_fc, dW_fc) / suv over accumulated weight gradients)."""
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 21)
    g = torch.Generator().manual_seed(8)
    X = torch.randn(16, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    m1, m2 = build(cfg, sd), build(cfg, sd)
    t1 = Trainer(m1, learning_rate=1e-3, grad_clip=0.0)
    t2 = Trainer(m2, learning_rate=1e-3, grad_clip=0.0, gradient_accumulation_steps=2)
    # compare the accumulated gradients before the optimizer consumes them
    for t in (t1, t2):
        t._ensure_state()
        t.loss_buf.zero_()
        t.engine.zero_grad()
        for k in range(t.grad_accum):
            t.micro_step(X, y, last=(k == t.grad_accum - 1))
    torch.cuda.synchronize()
    for n, _ in m1.named_parameters():
        s = m1.engine.slots[n]
        if s.off >= m1.engine.n_active:
            continue
        assert rel(m2.engine.g(n), m1.engine.g(n)) <= 5e-3, (n, rel(m2.engine.g(n), m1.engine.g(n)))
    assert abs(float(t2.loss_buf) / 2 - float(t1.loss_buf)) <= 1e-5 * abs(float(t1.loss_buf))     # one mean CE added per micro-step


def test_full_size_b16_batch256_sample_independence_linearity_and_step():
    """BASELINE config 2 at its FULL size (nViT-B/16, 224 px, batch 256, the bench workload), checked through
    size-independent properties, with the oracle anchoring one slice:
      * samples are independent (no batch statistics anywhere): the logits of the 256-image batch equal the
        logits of its 32-image slices run on their own;
      * the mean-CE gradient of the whole batch is the mean of the slices' gradients (linearity of the backward pass);
      * the first slice matches the fp32 oracle (logits and every gradient, the usual tolerance);
      * one full-size Trainer.step (graph-less) leaves every normalised weight row at unit norm and lowers the loss.
    """
    cfg = O.named_config("b16")
    sd = O.init_state_dict(cfg, 0)
    B, S = 256, 32
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(B, cfg.channels, cfg.image_size, cfg.image_size, generator=g).to(DEV)
    y = torch.randint(0, cfg.num_classes, (B,), generator=g).to(DEV)
    model = build(cfg, sd)
    logits, aux = model(X)
    assert logits.shape == (B, cfg.num_classes) and bool(torch.isfinite(logits).all())
    F.cross_entropy(logits, y).backward()
    full = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    full_logits = logits.detach().clone()
    del logits, aux
    acc = {n: torch.zeros_like(v, dtype=torch.float64) for n, v in full.items()}
    for s in range(0, B, S):
        model.zero_grad(set_to_none=True)
        ls, _ = model(X[s:s + S])
        # row results do not depend on which other rows share the launch: tiles never mix rows
        assert rel(ls.detach(), full_logits[s:s + S]) <= 1e-4, (s, rel(ls.detach(), full_logits[s:s + S]))
        F.cross_entropy(ls, y[s:s + S]).backward()
        for n, p in model.named_parameters():
            if p.grad is not None:
                acc[n] += p.grad.double()
        if s == 0:
            ref_logits, _, _, ref_grads = oracle_grads(cfg, sd, X[:S], y[:S])
            assert rel(ls.detach(), ref_logits) <= 1e-2
            check_grads(model, ref_grads)
            del ref_grads
    gnorm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in full.values())))
    for n, v in full.items():
        mean = acc[n] / (B // S)
        err = float((v.double() - mean).norm())
        # bf16 operands are rounded per element, not per batch: only fp32 / split-K summation order differs
        assert err <= 2e-2 * float(mean.norm()) + 1e-4 * gnorm, (n, err, float(mean.norm()))
    del full, acc
    model.zero_grad(set_to_none=True)
    tr = Trainer(model, learning_rate=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0)
    l0 = float(tr.step(X, y))
    l1 = float(tr.step(X, y))
    assert np.isfinite(l0) and l1 < l0, (l0, l1)
    params = dict(model.named_parameters())
    for i in range(cfg.n_layer):
        for nm, dim in O.NORMALIZED:
            w = params[f"transformer.h.{i}.{nm}.weight"].detach()
            assert float((w.norm(dim=dim) - 1).abs().max()) <= 1e-3, (i, nm)


def test_l16_shapes_match_oracle():
    """BASELINE config 3 shapes (nViT-L/16: C = 1024, 24 blocks, 16 heads) on a small batch against the fp32 oracle."""
    cfg = O.named_config("l16")
    sd = O.init_state_dict(cfg, 1)
    g = torch.Generator().manual_seed(4321)
    X = torch.randn(2, cfg.channels, cfg.image_size, cfg.image_size, generator=g).to(DEV)
    y = torch.randint(0, cfg.num_classes, (2,), generator=g).to(DEV)
    ref_logits, ref_recon, ref_loss, ref_grads = oracle_grads(cfg, sd, X, y)
    model = build(cfg, sd)
    logits, aux = model(X)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    assert rel(logits.detach(), ref_logits) <= 1e-2, rel(logits.detach(), ref_logits)
    assert abs(float(aux["reconstruction"]) - float(ref_recon)) <= 1e-2 * float(ref_recon)
    check_grads(model, ref_grads)


def _steps(trainer, X, y, n):
    return [float(trainer.step(X, y)) for _ in range(n)]


def _param_dist(ma, mb):
    worst = 0.0
    for (n, a), (_, b) in zip(ma.named_parameters(), mb.named_parameters()):
        worst = max(worst, rel(b.detach(), a.detach()))
    return worst


@pytest.mark.parametrize("fused", [True, False])
def test_cuda_graph_replay_equals_eager_steps(fused):
    """The bench's timed path: k steps of Trainer(cuda_graph=True) (2 eager warm-up steps, then capture + replays, learning
    rate and step count in device memory) against the same k eager steps from the same state, including a set_lr between
    replays, forwards at other batch sizes between replays (the captured activation set is pinned) and an optimizer-state
    reload (which drops and re-captures the graph).

    Two EAGER runs of the same steps are not bit-identical either (fp32 atomics: split-K reduce-adds, per-channel alpha / sqk
    / bias sums; AdamW's first steps turn that noise into +-lr on near-zero gradients), so the criterion is calibrated in
    place: a second eager trainer gives the noise floor (itself a noisy number: a handful of steps, measured between 1e-7 and
    3e-4), and graph-vs-eager must stay within max(10x floor, 5e-4 on losses / 5e-3 on parameters) and under absolute caps
    (loss 2e-3, parameters 2e-2) that a missed launch, a stale learning rate or a wrong step count break by an order of
    magnitude (MEASURED: deviations 2e-5 ... 4e-4)."""
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 13)
    g = torch.Generator().manual_seed(77)
    X = torch.randn(16, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    me, m2, mg = build(cfg, sd), build(cfg, sd), build(cfg, sd)
    te = Trainer(me, learning_rate=1e-3, fused_tail=fused)
    t2 = Trainer(m2, learning_rate=1e-3, fused_tail=fused)
    tg = Trainer(mg, learning_rate=1e-3, fused_tail=fused, cuda_graph=True, graph_warmup_steps=2)
    trainers = (te, t2, tg)

    def run(n):
        return [_steps(t, X, y, n) for t in trainers]

    def check(tag, losses):
        le, l2, lg = losses
        floor = max(abs(a - b) / abs(a) for a, b in zip(le, l2))
        dev = max(abs(a - b) / abs(a) for a, b in zip(le, lg))
        pfloor, pdev = _param_dist(me, m2), _param_dist(me, mg)
        print(f"[graph vs eager, {tag}] loss deviation {dev:.2e} (eager-vs-eager floor {floor:.2e}); "
              f"worst parameter rel-L2 {pdev:.2e} (floor {pfloor:.2e})")
        assert dev <= max(10 * floor, 5e-4) and dev <= 2e-3, (tag, dev, floor, le, lg)
        assert pdev <= max(10 * pfloor, 5e-3) and pdev <= 2e-2, (tag, pdev, pfloor)

    check("2 eager + capture + 3 replays", run(5))
    assert tg.replays == 3 and tg._graph is not None
    for t in trainers:
        t.set_lr(3e-4)
    check("set_lr between replays", run(2))
    assert tg.replays == 5
    assert te.opt_step == tg.opt_step == 7 and float(tg.hyper[1]) == 7.0 and abs(float(tg.hyper[0]) - 3e-4) < 1e-9
    # a forward at another batch size between replays must not disturb the captured activation set (it is pinned)
    mg.eval()
    with torch.no_grad():
        mg(X[:5])
        mg(X[:7])
        mg(X[:3])
    mg.train()
    assert 16 in mg.engine._acts and 16 in mg.engine._pinned
    check("after forwards at other batch sizes", run(2))
    assert tg.replays == 7
    # optimizer-state reload: the graph is dropped, re-captured, and the run continues like the eager ones
    state = te.optimizer_state_dict()
    msd = {k: v.detach().clone() for k, v in me.state_dict().items()}
    for m, t in ((me, te), (m2, t2), (mg, tg)):
        m.load_state_dict(msd)
        t.load_optimizer_state_dict(state)
    assert tg._graph is None
    check("after an optimizer-state reload", run(3))
    assert tg._graph is not None


def test_fused_optimizer_tail_equals_separate_kernels():
    """nvit_adamw_norm_fused (clip + AdamW + normalize_matrices + bf16 operands + zero_grad, one pass) against
    the separate sumsq / adamw_flat / weight_norm_multi / cast kernels ON IDENTICAL GRADIENTS AND STATE (two backward passes
    of the same batch differ in the last bits through fp32 atomics, which AdamW's first steps amplify): parameters,
    moments, bf16 operands, zeroed gradients and unit-norm rows / columns; also in the original-ViT branch (no
    normalisation) and with bias tensors."""
    for name, over in (("tiny", dict()), ("mini", dict(bias=True)), ("tiny", dict(use_nvit=False))):
        cfg = O.named_config(name, **over)
        sd = O.init_state_dict(cfg, 5)
        g = torch.Generator().manual_seed(3)
        X = torch.randn(8, 3, cfg.image_size, cfg.image_size, generator=g).to(DEV)
        y = torch.randint(0, cfg.num_classes, (8,), generator=g).to(DEV)
        ma, mb = build(cfg, sd), build(cfg, sd)
        ta, tb = Trainer(ma, fused_tail=True), Trainer(mb, fused_tail=False)
        ea, eb = ma.engine, mb.engine
        for it in range(3):
            ta._ensure_state()
            tb._ensure_state()
            # same parameters, moments and gradients on both sides, then one optimizer tail each
            eb.P32.copy_(ea.P32)
            tb.m.copy_(ta.m)
            tb.v.copy_(ta.v)
            ta.loss_buf.zero_()
            ta.micro_step(X, y)
            eb.G32.copy_(ea.G32)
            ta.optimizer_step()
            tb.optimizer_step()
            torch.cuda.synchronize()
            na = ea.n_active
            assert rel(ea.P32[:na], eb.P32[:na]) <= 1e-6, (name, it, rel(ea.P32[:na], eb.P32[:na]))
            assert rel(ta.m[:na], tb.m[:na]) <= 1e-6 and rel(ta.v[:na], tb.v[:na]) <= 1e-6, (name, it)
            assert float(ea.G32[:na].abs().max()) == 0.0 and float(eb.G32[:na].abs().max()) == 0.0      # zero_grad is part of the pass
            # the bf16 operands the fused pass wrote are exactly the rounding of its fp32 values
            assert torch.equal(ea.W16[:ea.n_gemm], ea.P32[:ea.n_gemm].to(torch.bfloat16)), (name, it)
            assert ta.opt_step == tb.opt_step == it + 1
        if cfg.use_nvit:
            params = dict(ma.named_parameters())
            for i in range(cfg.n_layer):
                for nm, dim in O.NORMALIZED:
                    w = params[f"transformer.h.{i}.{nm}.weight"].detach()
                    assert float((w.norm(dim=dim) - 1).abs().max()) <= 1e-5, (i, nm)
        # untrained (grad-less) tensors are not touched by the tail
        assert torch.equal(dict(ma.named_parameters())["reconstruction_head.0.weight"].detach().cpu(), sd["reconstruction_head.0.weight"])
        # and whole steps of the two trainers stay together (up to the atomics noise of two separate backward passes)
        la, lb = float(ta.step(X, y)), float(tb.step(X, y))
        assert abs(la - lb) <= 2e-3 * abs(lb), (name, la, lb)


def test_weights_written_behind_the_engines_back_are_seen():
    """The reference's normalize_matrices writes through `W.data.copy_` (train.py:474-480), which does not bump the tensor
    version: the next forward must still use the new weights (the bf16 operands are rebuilt on every forward that does
    not directly follow the Trainer's own fused tail)."""
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 9)
    g = torch.Generator().manual_seed(21)
    X = torch.randn(4, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (4,), generator=g).to(DEV)
    model = build(cfg, sd)
    tr = Trainer(model)
    tr.step(X, y)
    model.eval()
    with torch.no_grad():
        l0, _ = model(X)
        w = dict(model.named_parameters())["transformer.h.0.c_fc.weight"]
        v0 = w._version
        new = torch.randn(w.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
        new = new / new.norm(dim=1, keepdim=True)
        w.data.copy_(new)                                # exactly what the reference's normalize_matrices does
        assert w._version == v0                          # ... and it is invisible to version counters
        l1, _ = model(X)
        ref = {k: v.detach().clone() for k, v in model.state_dict().items()}
        lo, _ = O.vit_forward(ref, cfg, X)
    assert rel(l1, l0) > 1e-3                            # the new weights were used
    assert rel(l1, lo) <= 1e-2                           # and the result is the oracle's for the new weights
    # the same between two Trainer steps, announced through invalidate_operands()
    model.train()
    w.data.copy_(sd["transformer.h.0.c_fc.weight"].to(DEV))
    model.engine.invalidate_operands()
    l2 = float(tr.step(X, y))
    assert np.isfinite(l2)


def test_input_shape_and_label_validation():
    """a wrongly sized image or label tensor must raise instead of writing past the activation buffers / reading
    out-of-range classes; int32 labels are converted; an out-of-range label is ignored (zero loss / gradient row)."""
    from nvit_b200 import ops
    cfg = O.named_config("tiny")
    model = build(cfg, O.init_state_dict(cfg, 1))
    with pytest.raises(ValueError, match="input must be"):
        model(torch.zeros(2, 3, 40, 40, device=DEV))
    with pytest.raises(ValueError, match="input must be"):
        model(torch.zeros(2, 1, 32, 32, device=DEV))
    tr = Trainer(model)
    X = torch.randn(4, 3, 32, 32, device=DEV)
    with pytest.raises(ValueError, match="labels"):
        tr.step(X, torch.zeros(3, dtype=torch.int64, device=DEV))
    l32 = float(tr.step(X, torch.tensor([1, 2, 3, 4], dtype=torch.int32, device=DEV)))
    assert np.isfinite(l32)
    logits = torch.randn(4, 10, device=DEV)
    tgt = torch.tensor([1, -100, 3, 12], device=DEV)
    loss = torch.zeros(1, device=DEV)
    dl = torch.empty_like(logits)
    ops.cross_entropy(logits, tgt, loss, dl)
    ref = (F.cross_entropy(logits[[0, 2]], tgt[[0, 2]], reduction="sum") / 4)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert float(dl[1].abs().max()) == 0.0 and float(dl[3].abs().max()) == 0.0
    with pytest.raises(TypeError):
        ops.cross_entropy(logits, tgt.int(), loss, dl)


def test_reconstruction_gradient_request_fails_loudly_and_state_survives_a_noop_move():
    """without Kohonen maps the reconstruction loss has no backward here - asking for one must raise, not
    return zeros; and a no-op model.to(device) must not reset the engine or the Trainer's AdamW moments."""
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 2)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(4, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (4,), generator=g).to(DEV)
    model = build(cfg, sd)
    logits, aux = model(X)
    with pytest.raises(RuntimeError, match="reconstruction"):
        (F.cross_entropy(logits, y) + 0.1 * aux["reconstruction"]).backward()
    model.zero_grad(set_to_none=True)
    logits, aux = model(X)
    F.cross_entropy(logits, y).backward()                  # the supported objective still works
    assert dict(model.named_parameters())["transformer.h.0.c_fc.weight"].grad is not None
    model.zero_grad(set_to_none=True)
    tr = Trainer(model)
    tr.step(X, y)
    tr.step(X, y)
    P32, m_before = model.engine.P32, tr.m.clone()
    model.to(DEV)                                          # no parameter moves: nothing may be rebuilt
    assert model.engine.P32 is P32
    tr.step(X, y)
    assert tr._state_for is P32 and float((tr.m - m_before).abs().max()) > 0 and tr.opt_step == 3


def test_tensor_maps_are_encoded_once_per_buffer():
    """TMA tensor maps are a pure function of (base, type, shape, pitch, box); the engine's buffers are static, so after the
    first eager step no further map is encoded (nvit_tmap_cache_stats) and the steps that run on cached maps give the same
    losses as a fresh model whose first step encodes its own."""
    from nvit_b200 import _lib
    cfg = O.named_config("tiny")
    sd = O.init_state_dict(cfg, 5)
    g = torch.Generator().manual_seed(77)
    X = torch.randn(16, 3, 32, 32, generator=g).to(DEV)
    y = torch.randint(0, 10, (16,), generator=g).to(DEV)
    tr = Trainer(build(cfg, sd), learning_rate=1e-3)
    first = [float(tr.step(X, y)) for _ in range(2)]
    enc0, hit0 = _lib.tmap_cache_stats()
    more = [float(tr.step(X, y)) for _ in range(3)]
    enc1, hit1 = _lib.tmap_cache_stats()
    assert enc1 == enc0, f"{enc1 - enc0} tensor maps re-encoded over three steady-state steps"
    assert hit1 > hit0, (enc0, hit0, enc1, hit1)
    tr2 = Trainer(build(cfg, sd), learning_rate=1e-3)       # new buffers: new maps (or recycled addresses: identical maps)
    again = [float(tr2.step(X, y)) for _ in range(5)]
    for a, b in zip(again, first + more):          # split-K partial sums are added in arrival order: not bit-reproducible
        assert abs(a - b) <= 1e-3 * abs(b), (again, first + more)
