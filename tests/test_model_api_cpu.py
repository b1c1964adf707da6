"""Drop-in API checks that need no GPU: config fields, module/attribute names, state_dict layout and initial
distributions of nvit_b200.model against the live reference (skipped where /root/reference is absent) and against the
oracle's parameter table; the engine's flat parameter layout; loud failure without CUDA."""
import dataclasses

import pytest
import torch

from nvit_b200 import ViT, ViTConfig, Block, CrossAttentionBlock, RMSNorm, justnorm
from oracle import nvit_oracle as O


def test_config_fields_match_oracle_table():
    mine = {f.name: f.default for f in dataclasses.fields(ViTConfig)}
    theirs = {f.name: f.default for f in dataclasses.fields(O.OracleConfig)}
    assert mine == theirs


def test_config_fields_match_reference(reference_model_module):
    ref = reference_model_module
    mine = [(f.name, f.type if isinstance(f.type, str) else f.type.__name__, f.default) for f in dataclasses.fields(ViTConfig)]
    theirs = [(f.name, f.type if isinstance(f.type, str) else f.type.__name__, f.default) for f in dataclasses.fields(ref.ViTConfig)]
    assert mine == theirs


@pytest.mark.parametrize("name,over", [("micro", {}), ("micro", {"bias": True}), ("tiny", {}), ("mini", {"base_scale": 1 / 32}),
                                       ("micro", {"use_nvit": False}), ("tiny", {"use_nvit": False, "bias": True}),
                                       ("micro", {"use_kohonen": True, "kohonen_nodes": 32}), ("tiny", {"use_kohonen": True})])
def test_state_dict_keys_shapes_match_oracle(name, over):
    cfg = O.named_config(name, **over)
    m = ViT(ViTConfig(**cfg.as_dict()))
    sd = m.state_dict()
    shapes = dict(O.param_shapes(cfg))
    buffers = O.kohonen_buffers(cfg) if cfg.use_kohonen else {}
    assert set(sd) == set(shapes) | set(buffers)
    for k, v in sd.items():
        if k in buffers:
            assert torch.equal(v, buffers[k]) and v.dtype == torch.int64, k
        else:
            assert tuple(v.shape) == tuple(shapes[k]) and v.dtype == torch.float32, k


@pytest.mark.parametrize("name,over", [("micro", {}), ("tiny", {"bias": True}),
                                       ("micro", {"use_kohonen": True, "kohonen_nodes": 32, "kohonen_scheduler_enabled": True})])
def test_same_seed_gives_the_reference_initialisation(reference_model_module, name, over):
    """Modules are created in the reference's order, so the RNG stream — and every initial value — is identical."""
    ref = reference_model_module
    cfg = O.named_config(name, **over)
    torch.manual_seed(0)
    theirs = ref.ViT(ref.ViTConfig(**cfg.as_dict())).state_dict()
    torch.manual_seed(0)
    mine = ViT(ViTConfig(**cfg.as_dict())).state_dict()
    assert list(mine.keys()) == list(theirs.keys())
    for k in theirs:
        assert torch.equal(mine[k], theirs[k]), k


def test_attribute_surface_used_by_reference_trainer():
    cfg = O.named_config("micro")
    m = ViT(ViTConfig(**cfg.as_dict()))
    blk = m.transformer.h[0]
    assert isinstance(blk, Block) and isinstance(m.cross_attention, CrossAttentionBlock) and isinstance(blk.rmsnorm_att, RMSNorm)
    for attr in ("sqk", "sqk_init_value", "sqk_init_scaling", "attn_alpha", "attn_alpha_init_value", "attn_alpha_init_scaling",
                 "mlp_alpha", "suv", "suv_init_value", "suv_init_scaling", "skip_param", "query", "key", "value", "att_c_proj",
                 "c_fc", "mlp_c_proj"):
        assert hasattr(blk, attr), attr
    for attr in ("sz", "step", "total_steps", "config", "local_patch_embed", "global_patch_embed", "local_pos_embed",
                 "global_pos_embed", "reconstruction_head", "mlp_head", "configure_optimizers", "estimate_mfu", "num_params"):
        assert hasattr(m, attr), attr
    opt = m.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu")
    groups = opt.param_groups
    assert len(groups) == 3 and groups[0]["weight_decay"] == 0.1 and groups[1]["weight_decay"] == 0.0
    assert len(groups[2]["params"]) == 1 and groups[2]["params"][0] is m.sz
    x = torch.randn(3, 5)
    assert torch.allclose(justnorm(x).norm(dim=-1), torch.ones(3))


def test_engine_flat_layout_covers_every_parameter_once():
    from nvit_b200.engine import Engine
    cfg = O.named_config("mini", bias=True)
    m = ViT(ViTConfig(**cfg.as_dict()))
    eng = Engine(m)
    eng._build_layout()
    named = dict(m.named_parameters())
    assert set(eng.order) == set(named) and len(eng.order) == len(named)
    spans = sorted((s.off, s.off + s.numel) for s in eng.slots.values())
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= eng.n_total
    assert 0 < eng.n_gemm <= eng.n_decay <= eng.n_active <= eng.n_total
    for n, s in eng.slots.items():
        decays = named[n].dim() >= 2 and "sz" not in n
        inactive = ".rmsnorm_" in n or n.startswith("reconstruction_head.")
        if inactive:
            assert s.off >= eng.n_active, n
        elif decays:
            assert s.off + s.numel <= eng.n_decay, n
        else:
            assert eng.n_decay <= s.off and s.off + s.numel <= eng.n_active, n
        assert s.off % 8 == 0
    C = cfg.n_embd
    b = "transformer.h.1."
    assert eng.slots[b + "key.weight"].off == eng.slots[b + "query.weight"].off + C * C
    assert eng.slots[b + "value.weight"].off == eng.slots[b + "query.weight"].off + 2 * C * C
    assert eng.slots[b + "key.bias"].off == eng.slots[b + "query.bias"].off + C       # C % 8 == 0
    lo, hi = eng.block_grad_range(1)
    assert hi - lo == 16 * C * C


def test_no_cpu_fallback():
    cfg = O.named_config("micro")
    m = ViT(ViTConfig(**cfg.as_dict()))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 16, 16))
    with pytest.raises(ValueError, match="head_dim"):
        ViT(ViTConfig(n_embd=96, n_head=2, use_nvit=True))


def test_kohonen_surface_matches_reference(reference_model_module):
    """BASELINE config 5 API: map attributes, learning-rate schedule, optimizer groups, layout of the flat buffers."""
    ref = reference_model_module
    cfg = O.named_config("micro", use_kohonen=True, kohonen_nodes=32, kohonen_scheduler_enabled=True,
                         kohonen_scheduler_warmup_steps=5, kohonen_scheduler_decay_steps=20, kohonen_alpha=0.3)
    mine, theirs = ViT(ViTConfig(**cfg.as_dict())), ref.ViT(ref.ViTConfig(**cfg.as_dict()))
    for attr in ("m", "n", "grid_size", "input_dim", "alpha", "sigma", "periodic"):
        assert getattr(mine.local_kohonen, attr) == getattr(theirs.local_kohonen, attr), attr
    for step in (0, 1, 4, 5, 6, 12, 20, 21, 1000):
        assert mine.get_kohonen_lr(step) == theirs.get_kohonen_lr(step) == O.kohonen_lr(cfg, step)
    a, b = torch.randn(2, 3, 8), torch.randn(2, 3, 8)
    assert torch.equal(mine.combine_representations(a, b), theirs.combine_representations(a, b))
    go, gr = (m.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu").param_groups for m in (mine, theirs))
    assert [len(g["params"]) for g in go] == [len(g["params"]) for g in gr]
    from nvit_b200.engine import Engine
    eng = Engine(mine)
    eng._build_layout()
    s = eng.slots
    assert eng.rec_active and s["reconstruction_head.0.weight"].off < eng.n_gemm
    assert s["local_kohonen.nodes"].off < eng.n_decay and s["reconstruction_head.0.bias"].off < eng.n_active
    assert s["map_balance"].off >= eng.n_active          # never receives a gradient
    with pytest.raises(RuntimeError, match="CUDA"):
        mine.local_kohonen(torch.zeros(4, 64))


@pytest.mark.parametrize("over", [{}, {"use_kohonen": True, "kohonen_nodes": 32}, {"use_nvit": False}])
def test_checkpoint_interchange_with_the_reference(reference_model_module, tmp_path, over):
    """a checkpoint written the reference's way (train.py:629-655: model / optimizer / model_args) loads
    into this package the reference's way (train.py:375-381), and the other way round."""
    import dataclasses
    import torch.nn.functional as F
    ref = reference_model_module
    cfg = O.named_config("micro", **over)
    torch.manual_seed(1)
    theirs = ref.ViT(ref.ViTConfig(**cfg.as_dict()))
    if not cfg.use_nvit:
        for blk in theirs.transformer.h:
            blk.rmsnorm_att, blk.rmsnorm_mlp = ref.RMSNorm(cfg.n_embd), ref.RMSNorm(cfg.n_embd)
    opt = theirs.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu")
    logits, _ = theirs(torch.randn(2, 3, cfg.image_size, cfg.image_size))
    F.cross_entropy(logits, torch.tensor([1, 2])).backward()
    opt.step()                                        # non-trivial optimizer state
    path = tmp_path / "checkpoint_latest.pt"
    torch.save({"model": theirs.state_dict(), "optimizer": opt.state_dict(), "model_args": dataclasses.asdict(theirs.config),
                "iter_num": 7}, path)
    ck = torch.load(path, map_location="cpu")
    mine = ViT(ViTConfig(**ck["model_args"]))
    mine.load_state_dict(ck["model"])                 # strict
    my_opt = mine.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu")
    my_opt.load_state_dict(ck["optimizer"])
    named_theirs, named_mine = dict(theirs.named_parameters()), dict(mine.named_parameters())
    assert list(named_theirs) == list(named_mine)
    n_state = 0
    for k, prm in named_theirs.items():
        assert torch.equal(prm, named_mine[k]), k
        if prm in opt.state:
            assert torch.equal(opt.state[prm]["exp_avg"], my_opt.state[named_mine[k]]["exp_avg"]), k
            n_state += 1
    assert n_state > 20
    # and back: what this package saves loads into the reference
    path2 = tmp_path / "checkpoint_b200.pt"
    torch.save({"model": mine.state_dict(), "optimizer": my_opt.state_dict(), "model_args": dataclasses.asdict(mine.config)}, path2)
    ck2 = torch.load(path2, map_location="cpu")
    back = ref.ViT(ref.ViTConfig(**ck2["model_args"]))
    if not cfg.use_nvit:
        for blk in back.transformer.h:
            blk.rmsnorm_att, blk.rmsnorm_mlp = ref.RMSNorm(cfg.n_embd), ref.RMSNorm(cfg.n_embd)
    back.load_state_dict(ck2["model"])
    back.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cpu").load_state_dict(ck2["optimizer"])


def test_unsupported_shapes_are_rejected_at_construction():
    """Limits of the kernels (one sequence of <= 256 tokens per attention CTA, head_dim 64, rows of <= 1024 channels) raise
    when the model is built, not as a kernel error in the middle of a step."""
    from nvit_b200 import ViT, ViTConfig
    base = O.named_config("tiny").as_dict()
    with pytest.raises(ValueError, match="256 tokens"):
        ViT(ViTConfig(**dict(base, image_size=136, local_patch_size=8, global_patch_size=16)))     # 17 x 17 = 289 tokens
    with pytest.raises(ValueError, match="head_dim"):
        ViT(ViTConfig(**dict(base, n_embd=192, n_head=4)))
    with pytest.raises(ValueError, match="at most 1024"):
        ViT(ViTConfig(**dict(base, n_embd=2048, n_head=32)))
    ViT(ViTConfig(**dict(base, image_size=64, local_patch_size=4, global_patch_size=8)))             # 256 tokens: the limit
