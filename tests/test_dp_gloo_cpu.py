"""Data-parallel host logic on CPU with gloo, world_size 2.

(1) GradReducer: ranges of one flat gradient buffer handed over in backward order (large block slices immediately,
    small ranges coalesced) come out as the SUM over ranks, every element reduced exactly once.
(2) The multi-GPU oracle of the path: a 2-rank sharded step == the 1-rank step on the concatenated batch, using the
    fp32 oracle forward/backward for the arithmetic and GradReducer for the exchange (the CUDA engine hands GradReducer
    the same ranges on the GPU; the reference's own DDP path never reduces).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from nvit_b200 import ViT, ViTConfig
from nvit_b200.engine import Engine
from nvit_b200.train import GradReducer
from oracle import nvit_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)


def _reducer_worker(rank, world, port, q):
    _init(rank, world, port)
    n = 5_000
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    red = GradReducer(flat, min_bucket=1000)
    # backward order: last block first, then small tails, with a gap that is never handed over
    for lo, hi in [(3000, 4200), (1500, 3000), (300, 1500), (0, 100), (100, 300), (4200, 4300), (4300, 4900)]:
        red.ready(lo, hi)
    red.finish()
    others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    expect = sum(others)
    ok = torch.allclose(flat[:4900], expect[:4900], atol=1e-6) and torch.equal(flat[4900:], mine[4900:])
    # the same hand-over with adjacent large ranges merged into buckets of >= 2500 elements: fewer collectives, same sums
    flat2 = mine.clone()
    red2 = GradReducer(flat2, min_bucket=1000, bucket_elems=2500)
    issued = []
    orig_issue = red2._issue
    red2._issue = lambda lo, hi: (issued.append((lo, hi)), orig_issue(lo, hi))[1]
    for lo, hi in [(3000, 4200), (1500, 3000), (300, 1500), (0, 100), (100, 300), (4200, 4300), (4300, 4900)]:
        red2.ready(lo, hi)
    red2.finish()
    ok2 = torch.allclose(flat2[:4900], expect[:4900], atol=1e-6) and torch.equal(flat2[4900:], mine[4900:])
    ok2 = ok2 and issued[0] == (1500, 4200) and red2.reduced_elems == red.reduced_elems
    q.put((rank, bool(ok and ok2), red.reduced_elems))
    dist.destroy_process_group()


def _dp_worker(rank, world, port, q):
    _init(rank, world, port)
    cfg = O.named_config("micro")
    sd = O.formula_state_dict(cfg)
    X, y = O.formula_batch(cfg, 4)
    # engine layout (host-side only) gives the flat offsets the CUDA path reduces over
    eng = Engine(ViT(ViTConfig(**cfg.as_dict())))
    eng._build_layout()
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sl = slice(rank * 2, rank * 2 + 2)
    logits, _ = O.vit_forward(p, cfg, X[sl])
    (F.cross_entropy(logits, y[sl]) / world).backward()       # 1/world folded into the loss scale, as Trainer does
    flat = torch.zeros(eng.n_total)
    for n, s in eng.slots.items():
        if p[n].grad is not None:
            flat[s.off:s.off + s.numel] = p[n].grad.flatten()
    red = GradReducer(flat, min_bucket=1 << 12)
    for i in reversed(range(cfg.n_layer)):
        red.ready(*eng.block_grad_range(i))
    red.ready(0, eng.block_grad_range(0)[0])
    red.ready(eng.block_grad_range(cfg.n_layer - 1)[1], eng.n_active)
    red.finish()
    # single-rank reference on the concatenated batch
    pr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lg, _ = O.vit_forward(pr, cfg, X)
    F.cross_entropy(lg, y).backward()
    worst = 0.0
    for n, s in eng.slots.items():
        if pr[n].grad is None:
            continue
        got = flat[s.off:s.off + s.numel]
        ref = pr[n].grad.flatten()
        worst = max(worst, float((got - ref).norm() / (ref.norm() + 1e-12)))
    q.put((rank, worst < 1e-4, worst))
    dist.destroy_process_group()


def _run(worker):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.timeout(300)
def test_grad_reducer_sums_each_range_once_world2():
    out = _run(_reducer_worker)
    assert all(ok for _, ok, _ in out), out
    assert all(n == 4900 for _, _, n in out), out


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_equals_single_rank_on_concatenated_batch():
    out = _run(_dp_worker)
    assert all(ok for _, ok, _ in out), out
