"""Host half and oracle of the device-side AutoAugment (reference: train.py:1081-1092).

kornia (the reference's augmentation library) is absent from the image; the published uint8 definitions of the same
AutoAugment operations are taken from torchvision 0.26, run LIVE here where it is importable and through the committed
fixtures tests/golden/augment_golden.npz (tests/golden/make_augment_golden.py) everywhere.  Bar: bit-exact for the ten pixel
operations and for integer translations; rotations / shears may differ from grid_sample only where a source coordinate
lies ON a pixel boundary (checked pixel by pixel: every differing pixel is such a tie — a shear of exactly 0.2 puts every fifth row on one; the
measured rate is printed).
"""
import os

import numpy as np
import pytest

from nvit_b200 import augment as A
from oracle import augment_oracle as AO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "augment_golden.npz")
ALL_OPS = ["ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "Brightness", "Color", "Contrast", "Sharpness", "Posterize",
           "Solarize", "AutoContrast", "Equalize", "Invert"]
GEOMETRIC = {"ShearX", "ShearY", "Rotate"}


def images(S, n, seed):
    """Smooth + noisy uint8 images (flat regions, gradients, a saturated patch) so that every operation has work to do."""
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    out = []
    for i in range(n):
        base = 127 + 90 * np.sin(xs / (3.0 + i) + i) * np.cos(ys / (2.0 + i))
        img = np.stack([base + rng.normal(0, 20, (S, S)), base * 0.7 + 30 + rng.normal(0, 5, (S, S)),
                        255 - base + rng.normal(0, 40, (S, S))], -1)
        img = np.clip(img, 10 * (i % 3), 255 - 7 * (i % 4))
        img[: S // 4, : S // 4] = 255 if i % 2 else 0
        out.append(img.astype(np.uint8))
    out.append(np.full((S, S, 3), 77, np.uint8))               # constant image: autocontrast / equalize degenerate cases
    return np.stack(out)


def tv():
    return pytest.importorskip("torchvision.transforms.autoaugment")


def affine_mismatch_is_ties_only(mine, ref, p, S):
    """Rotations / shears: the oracle may differ from grid_sample only where the source coordinate lies on a pixel boundary
    (|frac - 0.5| < 1e-3 in x or y: the two float evaluations round such a tie differently).  Returns the mismatch rate."""
    bad = (mine != ref).any(-1)
    if bad.any():
        c = 0.5 * (S - 1)
        ys, xs = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(S, dtype=np.float64), indexing="ij")
        sx = p[0] * (xs - c) + p[1] * (ys - c) + p[2]
        sy = p[3] * (xs - c) + p[4] * (ys - c) + p[5]
        tie = (np.abs(sx - np.floor(sx) - 0.5) < 1e-3) | (np.abs(sy - np.floor(sy) - 0.5) < 1e-3)
        assert tie[bad].all(), f"{int((bad & ~tie).sum())} pixels differ away from a pixel boundary"
    return float(bad.mean())


def test_policy_tables_and_magnitudes_equal_torchvisions():
    aa = tv()
    import torchvision.transforms as T
    t = T.AutoAugment()
    for name, pol in (("imagenet", aa.AutoAugmentPolicy.IMAGENET), ("cifar10", aa.AutoAugmentPolicy.CIFAR10), ("svhn", aa.AutoAugmentPolicy.SVHN)):
        want = t._get_policies(pol)
        got = A.policies(name)
        assert len(got) == len(want) == 25
        for (g1, g2), (w1, w2) in zip(got, want):
            assert g1 == tuple(w1) and g2 == tuple(w2), (name, g1, g2, w1, w2)
    for S in (32, 224):
        space = t._augmentation_space(10, (S, S))
        for op, (mags, signed) in space.items():
            assert signed == (op in A._SIGNED)
            if mags.ndim == 0:
                continue
            for b in range(10):
                assert A.magnitude(op, b, S) == float(mags[b].item()), (op, b, S, A.magnitude(op, b, S), float(mags[b].item()))
    assert A.policies("cifar") == A.policies("cifar10")
    with pytest.raises(ValueError):
        A.policies("mnist")


def _torchvision_apply(img_hwc, op, mag):
    import torch
    from torchvision.transforms import InterpolationMode
    aa = tv()
    x = torch.from_numpy(img_hwc).permute(2, 0, 1).contiguous()
    y = aa._apply_op(x, op, mag, interpolation=InterpolationMode.NEAREST, fill=None)
    return y.permute(1, 2, 0).contiguous().numpy()


@pytest.mark.parametrize("S", [32, 57])
def test_oracle_matches_torchvision_live(S):
    tv()
    X = images(S, 4, seed=S)
    worst = 0.0
    for op in ALL_OPS:
        bins = [None] if op in ("AutoContrast", "Equalize", "Invert") else [0, 3, 7, 9]
        for b in bins:
            for sign in ((1, -1) if op in A._SIGNED else (1,)):
                mag = sign * A.magnitude(op, b, S) if b is not None else 0.0
                code, p = A.encode_op(op, mag, S)
                for i, img in enumerate(X):
                    mine = AO.apply_op(img, code, np.asarray(p, np.float32))
                    ref = _torchvision_apply(img, op, mag)
                    if op in GEOMETRIC:
                        rate = affine_mismatch_is_ties_only(mine, ref, p, S)
                        worst = max(worst, rate)
                    else:
                        assert np.array_equal(mine, ref), (op, b, sign, i, int((mine != ref).sum()))
    print(f"[augment oracle vs torchvision] S = {S}: worst pixel mismatch rate of a rotation / shear {worst:.4%}")


def test_oracle_matches_committed_torchvision_fixtures():
    g = np.load(GOLDEN, allow_pickle=False)
    X = g["images"]
    S = X.shape[1]
    names = [str(s) for s in g["op_names"]]
    for k, (op, mag, idx) in enumerate(zip(names, g["magnitudes"], g["image_index"])):
        code, p = A.encode_op(op, float(mag), S)
        mine = AO.apply_op(X[idx], code, np.asarray(p, np.float32))
        ref = g["outputs"][k]
        if op in GEOMETRIC:
            affine_mismatch_is_ties_only(mine, ref, p, S)
        else:
            assert np.array_equal(mine, ref), (op, float(mag), int((mine != ref).sum()))


def test_sampler_follows_the_policy_and_is_reproducible():
    aug = A.AutoAugment("cifar10", seed=7)
    ops, params = aug.plan(4096, 32)
    assert ops.shape == (4096, 2) and ops.dtype == np.int32 and params.shape == (4096, 2, 8) and params.dtype == np.float32
    ops2, params2 = A.AutoAugment("cifar10", seed=7).plan(4096, 32)
    assert np.array_equal(ops, ops2) and np.array_equal(params, params2)
    ops3, _ = A.AutoAugment("cifar10", seed=7, rank=1).plan(4096, 32)
    assert not np.array_equal(ops, ops3)
    # expected firing rate of the first / second operation = mean probability over the 25 sub-policies
    pol = A.policies("cifar10")
    for j in range(2):
        want = np.mean([sub[j][1] for sub in pol])
        got = float((ops[:, j] != A.IDENTITY).mean())
        assert abs(got - want) < 0.03, (j, got, want)
    # every code the policy can produce appears, none other
    codes = set(np.unique(ops))
    assert codes <= set(range(11)) and {A.AFFINE, A.EQUALIZE, A.AUTOCONTRAST, A.SHARPNESS, A.COLOR, A.BRIGHTNESS} <= codes
    # identity stages carry zero parameters; blend operations carry ratio and 1 - ratio
    assert not params[ops == A.IDENTITY].any()
    blend = np.isin(ops, [A.BRIGHTNESS, A.COLOR, A.CONTRAST, A.SHARPNESS])
    assert np.allclose(params[blend][:, 0] + params[blend][:, 1], 1.0, atol=1e-6)


def test_oracle_plan_composes_two_operations_and_identity_is_a_copy():
    X = images(16, 3, seed=1)
    aug = A.AutoAugment("svhn", seed=3)
    ops, params = aug.plan(len(X), 16)
    out = AO.apply_plan(X, ops, params)
    for b in range(len(X)):
        step = AO.apply_op(AO.apply_op(X[b], ops[b, 0], params[b, 0]), ops[b, 1], params[b, 1])
        assert np.array_equal(out[b], step)
    z = np.zeros_like(ops)
    assert np.array_equal(AO.apply_plan(X, z, np.zeros_like(params)), X)
    train_tf, val_tf = A.get_transforms("cifar10", seed=0)
    assert isinstance(train_tf, A.AutoAugment) and val_tf(X) is X


def test_integer_form_of_the_smoothing_division_used_by_the_kernel():
    """augment.cu rounds the 3 x 3 sum as (2 sum + 13) / 26 in integers; the oracle (and torchvision) round the float32 quotient
    sum / 13 half-to-even.  Identical for every reachable sum."""
    s = np.arange(0, 13 * 255 + 1)
    assert np.array_equal(np.rint((s.astype(np.float32) / np.float32(13)).astype(np.float32)).astype(np.int64), (2 * s + 13) // 26)


def test_the_public_call_refuses_host_tensors():
    import torch
    aug = A.AutoAugment("cifar10")
    with pytest.raises(RuntimeError, match="device only"):
        aug(torch.zeros(2, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(ValueError, match="HWC"):
        aug(torch.zeros(2, 3, 8, 8))


def test_operation_codes_agree_between_kernel_host_oracle_and_header():
    """One numbering in four places: the enum of csrc/augment.cu, nvit_b200/augment.py, oracle/augment_oracle.py and the text of
    include/nvit_b200.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cu = open(os.path.join(root, "nvit_b200", "csrc", "augment.cu")).read()
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"AUG_([A-Z]+) = (\d+)", cu))
    names = ["IDENTITY", "AFFINE", "BRIGHTNESS", "COLOR", "CONTRAST", "SHARPNESS", "POSTERIZE", "SOLARIZE", "AUTOCONTRAST", "EQUALIZE", "INVERT"]
    for n in names:
        assert enum[n] == getattr(A, n) == getattr(AO, n), n
    assert enum["NOPS"] == len(names) and int(re.search(r"AUG_NPARAM = (\d+)", cu).group(1)) == A.NPARAM
    header = open(os.path.join(root, "include", "nvit_b200.h")).read()
    doc = header[header.index("AutoAugment on the device"):header.index("int nvit_augment_u8")]
    for n in names:
        assert re.search(rf"\b{getattr(A, n)} {n.lower()}\b", doc), f"header does not document code {getattr(A, n)} as {n.lower()}"


def test_oracle_properties():
    """Size-independent properties of the operations (the GPU suite checks the same ones on the full-size batch)."""
    X = images(48, 3, seed=4)
    z = np.zeros(8, np.float32)
    for img in X:
        inv = AO.apply_op(img, AO.INVERT, z)
        assert np.array_equal(AO.apply_op(inv, AO.INVERT, z), img)
        for bits in range(0, 9):
            c, p = A.encode_op("Posterize", float(bits), 48)
            once = AO.apply_op(img, c, p)
            assert np.array_equal(AO.apply_op(once, c, p), once)                      # idempotent
            assert np.array_equal(once, img & np.uint8(256 - (1 << (8 - bits)) & 255))
        c, p = A.encode_op("Solarize", 0.0, 48)
        assert np.array_equal(AO.apply_op(img, c, p), inv)                            # threshold 0 inverts everything
        c, p = A.encode_op("Solarize", 255.5, 48)
        assert np.array_equal(AO.apply_op(img, c, p), img)
        for op in ("Brightness", "Color", "Contrast", "Sharpness"):                   # magnitude 0 = ratio 1 = identity
            c, p = A.encode_op(op, 0.0, 48)
            assert np.array_equal(AO.apply_op(img, c, p), img), op
        ac = AO.apply_op(img, AO.AUTOCONTRAST, z)
        for ch in range(3):
            if img[..., ch].max() > img[..., ch].min():
                assert ac[..., ch].min() == 0 and ac[..., ch].max() == 255
            else:
                assert np.array_equal(ac[..., ch], img[..., ch])
        eq = AO.apply_op(img, AO.EQUALIZE, z)
        for ch in range(3):                                                           # monotone per channel
            order = np.argsort(img[..., ch].reshape(-1), kind="stable")
            assert (np.diff(eq[..., ch].reshape(-1)[order].astype(int)) >= 0).all()
        c, p = A.encode_op("Rotate", 0.0, 48)
        assert np.array_equal(AO.apply_op(img, c, p), img)
        c, p = A.encode_op("Rotate", 180.0, 48)
        assert np.array_equal(AO.apply_op(img, c, p), img[::-1, ::-1])
        c, p = A.encode_op("Rotate", 90.0, 48)                                        # counter-clockwise
        assert np.array_equal(AO.apply_op(img, c, p), np.rot90(img, 1))
        c, p = A.encode_op("TranslateX", 5.0, 48)
        t = AO.apply_op(img, c, p)
        assert np.array_equal(t[:, 5:], img[:, :-5]) and not t[:, :5].any()
        c, p = A.encode_op("TranslateY", -7.9, 48)                                    # int() truncates toward zero
        t = AO.apply_op(img, c, p)
        assert np.array_equal(t[:-7], img[7:]) and not t[-7:].any()


def _plan_worker(rank, world, port, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    aug = A.AutoAugment("imagenet", seed=99, rank=dist.get_rank())
    ops, params = aug.plan(64, 224)
    import torch
    gathered = [torch.zeros(64, 2, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(ops))
    if rank == 0:
        out.put([g.numpy().tolist() for g in gathered])
    dist.destroy_process_group()


def test_data_parallel_ranks_draw_different_plans_from_one_seed():
    """N > 1 (gloo, two processes): replicas share the seed and differ by rank, so that they do not augment alike; a rank's plan
    is reproducible."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_plan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = np.array(got[0]), np.array(got[1])
    assert not np.array_equal(a, b)
    assert np.array_equal(a, A.AutoAugment("imagenet", seed=99, rank=0).plan(64, 224)[0])
    assert np.array_equal(b, A.AutoAugment("imagenet", seed=99, rank=1).plan(64, 224)[0])
