"""nvit_augment_u8 on a B200 against the numpy oracle (oracle/augment_oracle.py, pinned against torchvision in
tests/test_augment_cpu.py): BIT-EXACT uint8 output for every operation, every ordered pair of operations (the
global -> shared -> global path), whole sampled batches at 32 and 224 px, odd sizes (byte path), and the loader hook."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from nvit_b200 import DeviceLoader, augment as A, ops  # noqa: E402
from oracle import augment_oracle as AO  # noqa: E402
from test_augment_cpu import ALL_OPS, images  # noqa: E402

DEV = "cuda"


def run(X, ops_h, params_h):
    x = torch.from_numpy(X).to(DEV)
    out = torch.empty_like(x)
    ops.augment_u8(x, out, torch.from_numpy(ops_h).to(DEV), torch.from_numpy(params_h).to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), torch.from_numpy(X)), "the input batch was modified"
    return out.cpu().numpy()


def encoded_cases(S):
    cases = [("Identity", 0.0)]
    for op in ALL_OPS:
        bins = [None] if op in ("AutoContrast", "Equalize", "Invert") else [1, 5, 9]
        for b in bins:
            for sign in ((1, -1) if op in A._SIGNED else (1,)):
                cases.append((op, sign * A.magnitude(op, b, S) if b is not None else 0.0))
    return cases


@pytest.mark.parametrize("S", [32, 57, 224])
def test_every_operation_alone_in_either_stage_is_bit_exact(S):
    X = images(S, 3, seed=S + 1)
    cases = encoded_cases(S)
    for stage in (0, 1):
        n = len(cases) * len(X)
        batch = np.repeat(X, len(cases), axis=0)
        ops_h = np.zeros((n, 2), np.int32)
        params_h = np.zeros((n, 2, 8), np.float32)
        for i in range(len(X)):
            for k, (op, mag) in enumerate(cases):
                code, p = A.encode_op(op, mag, S)
                ops_h[i * len(cases) + k, stage] = code
                params_h[i * len(cases) + k, stage] = p
        got = run(batch, ops_h, params_h)
        want = AO.apply_plan(batch, ops_h, params_h)
        for j in range(n):
            assert np.array_equal(got[j], want[j]), (S, stage, cases[j % len(cases)], int((got[j] != want[j]).sum()))


def test_every_ordered_pair_of_operations_is_bit_exact():
    S = 40
    X = images(S, 2, seed=9)
    singles = [("Rotate", 20.0), ("ShearX", -0.2), ("TranslateY", 7.0), ("Brightness", 0.5), ("Color", -0.7), ("Contrast", 0.9),
               ("Sharpness", 0.9), ("Posterize", 4.0), ("Solarize", 113.3), ("AutoContrast", 0.0), ("Equalize", 0.0), ("Invert", 0.0)]
    enc = [A.encode_op(op, mag, S) for op, mag in singles]
    pairs = [(a, b) for a in range(len(enc)) for b in range(len(enc))]
    for i in range(len(X)):
        batch = np.repeat(X[i:i + 1], len(pairs), axis=0)
        ops_h = np.array([[enc[a][0], enc[b][0]] for a, b in pairs], np.int32)
        params_h = np.array([[enc[a][1], enc[b][1]] for a, b in pairs], np.float32)
        got = run(batch, ops_h, params_h)
        want = AO.apply_plan(batch, ops_h, params_h)
        for j, (a, b) in enumerate(pairs):
            assert np.array_equal(got[j], want[j]), (singles[a], singles[b], int((got[j] != want[j]).sum()))


@pytest.mark.parametrize("S", [1, 2, 3, 5])
def test_degenerate_image_sizes(S):
    """1 x 1 ... 5 x 5 images: sharpness leaves sides <= 2 alone, a 3 x 3 image has one interior pixel, statistics of a
    single pixel are degenerate (autocontrast / equalize return the image), gathers mostly fall outside."""
    rng = np.random.default_rng(S)
    X = rng.integers(0, 256, (6, S, S, 3), dtype=np.uint8)
    names = [("Identity", 0.0), ("Rotate", 30.0), ("ShearY", 0.3), ("TranslateX", 1.0), ("Brightness", -0.4), ("Color", 0.9), ("Contrast", -0.9),
             ("Sharpness", 0.9), ("Posterize", 5.0), ("Solarize", 28.3), ("AutoContrast", 0.0), ("Equalize", 0.0), ("Invert", 0.0)]
    enc = [A.encode_op(op, mag, S) for op, mag in names]
    pairs = [(a, b) for a in range(len(enc)) for b in range(len(enc))]
    for i in range(len(X)):
        batch = np.repeat(X[i:i + 1], len(pairs), axis=0)
        ops_h = np.array([[enc[a][0], enc[b][0]] for a, b in pairs], np.int32)
        params_h = np.array([[enc[a][1], enc[b][1]] for a, b in pairs], np.float32)
        got = run(batch, ops_h, params_h)
        want = AO.apply_plan(batch, ops_h, params_h)
        for j, (a, b) in enumerate(pairs):
            assert np.array_equal(got[j], want[j]), (S, names[a], names[b], got[j].tolist(), want[j].tolist())
    # out-of-range operation codes act as identity (the entry point's contract)
    bad = np.array([[99, -3]] * len(X), np.int32)
    assert np.array_equal(run(X, bad, np.zeros((len(X), 2, 8), np.float32)), X)


@pytest.mark.parametrize("S", [256, 258])
def test_large_images_run_sharpness_in_many_row_bands(S):
    """At 256 / 258 px the image leaves room for ~20 scratch rows only: the in-place sharpness walks many bands (word path at
    256, byte path at 258, where 3 S is not a multiple of four) and must still read ORIGINAL neighbour rows."""
    X = images(S, 1, seed=S)
    names = [("Sharpness", 0.9), ("Sharpness", -0.9), ("Rotate", 11.0), ("Equalize", 0.0), ("Color", 0.6), ("Identity", 0.0)]
    enc = [A.encode_op(op, mag, S) for op, mag in names]
    pairs = [(a, b) for a in range(len(enc)) for b in range(len(enc)) if 0 in (a, b) or 1 in (a, b)]
    for i in range(len(X)):
        batch = np.repeat(X[i:i + 1], len(pairs), axis=0)
        ops_h = np.array([[enc[a][0], enc[b][0]] for a, b in pairs], np.int32)
        params_h = np.array([[enc[a][1], enc[b][1]] for a, b in pairs], np.float32)
        got = run(batch, ops_h, params_h)
        want = AO.apply_plan(batch, ops_h, params_h)
        for j, (a, b) in enumerate(pairs):
            assert np.array_equal(got[j], want[j]), (S, names[a], names[b], int((got[j] != want[j]).sum()))


@pytest.mark.parametrize("dataset,S,B", [("cifar10", 32, 256), ("svhn", 32, 64), ("imagenet", 224, 48)])
def test_sampled_batches_are_bit_exact(dataset, S, B):
    base = images(S, 7, seed=5)
    X = base[np.arange(B) % len(base)]
    X = np.ascontiguousarray(np.roll(X, shift=3, axis=2))
    aug = A.AutoAugment(dataset, seed=11)
    for _ in range(2):
        ops_h, params_h = aug.plan(B, S)
        got = run(X, ops_h, params_h)
        want = AO.apply_plan(X, ops_h, params_h)
        assert np.array_equal(got, want), int((got != want).sum())
    # the public call: same generator state -> same plan -> same pixels
    a1, a2 = A.AutoAugment(dataset, seed=3), A.AutoAugment(dataset, seed=3)
    y = a1(torch.from_numpy(X).to(DEV))
    torch.cuda.synchronize()
    o, p = a2.plan(B, S)
    assert np.array_equal(a1.last_plan[0], o)
    assert np.array_equal(y.cpu().numpy(), AO.apply_plan(X, o, p))


def test_full_size_batch_properties_and_argument_checks():
    """BASELINE config 2's batch (256 x 224 x 224 x 3): identity is a copy, invert twice is the identity, equalize is
    idempotent on its own output's histogram class (second pass changes nothing the oracle would not), plus loud failures."""
    S, B = 224, 256
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).to(DEV)
    z_ops = torch.zeros(B, 2, dtype=torch.int32, device=DEV)
    z_par = torch.zeros(B, 2, 8, dtype=torch.float32, device=DEV)
    out = torch.empty_like(x)
    ops.augment_u8(x, out, z_ops, z_par)
    assert torch.equal(out, x)
    inv = torch.full((B, 2), A.INVERT, dtype=torch.int32, device=DEV)
    ops.augment_u8(x, out, inv, z_par)
    assert torch.equal(out, x)                       # invert . invert
    one = inv.clone()
    one[:, 1] = 0
    ops.augment_u8(x, out, one, z_par)
    assert torch.equal(out, 255 - x)
    # a 180-degree rotation twice is the identity (exact integer source coordinates)
    code, p = A.encode_op("Rotate", 180.0, S)
    rot = torch.zeros(B, 2, dtype=torch.int32, device=DEV)
    rot[:, :] = code
    rp = torch.tensor(p, dtype=torch.float32, device=DEV).repeat(B, 2, 1).contiguous()
    ops.augment_u8(x, out, rot, rp)
    assert torch.equal(out, x)
    rot[:, 1] = 0
    ops.augment_u8(x, out, rot, rp)
    assert torch.equal(out, x.flip(1).flip(2))
    with pytest.raises(ValueError):
        ops.augment_u8(x, x, z_ops, z_par)
    with pytest.raises(TypeError):
        ops.augment_u8(x.float(), out, z_ops, z_par)
    with pytest.raises(ValueError):
        ops.augment_u8(x, out, z_ops[:5], z_par)
    big = torch.zeros(1, 304, 304, 3, dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.augment_u8(big, torch.empty_like(big), z_ops[:1], z_par[:1])
    rgba = torch.zeros(1, 8, 8, 4, dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError, match="RGB"):
        ops.augment_u8(rgba, torch.empty_like(rgba), z_ops[:1], z_par[:1])


def test_device_loader_applies_the_train_transform():
    S, B = 32, 16
    base = images(S, 3, seed=21)
    batches = [(torch.from_numpy(base[np.arange(B) % len(base)].copy()), torch.arange(B) % 10) for _ in range(4)]
    train_tf, val_tf = A.get_transforms("cifar10", seed=5)
    twin = A.AutoAugment("cifar10", seed=5)
    for (Xd, yd), (Xh, yh) in zip(DeviceLoader(batches, DEV, transform=train_tf), batches):
        o, p = twin.plan(B, S)
        assert np.array_equal(Xd.cpu().numpy(), AO.apply_plan(Xh.numpy(), o, p))
        assert torch.equal(yd.cpu(), yh)
    for (Xd, yd), (Xh, yh) in zip(DeviceLoader(batches, DEV, transform=val_tf), batches):
        assert torch.equal(Xd.cpu(), Xh)
