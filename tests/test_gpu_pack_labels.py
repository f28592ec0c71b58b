"""GPU label packing through the C ABI against the reference-made fixture and the oracle:
fp32 labels bit-exact (the reference computes in float64 and casts once; so does the kernel)."""
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24.data import TrainTransform

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_pack_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "pack_labels.npz"))
    counts, shapes = g["counts"].tolist(), [tuple(s) for s in g["shapes"].tolist()]
    targets, off = [], 0
    for n in counts:
        targets.append(g["targets"][off:off + n] if n else np.zeros((1, 0)))
        off += n
    labels, nlabel = TrainTransform(50).pack(targets, shapes, (640, 640), "cuda:0", return_counts=True)
    for i in range(len(counts)):
        assert np.array_equal(labels[i].cpu().numpy(), g[f"labels{i}"]), f"image {i}"
    want_n = (labels.sum(dim=2) > 0).sum(dim=1)   # losses.py:190
    assert torch.equal(nlabel.long(), want_n)
    assert nlabel.tolist() == [min(c, 50) for c in counts]


@pytest.mark.parametrize("max_labels,dim", [(50, (640, 640)), (100, (1280, 1280)), (5, (416, 640))])
def test_pack_random_batches_vs_oracle(max_labels, dim):
    rng = np.random.default_rng(max_labels)
    B = 20
    counts = rng.integers(0, max_labels + 8, B).tolist()
    counts[3] = 0
    shapes = [(int(rng.integers(200, dim[0] + 1)), int(rng.integers(200, dim[1] + 1))) for _ in range(B)]
    targets = [np.concatenate([rng.integers(0, 80, (n, 1)).astype(np.float64), rng.random((n, 50))], 1) if n
               else np.zeros((1, 0)) for n in counts]
    labels = TrainTransform(max_labels).pack(targets, shapes, dim, "cuda:0").cpu().numpy()
    for b in range(B):
        assert np.array_equal(labels[b], orc.pack_labels(targets[b], shapes[b], dim, max_labels)), f"image {b}"


def test_pack_empty_batch_and_bad_rows():
    lab = TrainTransform(50).pack([np.zeros((1, 0)), np.zeros((0, 51))], [(640, 640), (480, 640)], (640, 640), "cuda:0")
    assert lab.shape == (2, 50, 51) and float(lab.abs().sum()) == 0.0
    with pytest.raises(IndexError):
        TrainTransform(50).pack([np.zeros((2, 27))], [(640, 640)], (640, 640), "cuda:0")


def test_packed_labels_feed_the_loss():
    """The packed batch goes straight into Loss_Function.forward (same device, no host round trip)."""
    from p24 import synth
    from p24.losses import Loss_Function
    lab = synth.make_labels(2, [4, 2], 50, 320, 80, seed=3, kind="smooth")          # pixels at 320
    targets = []
    for b, n in enumerate([4, 2]):
        t = lab[b, :n].double().numpy().copy()
        t[:, 1:] /= 320.0
        targets.append(t)
    packed = TrainTransform(50).pack(targets, [(320, 320), (320, 320)], (320, 320), "cuda:0")
    np.testing.assert_allclose(packed.cpu().numpy(), lab.numpy(), rtol=1e-6, atol=1e-4)
    out = synth.make_head_outputs(2, 320, 80, seed=3).to("cuda:0")
    xs, ys, ss = synth.make_grids(320, device="cuda:0")
    r = Loss_Function(80).forward((xs, ys, ss, out, []), packed)
    assert torch.isfinite(r[0]) and r[5] > 0
