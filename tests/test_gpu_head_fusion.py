"""Head-decode fusion: the loss takes the head's RAW per-level conv outputs
(``p24.engine.RawLevels``) and decodes on load (yolo_head_24p.py:212-237).  Checked three ways:
  * against the oracle's head decode + loss run ON THE SAME GPU (bit-exact assignments, 1e-5 relative on values);
  * against this library's own decoded-buffer entry fed with the oracle-decoded tensor (every output bit-identical:
    same arithmetic, different loads);
  * the reference-made decode fixture (tests/golden/head_s64.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import synth
from p24.engine import RawLevels
from p24.losses import Loss_Function

from test_gpu_simota import _assert_assignment_equal, DEV, RTOL

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _levels(B, size, nc, seed):
    reg, obj, cls = synth.make_raw_levels(B, size, nc, seed=seed)
    return [t.to(DEV) for t in reg], [t.to(DEV) for t in obj], [t.to(DEV) for t in cls]


def _both(B, size, nc, lab, seed, steps=1):
    reg, obj, cls = _levels(B, size, nc, seed)
    gx, gy, gs, dec = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    labd = lab.to(DEV)
    raw_lf, row_lf, o = Loss_Function(nc), Loss_Function(nc), orc.LossOracle(nc)
    for _ in range(steps):
        r = o.forward((gx, gy, gs, dec.clone(), []), labd)
        res_raw, w_raw, a_raw = raw_lf.forward_async((gx, gy, gs, RawLevels(reg, obj, cls), []), labd)
        res_row, w_row, a_row = row_lf.forward_async((gx, gy, gs, dec, []), labd)
        _assert_assignment_equal(a_raw, o.trace)
        for name in ("fg_mask", "matched_gt", "pred_iou", "num_fg", "dyn_k", "sums28"):
            assert torch.equal(getattr(a_raw, name), getattr(a_row, name)), name
        assert torch.equal(res_raw, res_row) and torch.equal(w_raw, w_row)
        np.testing.assert_allclose(float(res_raw[0]), float(r[0]), rtol=RTOL)
        np.testing.assert_allclose(res_raw[1:25].cpu().numpy(), r[1].cpu().numpy(), rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(float(res_raw[25]), float(r[2]), rtol=RTOL)
        np.testing.assert_allclose(float(res_raw[26]), float(r[3]), rtol=RTOL, atol=1e-7)
    raw_lf.check_errors()


@pytest.mark.parametrize("kind", ["smooth", "spiky"])
def test_raw_levels_small(kind):
    lab = synth.make_labels(3, [6, 0, 9], 10, 256, 80, seed=7, kind=kind)
    _both(3, 256, 80, lab, seed=7, steps=2)


def test_raw_levels_config1_full_batch20():
    lab = synth.make_labels(20, 20, 50, 640, 80, seed=1, kind="smooth")
    _both(20, 640, 80, lab, seed=1)


def test_raw_levels_crowded_and_hires():
    _both(4, 640, 80, synth.make_labels(4, 100, 100, 640, 80, seed=2, kind="smooth"), seed=2)
    _both(2, 1280, 80, synth.make_labels(2, 20, 50, 1280, 80, seed=4, kind="spiky"), seed=4)


def test_raw_levels_channel_slices_of_one_tensor_and_forward_tuple():
    """The conv outputs may be channel slices of one [B, 27 + nc, H, W] tensor (planes dense, batch stride larger);
    forward() keeps the reference's 7-tuple with the drawing entries decoded from the raw planes."""
    nc, size = 80, 256
    reg, obj, cls = _levels(2, size, nc, 9)
    lab = synth.make_labels(2, [5, 3], 8, size, nc, seed=9, kind="smooth").to(DEV)
    gx, gy, gs, dec = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    whole = [torch.cat([r, o, c], 1) for r, o, c in zip(reg, obj, cls)]
    sl = RawLevels([w[:, :26] for w in whole], [w[:, 26:27] for w in whole], [w[:, 27:] for w in whole])
    a, b = Loss_Function(nc), Loss_Function(nc)
    ra = a.forward((gx, gy, gs, sl, []), lab)
    rb = b.forward((gx, gy, gs, dec, []), lab)
    for i in range(4):
        assert torch.equal(torch.as_tensor(ra[i]), torch.as_tensor(rb[i]))
    assert ra[4] == rb[4] and ra[5] == rb[5]
    for x, y in zip(ra[6], rb[6]):
        assert torch.equal(torch.as_tensor(x), torch.as_tensor(y))
    with pytest.raises(IndexError):
        a.forward_async((gx, gy, gs, RawLevels(reg[:2], obj[:2], cls[:2]), []), lab)


def test_raw_levels_reference_fixture():
    """Decode-on-load against the reference-made fixture: the foreground rows forward() returns for drawing are the
    reference's decoded rows (exp() may differ in the last bit between the CPU that made the fixture and the GPU)."""
    g = np.load(os.path.join(GOLD, "head_s64.npz"))
    B, size = int(g["batch"]), int(g["img_size"])
    reg, obj, cls = _levels(B, size, 80, int(g["seed"]))
    gx, gy, gs, dec = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    np.testing.assert_allclose(dec.cpu().numpy(), g["train"], rtol=2e-6, atol=0)
    lab = synth.make_labels(B, [2, 1], 4, size, 80, seed=3, kind="smooth", radius_range=(0.1, 0.3)).to(DEV)
    lf = Loss_Function(80)
    out = lf.forward((gx, gy, gs, RawLevels(reg, obj, cls), []), lab)
    fg = lf.last_assignment.fg_mask.view(-1).bool().cpu().numpy()
    want = g["train"].reshape(-1, 107)[fg]
    assert fg.sum() > 0
    np.testing.assert_allclose(out[6][0].cpu().numpy(), want[:, 0], rtol=2e-6)
    np.testing.assert_allclose(out[6][2].cpu().numpy(), want[:, 2:26], rtol=2e-6)


def test_raw_levels_backward_matches_autograd_through_torch_decode():
    """d(loss)/d(raw conv outputs): the fused raw backward (decode chain rule inside the kernel) against torch autograd
    through the oracle's decode into this library's decoded-buffer backward, and against the oracle end to end."""
    nc, size = 80, 256
    reg, obj, cls = _levels(2, size, nc, 15)
    lab = synth.make_labels(2, [5, 7], 8, size, nc, seed=15, kind="smooth").to(DEV)

    def leaves():
        return [[t.clone().requires_grad_(True) for t in lst] for lst in (reg, obj, cls)]

    # (1) fused
    r1, o1, c1 = leaves()
    gx, gy, gs, _ = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    out = Loss_Function(nc).forward((gx, gy, gs, RawLevels(r1, o1, c1), []), lab)
    out[0].backward()
    # (2) torch decode (autograd) -> decoded-buffer entry
    r2, o2, c2 = leaves()
    dec = orc.head_decode_train([t * 1.0 for t in r2], o2, c2, list(synth.STRIDES))[3]
    out2 = Loss_Function(nc).forward((gx, gy, gs, dec, []), lab)
    out2[0].backward()
    # (3) the oracle, all torch
    r3, o3, c3 = leaves()
    dec3 = orc.head_decode_train([t * 1.0 for t in r3], o3, c3, list(synth.STRIDES))[3]
    out3 = orc.LossOracle(nc).forward((gx, gy, gs, dec3, []), lab)
    out3[0].backward()
    assert float(out[0]) == float(out2[0])
    for mine, viarows, oracle in zip(r1 + o1 + c1, r2 + o2 + c2, r3 + o3 + c3):
        assert mine.grad is not None and torch.isfinite(mine.grad).all()
        scale = float(oracle.grad.abs().max()) + 1e-12
        torch.testing.assert_close(mine.grad, viarows.grad, rtol=1e-5, atol=scale * 1e-6)
        torch.testing.assert_close(mine.grad, oracle.grad, rtol=1e-4, atol=scale * 2e-6)


@pytest.mark.parametrize("conf,nms,agnostic", [(0.25, 0.45, False), (0.01, 0.3, True), (0.01, 0.65, False)])
def test_raw_levels_postprocess_vs_oracle_decode(conf, nms, agnostic):
    """Inference side of the fusion: postprocess on the raw conv outputs (sigmoid + decode inside the filter pass)
    against (1) this library's postprocess on the oracle-decoded prediction (yolo_head_24p.py:191, 201-211, 239-256) and
    (2) the oracle postprocess of that prediction: rows, order and counts bit for bit."""
    from p24 import boxes as p24_boxes
    from p24 import head as p24_head
    reg, obj, cls = _levels(4, 320, 80, 33)
    for t in obj + cls:  # raise the scores into the interesting range (the synthetic logits sit at the prior)
        t += 4.5
    pred = orc.head_decode_infer(reg, obj, cls, list(synth.STRIDES))
    fused = p24_boxes.postprocess(p24_head.infer_outputs(reg, obj, cls, synth.STRIDES), 80, conf, nms, agnostic)
    plain = p24_boxes.postprocess(pred, 80, conf, nms, agnostic)
    assert torch.equal(p24_head.infer_outputs(reg, obj, cls, synth.STRIDES, fused=False), pred)
    nonempty = 0
    for i in range(4):
        want = orc.postprocess_image(pred[i], 80, conf, nms, agnostic)
        for got in (fused[i], plain[i]):
            assert (got is None) == (want is None or want.shape[0] == 0)
            if got is not None:
                assert torch.equal(got, want), f"image {i}"
                nonempty += 1
    assert nonempty > 0


@pytest.mark.parametrize("nc", [1, 5, 81])
def test_raw_levels_other_class_counts(nc):
    """Class counts other than 80 through both raw entries (81 = the largest the training entry stages: 27 + nc <= 108)."""
    from p24 import boxes as p24_boxes
    from p24 import head as p24_head
    lab = synth.make_labels(2, [6, 3], 8, 256, nc, seed=11, kind="smooth")
    _both(2, 256, nc, lab, seed=11)
    reg, obj, cls = _levels(2, 256, nc, 12)
    for t in obj + cls:
        t += 4.0
    pred = orc.head_decode_infer(reg, obj, cls, list(synth.STRIDES))
    fused = p24_boxes.postprocess(p24_head.infer_outputs(reg, obj, cls, synth.STRIDES), nc, 0.2, 0.5, False)
    for i in range(2):
        want = orc.postprocess_image(pred[i], nc, 0.2, 0.5, False)
        assert (fused[i] is None) == (want is None)
        if want is not None:
            assert torch.equal(fused[i], want)


def test_raw_levels_postprocess_odd_level_sizes():
    """96 x 96 input: level sizes 144 / 36 / 9 (odd, not a multiple of the warp or of 4): tiles straddle two levels and
    the last tile is partial; same result as the decoded path."""
    from p24 import boxes as p24_boxes
    from p24 import head as p24_head
    reg, obj, cls = _levels(3, 96, 80, 44)
    for t in obj + cls:
        t += 4.5
    pred = orc.head_decode_infer(reg, obj, cls, list(synth.STRIDES))
    for conf, nms, ag in [(0.25, 0.45, False), (0.01, 0.65, False)]:
        fused = p24_boxes.postprocess(p24_head.infer_outputs(reg, obj, cls, synth.STRIDES), 80, conf, nms, ag)
        for i in range(3):
            want = orc.postprocess_image(pred[i], 80, conf, nms, ag)
            assert (fused[i] is None) == (want is None)
            if want is not None:
                assert torch.equal(fused[i], want)
