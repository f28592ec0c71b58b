"""Full-size BASELINE.json configurations against the oracle RUN ON THE SAME GPU (SURVEY.md 8c mode ii: torch CUDA
eager shares IEEE ops and libdevice transcendentals with the kernels, so discrete outputs must agree bit for bit).

  configs[1]  B=20, 640x640 (8400 anchors), 20 GT/img            both label kinds
  configs[2]  B=20, 640x640, 100 GT/img (crowded)
  configs[4]  1280x1280 (33600 anchors), 20 GT/img, B=4 slice    both label kinds

Bar: fg masks, matched GT indices, dynamic-k counts bit-exact; pair values / losses within RTOL = 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import synth
from p24.losses import Loss_Function

from test_gpu_simota import _assert_assignment_equal, _grids, DEV, RTOL

pytestmark = pytest.mark.gpu


def _compare(out, lab, size):
    gx, gy, gs = _grids(size)
    outd, labd = out.to(DEV), lab.to(DEV)
    mine, o = Loss_Function(80), orc.LossOracle(80)
    r = o.forward((gx, gy, gs, outd.clone(), []), labd)
    res, _, asg = mine.forward_async((gx, gy, gs, outd, []), labd)
    _assert_assignment_equal(asg, o.trace)
    np.testing.assert_allclose(float(res[0]), float(r[0]), rtol=RTOL)
    np.testing.assert_allclose(res[1:25].cpu().numpy(), r[1].cpu().numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(float(res[25]), float(r[2]), rtol=RTOL)
    np.testing.assert_allclose(float(res[26]), float(r[3]), rtol=RTOL, atol=1e-7)
    assert float(res[27]) == pytest.approx(r[5], rel=1e-6)


@pytest.mark.parametrize("kind", ["smooth", "spiky"])
def test_config1_full_batch20_vs_oracle_on_gpu(kind):
    out = synth.make_head_outputs(20, 640, 80, seed=1)
    lab = synth.make_labels(20, 20, 50, 640, 80, seed=1, kind=kind)
    _compare(out, lab, 640)


def test_config2_crowded_full_batch20_vs_oracle_on_gpu():
    out = synth.make_head_outputs(20, 640, 80, seed=2)
    lab = synth.make_labels(20, 100, 100, 640, 80, seed=2, kind="smooth")
    _compare(out, lab, 640)


@pytest.mark.parametrize("kind", ["smooth", "spiky"])
def test_config5_hires_batch4_vs_oracle_on_gpu(kind):
    out = synth.make_head_outputs(4, 1280, 80, seed=4)
    lab = synth.make_labels(4, 20, 50, 1280, 80, seed=4, kind=kind)
    _compare(out, lab, 1280)


def test_get_assignments_repeated_same_shape_then_forward():
    """The reference calls get_assignments once per image, so the same (1, A, num_gt) shape repeats: the second call on
    the same workspace must not see state left by the first, nor disturb a following forward()."""
    size = 320
    out = synth.make_head_outputs(3, size, 80, seed=43).to(DEV)
    lab = synth.make_labels(3, [4, 4, 4], 4, size, 80, seed=43, kind="smooth").to(DEV)
    xs, ys, ss = synth.make_grids(size)
    X, Y, S = torch.cat(xs, 1).to(DEV), torch.cat(ys, 1).to(DEV), torch.cat(ss, 1).to(DEV)
    bbox, obj, cls = out[:, :, :26], out[:, :, 26].unsqueeze(-1), out[:, :, 27:]
    lf = Loss_Function(80)
    for rep in range(2):
        for b in range(3):
            gt50, gcls = lab[b, :, 1:], lab[b, :, 0]
            want = orc.get_assignments(4, out.shape[1], gt50, gcls, bbox[b], S, X, Y, cls[b], obj[b], 80)
            got = lf.get_assignments(b, 4, out.shape[1], gt50, gcls, bbox[b], S, X, Y, cls, bbox, obj)
            assert torch.equal(got[1], want[1]) and torch.equal(got[3], want[3]) and got[4] == want[4], (rep, b)
    # forward with B=1 and Lmax=4 shares the (1, A, 4) workspace of the calls above
    gx, gy, gs = _grids(size)
    o = orc.LossOracle(80)
    r = o.forward((gx, gy, gs, out[:1].clone(), []), lab[:1])
    res, _, asg = lf.forward_async((gx, gy, gs, out[:1], []), lab[:1])
    _assert_assignment_equal(asg, o.trace)
    np.testing.assert_allclose(float(res[0]), float(r[0]), rtol=RTOL)


@pytest.mark.parametrize("nc,shift", [(80, 6.0), (365, 3.6), (400, 0.0)])
def test_class_bce_sum_does_not_overflow(nc, shift):
    """Uniformly large class logits / many classes: sum_j softplus(x_j) exceeds 88.7 at a foreground anchor (the fp32
    product form overflows there); the loss must stay finite and agree with the oracle."""
    size = 256
    out = synth.make_head_outputs(2, size, nc, seed=61)
    out[:, :, 27:] += shift
    lab = synth.make_labels(2, [3, 2], 6, size, nc, seed=61, kind="smooth")
    gx, gy, gs = _grids(size)
    outd, labd = out.to(DEV), lab.to(DEV)
    mine, o = Loss_Function(nc), orc.LossOracle(nc)
    r = o.forward((gx, gy, gs, outd.clone(), []), labd)
    res, _, asg = mine.forward_async((gx, gy, gs, outd, []), labd)
    assert torch.isfinite(res).all()
    _assert_assignment_equal(asg, o.trace)
    np.testing.assert_allclose(float(res[26]), float(r[3]), rtol=RTOL)
    np.testing.assert_allclose(float(res[0]), float(r[0]), rtol=RTOL)


def test_pipelined_steps_give_the_bits_of_plain_steps():
    """Loss_Function.pipelined (P24_F_EARLY_PREP): the preparation kernel of a step runs beside the last kernel of the
    step before it, on double-buffered workspace halves.  Eight back-to-back steps over three resident batches (different
    GT counts, one image without GTs) must reproduce the plain sequence bit for bit, including the stateful re-weighting."""
    gx, gy, gs = _grids(640)
    sets = []
    for i, counts in enumerate([[20] * 6, [3, 0, 50, 7, 1, 12], [33] * 6]):
        sets.append((synth.make_head_outputs(6, 640, 80, seed=70 + i).to(DEV),
                     synth.make_labels(6, counts, 50, 640, 80, seed=70 + i, kind="spiky" if i == 1 else "smooth").to(DEV)))
    plain, piped = Loss_Function(80), Loss_Function(80)
    piped.pipelined = True
    want = []
    for s in range(8):
        o, l = sets[s % 3]
        r, _, a = plain.forward_async((gx, gy, gs, o, []), l)
        want.append((r.clone(), a.fg_mask.clone(), a.matched_gt.clone(), a.dyn_k.clone(), a.num_gt.clone(), a.num_fg.clone()))
    torch.cuda.synchronize()
    got = []
    for s in range(8):   # enqueued back to back, no synchronisation in between
        o, l = sets[s % 3]
        r, _, a = piped.forward_async((gx, gy, gs, o, []), l)
        got.append((r, a.fg_mask, a.matched_gt, a.dyn_k, a.num_gt, a.num_fg))
    torch.cuda.synchronize()
    for s in range(8):
        for x, y in zip(want[s], got[s]):
            assert torch.equal(x, y), s
    piped.check_errors()
