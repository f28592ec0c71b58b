"""GPU parity tests of the element-wise IoU loss, the pairwise pair value, dynamic_k_matching on materialised matrices
and the backward passes, through the C ABI, against the oracle (same GPU) and the reference-made goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import boxes as p24_boxes
from p24 import synth
from p24.losses import IOUloss, Loss_Function

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5  # BASELINE.json: IoU / loss values within 1e-5 relative (fp32)
DEV = "cuda:0"


def _pairs(n, size, seed, kind="smooth"):
    lab = synth.make_labels(1, n, n, size, 80, seed=seed, kind=kind)[0, :, 1:].to(DEV)
    pred = synth.make_head_outputs(1, size, 80, seed=seed)[0, :, :26].to(DEV)
    # matched-like pairs: predictions near their GT (partial-overlap regime) plus random far ones
    near = pred[:n].clone()
    near[:, 0] = lab[:, 0] + 6.0 * torch.randn(n, device=DEV, generator=torch.Generator(DEV).manual_seed(seed))
    near[:, 1] = lab[:, 1] + 6.0 * torch.randn(n, device=DEV, generator=torch.Generator(DEV).manual_seed(seed + 1))
    near[:, 2:] = near[:, 2:] * 4.0
    return lab, near, pred


def test_known_answer_table():
    g = np.load(os.path.join(GOLD, "known_answers.npz"))
    gt, pd = torch.from_numpy(g["gt"]).to(DEV), torch.from_numpy(g["pred"]).to(DEV)
    loss24, draw = IOUloss().forward(pd, gt)
    np.testing.assert_allclose(loss24.cpu().numpy(), g["loss24"], rtol=RTOL, atol=1e-6)
    cx, cy, rg = orc._gt_radii(gt)
    inter, dist = IOUloss().circle_inter(cx, cy, rg, pd[:, 0], pd[:, 1], pd[:, 2:])
    np.testing.assert_allclose(inter.cpu().numpy(), g["inter"], rtol=RTOL, atol=1e-3)
    np.testing.assert_allclose(dist.cpu().numpy(), g["dist"], rtol=RTOL)
    pair = torch.stack([p24_boxes.bboxes_iou(gt[i:i + 1], pd[i:i + 1])[0, 0] for i in range(gt.shape[0])])
    np.testing.assert_allclose(pair.cpu().numpy(), g["pair_iou"], rtol=RTOL, atol=1e-7)
    assert torch.equal(draw[2], pd[:, 2:])


def test_iou_loss_forward_matches_oracle_and_error_behaviour():
    lab, near, far = _pairs(40, 640, 3)
    for pred in (near, far[:40]):
        got, _ = IOUloss().forward(pred, lab)
        want, _ = orc.iou_loss_forward(pred, lab)
        torch.testing.assert_close(got, want, rtol=RTOL, atol=1e-6)
    with pytest.raises(IndexError):
        IOUloss().forward(near[:, :25], lab)
    e, d = IOUloss().forward(near[:0], lab[:0])
    assert e.shape == (1, 24) and float(e.abs().sum()) == 0 and d[0].shape == (1, 24)
    r, dd = IOUloss().circle_inter(lab[:0, 0], lab[:0, 1], lab[:0, 2:26], near[:0, 0], near[:0, 1], near[:0, 2:])
    assert r.shape == (0, 24) and dd.shape == (0, 24)


def test_iou_loss_backward_matches_autograd_of_the_oracle():
    lab, near, far = _pairs(64, 640, 5, kind="spiky")
    pred = torch.cat([near, far[:64]], 0)
    tgt = torch.cat([lab, lab], 0)
    w = torch.rand(128, 24, device=DEV, generator=torch.Generator(DEV).manual_seed(9))
    p1 = pred.clone().requires_grad_(True)
    (IOUloss().forward(p1, tgt)[0] * w).sum().backward()
    p2 = pred.clone().requires_grad_(True)
    (orc.iou_loss_forward(p2, tgt)[0] * w).sum().backward()
    scale = p2.grad.abs().max()
    torch.testing.assert_close(p1.grad, p2.grad, rtol=1e-4, atol=float(scale) * 2e-6)


def test_pairwise_value_and_circle_inter_match_oracle():
    lab, near, far = _pairs(9, 320, 7, kind="spiky")
    pred = far[:700]
    torch.testing.assert_close(p24_boxes.bboxes_iou(lab, pred), orc.bboxes_iou(lab, pred), rtol=RTOL, atol=1e-7)
    cx, cy, rg = orc._gt_radii(lab)
    got = p24_boxes.circle_inter(cx, cy, rg, pred[:50, 0], pred[:50, 1], pred[:50, 2:])
    want = orc.circle_inter_pairwise(cx, cy, rg, pred[:50, 0], pred[:50, 1], pred[:50, 2:])
    torch.testing.assert_close(got[0], want[0], rtol=RTOL, atol=1e-3)
    torch.testing.assert_close(got[1], want[1], rtol=RTOL, atol=1e-6)
    with pytest.raises(IndexError):
        p24_boxes.bboxes_iou(lab[:, :49], pred)


def test_dynamic_k_matching_on_materialised_matrices():
    size = 320
    out = synth.make_head_outputs(1, size, 80, seed=13).to(DEV)[0]
    lab = synth.make_labels(1, 6, 8, size, 80, seed=13, kind="smooth").to(DEV)[0, :6]
    xs, ys, ss = synth.make_grids(size)
    X, Y, S = torch.cat(xs, 1).to(DEV), torch.cat(ys, 1).to(DEV), torch.cat(ss, 1).to(DEV)
    fg, both = orc.get_in_boxes_info(lab[:, 1:], S, X, Y, out.shape[0])
    cost, ious = orc.pair_cost(lab[:, 1:], lab[:, 0], out[fg, :26], out[fg, 27:], out[fg, 26:27], both, 80)
    fg_o, fg_m = fg.clone(), fg.clone()
    want = orc.dynamic_k_matching(cost, ious, lab[:, 0], 6, fg_o)
    lf = Loss_Function(80)
    got = lf.dynamic_k_matching(cost, ious, lab[:, 0], 6, fg_m)
    assert got[0] == want[0] and torch.equal(fg_o, fg_m)
    assert torch.equal(got[1], want[1]) and torch.equal(got[3], want[3])
    torch.testing.assert_close(got[2], want[2], rtol=0, atol=0)
    assert lf.last_dynamic_ks.tolist() == list(want[4])


def test_loss_backward_matches_autograd_of_the_oracle():
    """Gradient of the total loss w.r.t. the head output (train_24p.py:101): reg channels of the fg rows, the obj
    channel of every row, the cls channels of the fg rows."""
    size = 320
    out = synth.make_head_outputs(3, size, 80, seed=17).to(DEV)
    lab = synth.make_labels(3, [5, 0, 3], 8, size, 80, seed=17, kind="smooth").to(DEV)
    xs, ys, ss = synth.make_grids(size)
    gx, gy, gs = [t.to(DEV) for t in xs], [t.to(DEV) for t in ys], [t.to(DEV) for t in ss]
    mine, o = Loss_Function(80), orc.LossOracle(80)
    for step in range(2):  # the second step exercises non-trivial weights
        a = out.clone().requires_grad_(True)
        r = mine.forward((gx, gy, gs, a, []), lab)
        (r[0] * 1.7).backward()
        b = out.clone().requires_grad_(True)
        ro = o.forward((gx, gy, gs, b, []), lab)
        (ro[0] * 1.7).backward()
        np.testing.assert_allclose(float(r[0]), float(ro[0]), rtol=RTOL)
        scale = float(b.grad.abs().max())
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=scale * 2e-6)
        assert float(a.grad[1, :, :26].abs().sum()) == 0.0  # image without GT: only the obj channel has gradient
