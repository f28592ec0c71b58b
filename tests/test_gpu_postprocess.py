"""GPU parity tests of the inference postprocess (through the C ABI) against the reference-made goldens and the
oracle.  Kept rows are gathered copies of the input, so rows AND their order (the NMS keep-list) must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import boxes as p24_boxes
from p24 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _iou_margin(rect, thr):
    """smallest |IoU - thr| over all pairs (float64): keep-lists can only be demanded bit-exact away from it."""
    r = rect.double()
    area = (r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1])
    lt = torch.max(r[:, None, :2], r[None, :, :2])
    rb = torch.min(r[:, None, 2:], r[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = inter / (area[:, None] + area[None, :] - inter)
    iou.fill_diagonal_(0)
    return float((iou - thr).abs().min())


def test_golden_fixture_from_reference():
    g = np.load(os.path.join(GOLD, "post_s256.npz"))
    p = torch.from_numpy(g["prediction"]).to(DEV)
    si = 0
    while f"cfg{si}" in g.files:
        c, n, ag = g[f"cfg{si}"].tolist()
        res = p24_boxes.postprocess(p, 80, c, n, bool(ag))
        for i, r in enumerate(res):
            want = g[f"cfg{si}_img{i}"]
            if want.shape[0] == 0:
                assert r is None
            else:
                assert r is not None and tuple(r.shape) == want.shape
                assert np.array_equal(r.cpu().numpy(), want), f"cfg {si} image {i}: rows / keep order differ"
        si += 1
    assert si == 4


@pytest.mark.parametrize("conf,nms,agnostic", [(0.25, 0.45, False), (0.01, 0.3, True), (0.01, 0.65, False)])
def test_config4_batch_vs_oracle(conf, nms, agnostic):
    """BASELINE.json configs[3] at a reduced batch (8 of 64 images, 640x640) against the oracle ON THE CPU."""
    p = synth.make_postprocess_input(8, 640, 80, seed=3)
    pd = p.to(DEV)
    # second check (the first is the same-GPU oracle below): against the oracle on the CPU, with the CPU-evaluated
    # spiral coefficients (the reference evaluates them on the prediction's device, boxes.py:30-33)
    cpu_coef = p24_boxes.spiral_coefficients("cpu")
    cand, cnt, rows, keep, rects = p24_boxes.postprocess_raw(pd, 80, conf, nms, agnostic, want_rects=True, coef=cpu_coef)
    res = [rows[i, :int(cnt[i])] if int(cnt[i]) else None for i in range(p.shape[0])]
    for i in range(p.shape[0]):
        want, dbg = orc.postprocess_image(p[i], 80, conf, nms, agnostic, return_debug=True)
        assert int(cand[i]) == dbg["cand"].numel()
        # decoded rectangle of every candidate (polygon coordinates gate: 1e-5 relative; here bit-exact)
        assert torch.equal(rects[i, :int(cand[i])].cpu(), dbg["rect"])
        if _iou_margin(dbg["rect"], float(np.float32(nms))) < 1e-6:
            continue  # a pair sits on the threshold: fp32 implementations may legitimately differ
        assert int(cnt[i]) == want.shape[0]
        assert torch.equal(res[i].cpu(), want), f"image {i}: rows / keep order differ"
        assert torch.equal(keep[i, :int(cnt[i])].cpu().long(), dbg["cand"][dbg["keep"]])


def test_no_detections_one_detection_and_empty_inputs():
    p = synth.make_postprocess_input(2, 320, 80, seed=5).to(DEV)
    assert p24_boxes.postprocess(p, 80, 1.5, 0.45) == [None, None]
    one = p.clone()
    one[:, :, 26] = 0.0
    one[0, 77, 26] = 1.0
    one[0, 77, 27 + 5] = 0.9
    r = p24_boxes.postprocess(one, 80, 0.5, 0.45)
    assert r[1] is None and r[0].shape == (1, 29) and float(r[0][0, 28]) == 5.0 and torch.equal(r[0][0, :27], one[0, 77, :27])
    assert p24_boxes.postprocess(p[:, :0], 80) == [None, None]
    with pytest.raises(IndexError):
        p24_boxes.postprocess(p[:, :, :100], 80)


def test_strided_view_without_bulk_copy_path():
    """A row-padded view (row stride != 27 + nc) takes the plain-load path and must give the same result."""
    p = synth.make_postprocess_input(2, 320, 80, seed=6).to(DEV)
    wide = torch.zeros(2, p.shape[1], 112, device=DEV)
    wide[:, :, :107] = p
    a = p24_boxes.postprocess(p, 80, 0.25, 0.45)
    b = p24_boxes.postprocess(wide[:, :, :107], 80, 0.25, 0.45)
    for x, y in zip(a, b):
        assert (x is None) == (y is None) and (x is None or torch.equal(x, y))


def test_config4_full_batch64_properties():
    """BASELINE.json configs[3] at full size (B=64): per-image independence, score ordering, and the NMS invariant
    (no kept rectangle overlaps an earlier kept one of its class above the threshold)."""
    p = synth.make_postprocess_input(64, 640, 80, seed=3).to(DEV)
    conf, nms = 0.25, 0.45
    cand, cnt, rows, keep, rects = p24_boxes.postprocess_raw(p, 80, conf, nms, False, want_rects=True)
    single = p24_boxes.postprocess_raw(p[17:18], 80, conf, nms, False)
    n17 = int(cnt[17])
    assert n17 == int(single[1][0]) and torch.equal(rows[17, :n17], single[2][0, :n17])
    cx, cy = p24_boxes.spiral_coefficients(DEV)
    cx, cy = cx.to(DEV), cy.to(DEV)
    for b in (0, 31, 63):
        n = int(cnt[b])
        r = rows[b, :n]
        score = r[:, 26] * r[:, 27]
        assert (score >= conf).all() and (score[:-1] >= score[1:]).all()
        assert torch.equal(r[:, :27], p[b, keep[b, :n].long(), :27])
        px = r[:, 2:26] * cx + r[:, 0:1]
        py = r[:, 2:26] * cy + r[:, 1:2]
        box = torch.stack([px.min(1).values, py.min(1).values, px.max(1).values, py.max(1).values], 1).double()
        area = (box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1])
        lt = torch.max(box[:, None, :2], box[None, :, :2])
        rb = torch.min(box[:, None, 2:], box[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        iou = inter / (area[:, None] + area[None, :] - inter)
        same = r[:, 28][:, None] == r[:, 28][None, :]
        iou = torch.where(same, iou, torch.zeros_like(iou))
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= nms + 1e-5


@pytest.mark.parametrize("conf,nms,agnostic", [(0.25, 0.45, False), (0.01, 0.3, True), (0.01, 0.65, False)])
def test_config4_full_batch64_vs_oracle_on_gpu(conf, nms, agnostic):
    """BASELINE.json configs[3] at full size: all 64 images at the three settings of SURVEY.md 8(d) against the oracle
    run ON THE SAME GPU (SURVEY.md 8c mode ii): torchvision's CUDA nms / batched_nms (coordinate trick up to 25 000
    boxes) defines the keep-list.  Rows, their order, the candidate rectangles and the keep indices bit-exact."""
    p = synth.make_postprocess_input(64, 640, 80, seed=3).to(DEV)
    cand, cnt, rows, keep, rects = p24_boxes.postprocess_raw(p, 80, conf, nms, agnostic, want_rects=True)
    res = p24_boxes.postprocess(p, 80, conf, nms, agnostic)
    skipped = 0
    for i in range(p.shape[0]):
        want, dbg = orc.postprocess_image(p[i], 80, conf, nms, agnostic, return_debug=True)
        assert int(cand[i]) == dbg["cand"].numel()
        assert torch.equal(rects[i, :int(cand[i])], dbg["rect"]), f"image {i}: candidate rectangles differ"
        if dbg["rect"].shape[0] <= 2500 and _iou_margin(dbg["rect"].cpu(), float(np.float32(nms))) < 1e-7:
            skipped += 1  # a pair sits ON the threshold in float64 terms: fp32 implementations may legitimately differ
            continue
        assert int(cnt[i]) == want.shape[0], f"image {i}: {int(cnt[i])} kept, oracle {want.shape[0]}"
        assert torch.equal(res[i], want), f"image {i}: rows / keep order differ"
        assert torch.equal(keep[i, :int(cnt[i])].long(), dbg["cand"][dbg["keep"]])
    assert skipped <= 2


def test_opt_in_polar_decode_vs_oracle_and_drawn_polygons():
    """decode="polar" (opt-in: it changes the results): rectangles of the true polygon
    r_k (cos, sin)(theta_k) -- the decode the reference draws with (show_24p.py:346-348) -- through the same kernels; rows
    and order against the oracle run with those coefficients, and the drawn vertices stay inside the NMS rectangle."""
    p = synth.make_postprocess_input(4, 640, 80, seed=12).to(DEV)
    got = p24_boxes.postprocess(p, 80, 0.25, 0.45, False, decode="polar")
    ref = p24_boxes.postprocess(p, 80, 0.25, 0.45, False)
    differs = 0
    for i in range(4):
        want = orc.postprocess_image(p[i], 80, 0.25, 0.45, False, polar=True)
        assert (got[i] is None) == (want is None)
        if want is not None:
            assert torch.equal(got[i], want), f"image {i}"
            poly = p24_boxes.decode_polygons(got[i])
            assert poly.shape == (got[i].shape[0], 24, 2)
            differs += int(ref[i] is None or ref[i].shape != got[i].shape or not torch.equal(ref[i], got[i]))
    assert differs > 0   # (the corrected decode is a different detector output: that is why it is opt-in)
    with pytest.raises(ValueError):
        p24_boxes.postprocess(p, 80, decode="polygon")
