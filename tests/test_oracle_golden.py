"""Oracle (oracle/p24_oracle.py) against the golden fixtures produced by the UNMODIFIED reference
(tests/tools/make_golden.py).  CPU only.  Discrete outputs must be identical; floats may differ in
the last bits between hosts (vectorised ATen reductions / SLEEF variants depend on the CPU's ISA)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 2e-6


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_known_answer_table():
    g = _load("known_answers.npz")
    gt, pd = torch.from_numpy(g["gt"]), torch.from_numpy(g["pred"])
    loss24, draw = orc.iou_loss_forward(pd, gt)
    np.testing.assert_allclose(loss24.numpy(), g["loss24"], rtol=RTOL, atol=1e-6)
    cx, cy, rg = orc._gt_radii(gt)
    inter, dist = orc.circle_inter_matched(cx, cy, rg, pd[:, 0], pd[:, 1], pd[:, 2:])
    np.testing.assert_allclose(inter.numpy(), g["inter"], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(dist.numpy(), g["dist"], rtol=RTOL)
    pair = torch.stack([orc.bboxes_iou(gt[i:i + 1], pd[i:i + 1])[0, 0] for i in range(gt.shape[0])])
    np.testing.assert_allclose(pair.numpy(), g["pair_iou"], rtol=RTOL, atol=1e-7)
    # the survey's table, ray 0: identical, contained, partial d=r, partial, ...
    expect = [0.0, 0.75, 1.04188001, 1.04254508, 0.75, 1.44444442, 1.88165689, 1.9958446]
    np.testing.assert_allclose(loss24[:, 0].numpy(), expect, rtol=1e-6, atol=1e-6)
    inside = orc.pts_in_poly(gt[0:1], torch.from_numpy(g["probes"][:, 0]), torch.from_numpy(g["probes"][:, 1]))
    assert inside.numpy().tolist() == g["inside"].tolist() == [[True, True, True, True, False, False, False, False]]
    cxk, cyk = orc.spiral_coefficients()
    np.testing.assert_allclose(cxk.numpy(), g["coef_x"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(cxk[:6].numpy(), [0, 0.25287879, 0.45344985, 0.55536038, 0.52359873, 0.33879337],
                               rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cyk[:6].numpy(), [0, 0.067758672, 0.26179940, 0.55536038, 0.90689969, 1.2643939],
                               rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "loss_*.npz"))), ids=os.path.basename)
def test_loss_forward_matches_reference_fixture(path):
    g = np.load(path)
    out, lab = torch.from_numpy(g["outputs"]), torch.from_numpy(g["labels"])
    xs, ys, ss = synth.make_grids(int(g["img_size"]))
    o = orc.LossOracle(80)
    for s in range(int(g["steps"])):
        r = o.forward((xs, ys, ss, out.clone(), []), lab)
        np.testing.assert_allclose(r[0].numpy(), g[f"s{s}_loss"], rtol=1e-5)
        np.testing.assert_allclose(r[1].numpy(), g[f"s{s}_loss_iou_w"], rtol=1e-5)
        np.testing.assert_allclose(r[2].numpy(), g[f"s{s}_loss_obj"], rtol=1e-5)
        np.testing.assert_allclose(r[3].numpy(), g[f"s{s}_loss_cls"], rtol=1e-5)
        assert r[4] == 0.0 and r[5] == float(g[f"s{s}_ratio"])
        np.testing.assert_allclose(r[6][3].numpy(), g[f"s{s}_reg_w"], rtol=1e-5)
    if bool(g["certified"]):
        for b, tr in enumerate(o.trace):
            fg = tr["fg_mask"].numpy()
            assert np.array_equal(fg, g["fg_mask"][b])
            if tr["num_gt"]:
                assert np.array_equal(tr["matched"].numpy(), g["matched_gt"][b][fg])
                assert tr["dyn_k"] == g["dyn_k"][b][:tr["num_gt"]].tolist()
                np.testing.assert_allclose(tr["ious"].numpy(), g["pred_iou"][b][fg], rtol=RTOL)


def test_postprocess_matches_reference_fixture():
    g = _load("post_s256.npz")
    p = torch.from_numpy(g["prediction"])
    si = 0
    while f"cfg{si}" in g.files:
        c, n, ag = g[f"cfg{si}"].tolist()
        res = orc.postprocess(p.clone(), 80, c, n, bool(ag))
        for i, r in enumerate(res):
            want = g[f"cfg{si}_img{i}"]
            if want.shape[0] == 0:
                assert r is None
            else:
                assert r.shape == want.shape
                # kept rows are gathered copies of the input: bit-exact, order included
                assert np.array_equal(r.numpy(), want)
        si += 1
    assert si == 4


def test_head_decode_matches_reference_fixture():
    """Head decode restatement (yolo_head_24p.py:212-256) against the reference-made fixture: raw conv outputs are
    re-drawn from the seed; exp() may differ in the last bit between hosts."""
    g = _load("head_s64.npz")
    reg, obj, cls = synth.make_raw_levels(int(g["batch"]), int(g["img_size"]), 80, seed=int(g["seed"]))
    xs, ys, ss, train = orc.head_decode_train(reg, obj, cls, list(synth.STRIDES))
    np.testing.assert_allclose(train.numpy(), g["train"], rtol=RTOL, atol=0)
    assert np.array_equal(train[:, :, 26:].numpy(), g["train"][:, :, 26:])   # logits pass through untouched
    mx, my, ms = synth.make_grids(int(g["img_size"]))
    assert all(torch.equal(a, b) for a, b in zip(xs + ys + ss, mx + my + ms))
    np.testing.assert_allclose(orc.head_decode_infer(reg, obj, cls, list(synth.STRIDES)).numpy(), g["infer"], rtol=RTOL, atol=0)


def test_label_packing_matches_reference_fixture():
    """oracle.pack_labels against the reference-made fixture (TrainTransform, datasets/data_augment.py:131-174):
    float64 arithmetic cast once to fp32 -> bit-exact on every host."""
    g = _load("pack_labels.npz")
    off = 0
    for i, (n, hw) in enumerate(zip(g["counts"].tolist(), g["shapes"].tolist())):
        t = g["targets"][off:off + n] if n else np.zeros((1, 0))
        off += n
        assert np.array_equal(orc.pack_labels(t, tuple(hw), (640, 640), 50), g[f"labels{i}"])
