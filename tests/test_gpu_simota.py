"""GPU parity tests of the fused SimOTA + loss path (through the C ABI) against the oracle.

Bar (BASELINE.json): matched-anchor indices, foreground masks and dynamic-k counts bit-exact;
pair values / losses within 1e-5 relative (fp32).  The oracle (torch restatement of the reference,
pinned to the reference by tests/golden + tests/test_oracle_vs_reference.py) runs on the same GPU,
and the goldens produced by the UNMODIFIED reference on CPU are checked as well (bit-exact
decisions where the float64 margin certifier proved the input margin-safe).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import p24_oracle as orc
from p24 import engine as eng
from p24 import synth
from p24.losses import Loss_Function

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5  # BASELINE.json: IoU / loss values within 1e-5 relative (fp32)
DEV = "cuda:0"


def _to_dev(lst):
    return [t.to(DEV) for t in lst]


def _grids(size):
    xs, ys, ss = synth.make_grids(size)
    return _to_dev(xs), _to_dev(ys), _to_dev(ss)


def _assert_assignment_equal(asg, trace):
    for b, tr in enumerate(trace):
        fg_o = tr["fg_mask"].cpu().numpy()
        fg_m = asg.fg_mask[b].bool().cpu().numpy()
        assert np.array_equal(fg_o, fg_m), f"image {b}: fg_mask differs at {np.nonzero(fg_o != fg_m)[0][:10]}"
        assert int(asg.num_fg[b]) == int(fg_o.sum())
        assert int(asg.num_gt[b]) == tr["num_gt"]
        mm = asg.matched_gt[b].cpu().numpy()
        assert (mm[~fg_m] == -1).all()
        if tr["num_gt"]:
            assert np.array_equal(mm[fg_m], tr["matched"].cpu().numpy()), f"image {b}: matched GT indices differ"
            assert asg.dyn_k[b, :tr["num_gt"]].cpu().tolist() == list(tr["dyn_k"]), f"image {b}: dynamic k differs"
            np.testing.assert_allclose(asg.pred_iou[b].cpu().numpy()[fg_m], tr["ious"].cpu().numpy(), rtol=RTOL)
        assert (asg.dyn_k[b, tr["num_gt"]:] == 0).all()
        assert (asg.pred_iou[b].cpu().numpy()[~fg_m] == 0).all()


def _run_both(out, lab, size, steps=2, flags=0):
    gx, gy, gs = _grids(size)
    outd, labd = out.to(DEV), lab.to(DEV)
    mine, o = Loss_Function(80), orc.LossOracle(80)
    for _ in range(steps):
        r = o.forward((gx, gy, gs, outd.clone(), []), labd)
        res, w, asg = mine.forward_async((gx, gy, gs, outd, []), labd, flags=flags)
        _assert_assignment_equal(asg, o.trace)
        np.testing.assert_allclose(float(res[0]), float(r[0]), rtol=RTOL)
        np.testing.assert_allclose(res[1:25].cpu().numpy(), r[1].cpu().numpy(), rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(float(res[25]), float(r[2]), rtol=RTOL)
        np.testing.assert_allclose(float(res[26]), float(r[3]), rtol=RTOL, atol=1e-7)
        assert float(res[27]) == pytest.approx(r[5], rel=1e-6)
        np.testing.assert_allclose(res[28:52].cpu().numpy(), r[6][3].cpu().numpy(), rtol=RTOL)
    return mine, o


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "loss_*.npz"))), ids=os.path.basename)
def test_golden_fixture_from_reference(path):
    """Fixtures produced by the UNMODIFIED reference on CPU (tests/tools/make_golden.py)."""
    g = np.load(path)
    out, lab = torch.from_numpy(g["outputs"]).to(DEV), torch.from_numpy(g["labels"]).to(DEV)
    gx, gy, gs = _grids(int(g["img_size"]))
    lf = Loss_Function(80)
    for s in range(int(g["steps"])):
        r = lf.forward((gx, gy, gs, out, []), lab)
        np.testing.assert_allclose(float(r[0]), g[f"s{s}_loss"], rtol=RTOL)
        np.testing.assert_allclose(r[1].cpu().numpy(), g[f"s{s}_loss_iou_w"], rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(float(r[2]), g[f"s{s}_loss_obj"], rtol=RTOL)
        np.testing.assert_allclose(float(r[3]), g[f"s{s}_loss_cls"], rtol=RTOL, atol=1e-7)
        assert r[4] == 0.0 and r[5] == pytest.approx(float(g[f"s{s}_ratio"]), rel=1e-6)
        np.testing.assert_allclose(r[6][3].cpu().numpy(), g[f"s{s}_reg_w"], rtol=RTOL)
        if bool(g["certified"]):
            assert np.array_equal(r[6][0].cpu().numpy().reshape(-1), g[f"s{s}_draw_cx"].reshape(-1))
    asg = lf.last_assignment
    if bool(g["certified"]):
        assert np.array_equal(asg.fg_mask.bool().cpu().numpy(), g["fg_mask"])
        assert np.array_equal(asg.matched_gt.cpu().numpy(), g["matched_gt"])
        assert np.array_equal(asg.dyn_k.cpu().numpy(), g["dyn_k"])
        np.testing.assert_allclose(asg.pred_iou.cpu().numpy(), g["pred_iou"], rtol=RTOL)


@pytest.mark.parametrize("kind,seed", [("smooth", 0), ("spiky", 1)])
def test_config1_batch2_640_vs_oracle_on_gpu(kind, seed):
    """BASELINE.json configs[0]: B=2, 640x640 (8400 anchors), 20 GT/img."""
    out = synth.make_head_outputs(2, 640, 80, seed=seed)
    lab = synth.make_labels(2, 20, 50, 640, 80, seed=seed, kind=kind)
    _run_both(out, lab, 640)


def test_pruning_and_filter_are_decision_preserving():
    """The geometric pruning of the polygon test and the bound filter of the top-10 selection must not
    change any output bit (P24_F_NO_PRUNE | P24_F_NO_FILTER evaluates everything exactly)."""
    gx, gy, gs = _grids(640)
    for kind, seed, n in [("smooth", 3, 20), ("spiky", 4, 20), ("spiky", 5, 3), ("smooth", 6, 1)]:
        out = synth.make_head_outputs(3, 640, 80, seed=seed).to(DEV)
        lab = synth.make_labels(3, n, 50, 640, 80, seed=seed, kind=kind).to(DEV)
        a = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab)
        b = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab, flags=eng.F_NO_PRUNE | eng.F_NO_FILTER)
        for name in ("fg_mask", "matched_gt", "pred_iou", "num_fg", "dyn_k", "sums28"):
            assert torch.equal(getattr(a[2], name), getattr(b[2], name)), (kind, seed, name)
        assert torch.equal(a[0], b[0])


def test_launch_modes_give_identical_bits():
    """Programmatic dependent launch is a scheduling choice: same output bits as plain stream order."""
    gx, gy, gs = _grids(640)
    out = synth.make_head_outputs(20, 640, 80, seed=8).to(DEV)
    lab = synth.make_labels(20, 20, 50, 640, 80, seed=8, kind="smooth").to(DEV)
    ref = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab)
    for flags in (eng.F_NO_PDL,):
        got = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab, flags=flags)
        for name in ("fg_mask", "matched_gt", "pred_iou", "num_fg", "num_gt", "dyn_k", "sums28"):
            assert torch.equal(getattr(ref[2], name), getattr(got[2], name)), (flags, name)
        assert torch.equal(ref[0], got[0])


def test_edge_cases_empty_images_single_gt_border_and_duplicates():
    size = 320
    out = synth.make_head_outputs(5, size, 80, seed=21)
    lab = synth.make_labels(5, [0, 1, 4, 0, 2], 8, size, 80, seed=21, kind="smooth")
    # image 2: a duplicated GT (exact cost ties between two GT rows) and a GT hanging over the image border
    lab[2, 1] = lab[2, 0]
    shift = lab[2, 2, 1].item() - 4.0
    lab[2, 2, 1::2] -= shift
    _run_both(out, lab, size)
    # all-background batch: num_fg = max(0, 1), loss_iou = 0, ratio = 1.0
    lab0 = torch.zeros(2, 8, 51)
    _run_both(out[:2], lab0, size, steps=1)
    gx, gy, gs = _grids(size)
    r = Loss_Function(80).forward((gx, gy, gs, out[:2].to(DEV), []), lab0.to(DEV))
    assert r[5] == 1.0 and float(r[1].abs().sum()) == 0.0 and r[6][0].shape == (1, 24)


def test_tiny_gt_spills_into_penalised_regime():
    """1-px GT: no valid anchor, every match comes from the 1e5-penalised regime."""
    size = 256
    out = synth.make_head_outputs(2, size, 80, seed=31)
    lab = synth.make_labels(2, [3, 2], 10, size, 80, seed=31, kind="smooth", radius_range=(0.08, 0.3))
    c = lab[0, 1, 1:3].clone()
    lab[0, 1, 3:] = (lab[0, 1, 3:].view(24, 2) - c).mul(1.0 / 64.0).add(c).view(-1)
    _run_both(out, lab, size, steps=1)


def test_crowded_100_gt_vs_oracle_on_gpu():
    """BASELINE.json configs[2] shape at a reduced batch: 100 GT/img (dynamic-k conflicts)."""
    out = synth.make_head_outputs(2, 640, 80, seed=2)
    lab = synth.make_labels(2, 100, 100, 640, 80, seed=2, kind="smooth")
    _run_both(out, lab, 640, steps=1)


def test_get_assignments_drop_in_signature():
    size = 320
    out = synth.make_head_outputs(2, size, 80, seed=41).to(DEV)
    lab = synth.make_labels(2, [5, 3], 10, size, 80, seed=41, kind="smooth").to(DEV)
    xs, ys, ss = synth.make_grids(size)
    X, Y, S = torch.cat(xs, 1).to(DEV), torch.cat(ys, 1).to(DEV), torch.cat(ss, 1).to(DEV)
    bbox, obj, cls = out[:, :, :26], out[:, :, 26].unsqueeze(-1), out[:, :, 27:]
    lf = Loss_Function(80)
    for b, n in enumerate([5, 3]):
        gt50, gcls = lab[b, :n, 1:], lab[b, :n, 0]
        want = orc.get_assignments(n, out.shape[1], gt50, gcls, bbox[b], S, X, Y, cls[b], obj[b], 80)
        got = lf.get_assignments(b, n, out.shape[1], gt50, gcls, bbox[b], S, X, Y, cls, bbox, obj)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[3], want[3])
        assert got[4] == want[4] and got[3].dtype == torch.int64 and got[1].dtype == torch.bool
        torch.testing.assert_close(got[2], want[2], rtol=RTOL, atol=0)


def test_full_size_batch20_properties():
    """BASELINE.json configs[1] at full size (B=20): size-independent properties — per-image independence
    (a batch result equals the per-image results), determinism, and count consistency."""
    out = synth.make_head_outputs(20, 640, 80, seed=1).to(DEV)
    lab = synth.make_labels(20, 20, 50, 640, 80, seed=1, kind="smooth").to(DEV)
    gx, gy, gs = _grids(640)
    res, _, a = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab)
    res2, _, a2 = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab)
    assert torch.equal(res, res2) and torch.equal(a.fg_mask, a2.fg_mask) and torch.equal(a.matched_gt, a2.matched_gt)
    assert int(a.sums28[26]) == int(a.num_fg.sum()) == int(a.fg_mask.sum()) and int(a.sums28[27]) == 400
    assert ((a.matched_gt >= 0) == a.fg_mask.bool()).all() and int(a.matched_gt.max()) < 20
    assert (a.dyn_k[:, :20] >= 1).all() and (a.dyn_k[:, :20] <= 10).all() and (a.dyn_k[:, 20:] == 0).all()
    for b in (0, 7, 19):
        _, _, s = Loss_Function(80).forward_async((gx, gy, gs, out[b:b + 1], []), lab[b:b + 1])
        assert torch.equal(s.fg_mask[0], a.fg_mask[b]) and torch.equal(s.matched_gt[0], a.matched_gt[b])
        assert torch.equal(s.pred_iou[0], a.pred_iou[b]) and torch.equal(s.dyn_k[0], a.dyn_k[b])


def test_config5_highres_1280_vs_oracle_on_gpu():
    """BASELINE.json configs[4] shape (1280x1280, 33600 anchors, 20 GT/img) at a reduced batch."""
    out = synth.make_head_outputs(2, 1280, 80, seed=4)
    lab = synth.make_labels(2, 20, 50, 1280, 80, seed=4, kind="smooth")
    _run_both(out, lab, 1280, steps=1)


def test_crowded_full_batch_properties():
    """BASELINE.json configs[2] at full size (B=20, 100 GT/img): per-image independence and count consistency."""
    out = synth.make_head_outputs(20, 640, 80, seed=2).to(DEV)
    lab = synth.make_labels(20, 100, 100, 640, 80, seed=2, kind="smooth").to(DEV)
    gx, gy, gs = _grids(640)
    res, _, a = Loss_Function(80).forward_async((gx, gy, gs, out, []), lab)
    assert int(a.sums28[26]) == int(a.num_fg.sum()) == int(a.fg_mask.sum()) and int(a.sums28[27]) == 2000
    assert ((a.matched_gt >= 0) == a.fg_mask.bool()).all() and int(a.matched_gt.max()) < 100
    assert (a.dyn_k >= 1).all() and (a.dyn_k <= 10).all()
    # every GT keeps at most dyn_k anchors after conflict resolution
    for b in (0, 13):
        counts = torch.bincount(a.matched_gt[b][a.fg_mask[b].bool()].long(), minlength=100)
        assert (counts <= a.dyn_k[b].long()).all()
        _, _, s = Loss_Function(80).forward_async((gx, gy, gs, out[b:b + 1], []), lab[b:b + 1])
        assert torch.equal(s.fg_mask[0], a.fg_mask[b]) and torch.equal(s.matched_gt[0], a.matched_gt[b])
        assert torch.equal(s.dyn_k[0], a.dyn_k[b])
    assert torch.isfinite(res).all()


def test_mixed_label_kinds_many_seeds_vs_oracle_on_gpu():
    """Several seeds / label kinds / GT counts at 640x640 against the oracle on the same GPU (decisions bit-exact)."""
    for seed, kind, n in [(51, "smooth", 7), (52, "spiky", 13), (53, "smooth", 33), (54, "spiky", 2)]:
        out = synth.make_head_outputs(2, 640, 80, seed=seed)
        lab = synth.make_labels(2, [n, max(n // 2, 1)], 50, 640, 80, seed=seed, kind=kind)
        _run_both(out, lab, 640, steps=1)


def test_two_gpu_fused_allreduce_matches_full_batch():
    """Images sharded over 2 GPUs, the 28 sums all-reduced inside the last kernel over peer memory (and, for comparison,
    with NCCL): same loss as the whole batch on one GPU, bit-identical on both ranks.  Needs 2 GPUs."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(root, "tests", "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
