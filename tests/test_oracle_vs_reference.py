"""Build-container-only check: the oracle restatement reproduces the UNMODIFIED reference bit for
bit (same torch build, same CPU).  Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import pytest
import torch

from ref_loader import cuda0_shim, load_reference, reference_available
from oracle import p24_oracle as orc
from p24 import synth

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("kind,seed", [("smooth", 11), ("spiky", 12)])
def test_loss_forward_bit_equal(kind, seed):
    ref_models, _ = load_reference()
    out = synth.make_head_outputs(2, 320, 80, seed=seed)
    lab = synth.make_labels(2, [7, 3], 10, 320, 80, seed=seed, kind=kind)
    xs, ys, ss = synth.make_grids(320)
    ref, mine = ref_models.Loss_Function(80), orc.LossOracle(80)
    for _ in range(2):
        with cuda0_shim("cpu"):
            r = ref.forward((xs, ys, ss, out.clone(), []), lab)
        o = mine.forward((xs, ys, ss, out.clone(), []), lab)
        for i in range(4):
            assert torch.equal(r[i], o[i])
        assert r[4] == o[4] and r[5] == o[5]
        assert all(torch.equal(a, b) for a, b in zip(r[6], o[6]))


def test_pairwise_and_matched_bit_equal():
    ref_models, ref_utils = load_reference()
    lab = synth.make_labels(1, 9, 10, 320, 80, seed=5, kind="spiky")[0, :9, 1:]
    pred = synth.make_head_outputs(1, 320, 80, seed=5)[0, :700, :26]
    assert torch.equal(ref_utils.bboxes_iou(lab, pred), orc.bboxes_iou(lab, pred))
    a, _ = ref_models.IOUloss().forward(pred[:9], lab)
    b, _ = orc.iou_loss_forward(pred[:9], lab)
    assert torch.equal(a, b)
    with pytest.raises(IndexError):
        orc.iou_loss_forward(pred[:, :25], lab)
    e, d = orc.iou_loss_forward(pred[:0], lab[:0])
    er, dr = ref_models.IOUloss().forward(pred[:0], lab[:0])
    assert torch.equal(e, er) and e.shape == (1, 24)


def test_postprocess_bit_equal():
    _, ref_utils = load_reference()
    p = synth.make_postprocess_input(2, 320, 80, seed=9)
    for c, n, ag in [(0.25, 0.45, False), (0.01, 0.3, True)]:
        for i in range(2):
            r = ref_utils.postprocess(p[i:i + 1].clone(), 80, c, n, ag)[0]
            o = orc.postprocess(p[i:i + 1].clone(), 80, c, n, ag)[0]
            assert (r is None) == (o is None) and (r is None or torch.equal(r, o))


def test_head_decode_bit_equal():
    """The head-decode restatement (train and inference) against the reference's own methods
    (yolo_head_24p.py:212-256), called unbound on a stand-in for the module state they read."""
    import types
    ref_models, _ = load_reference()
    head = ref_models.YOLOXHead
    reg, obj, cls = synth.make_raw_levels(2, 160, 80, seed=21)
    strides = [8, 16, 32]
    me = types.SimpleNamespace(grids=[torch.zeros(1)] * 3, num_classes=80, n_anchors=1)
    outs, xs, ys = [], [], []
    for k in range(3):
        o, grid = head.get_output_and_grid(me, torch.cat([reg[k], obj[k], cls[k]], 1), k, strides[k], "torch.FloatTensor")
        outs.append(o), xs.append(grid[:, :, 0]), ys.append(grid[:, :, 1])
    want = torch.cat(outs, 1)
    gx, gy, gs, got = orc.head_decode_train(reg, obj, cls, strides)
    assert torch.equal(got, want)
    assert all(torch.equal(a, b) for a, b in zip(gx, xs)) and all(torch.equal(a, b) for a, b in zip(gy, ys))
    assert all(float(s[0, 0]) == st and s.shape[1] == x.shape[1] for s, st, x in zip(gs, strides, xs))
    # inference: cat(sigmoid) -> flatten/cat/permute -> decode_outputs
    flat = [torch.cat([reg[k], obj[k].sigmoid(), cls[k].sigmoid()], 1) for k in range(3)]
    me.hw, me.strides = [x.shape[-2:] for x in flat], strides
    pred = torch.cat([x.flatten(start_dim=2) for x in flat], dim=2).permute(0, 2, 1)
    want = head.decode_outputs(me, pred, dtype="torch.FloatTensor")
    assert torch.equal(orc.head_decode_infer(reg, obj, cls, strides), want)


def _ref_train_transform():
    """The reference's datasets/data_augment.py imported as a stand-alone file (the package __init__ needs pycocotools)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data_augment", "/root/reference/yolox_24p/datasets/data_augment.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TrainTransform


def test_label_packing_bit_equal():
    """oracle.pack_labels against TrainTransform.__call__ of the reference (datasets/data_augment.py:131-174): ragged
    counts, more targets than max_labels, an empty label file, non-square images."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")  # the reference resizes the image in the same call
    TT = _ref_train_transform()
    rng = np.random.default_rng(5)
    for n, (h, w), max_labels in [(7, (480, 640), 50), (60, (640, 427), 50), (1, (333, 500), 10), (0, (640, 640), 50)]:
        img = np.zeros((h, w, 3), dtype=np.uint8)
        if n:
            t = np.concatenate([rng.integers(0, 80, (n, 1)).astype(np.float64), rng.random((n, 50))], 1)
        else:
            t = np.zeros((0,))[np.newaxis, :]   # what pull_item yields for an empty label file (coco24p.py:84-85)
        _, want = TT(max_labels=max_labels)(img, t.copy(), [640, 640])
        got = orc.pack_labels(t, (h, w), (640, 640), max_labels)
        assert want.dtype == got.dtype and np.array_equal(want, got)


def test_staged_reference_tree_is_the_unmodified_reference(tmp_path, monkeypatch):
    """oracle/make_ref.py (the recipe behind bench.py's `kind: "reference"` CPU arm): the staged files are byte-identical
    to the reference's, the manifest says so, and the tree loaded through oracle/ref_runtime.py gives the oracle's bits."""
    import hashlib
    import json
    import os
    import subprocess
    import sys
    from oracle import make_ref
    monkeypatch.setattr(make_ref, "DEST", str(tmp_path / "yolox_24p"))
    assert make_ref.make("/root/reference")
    man = json.load(open(tmp_path / "yolox_24p" / "MANIFEST.json"))
    for f in make_ref.FILES:
        a = open(tmp_path / "yolox_24p" / f, "rb").read()
        b = open(os.path.join("/root/reference/yolox_24p", f), "rb").read()
        assert a == b and hashlib.sha256(a).hexdigest() == man["sha256"][f]
    # load it in a fresh interpreter (this one already holds the reference's top-level `models` / `utils` modules)
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'exploration-of-potential_b200')!r})\n"
        "from oracle import ref_runtime, p24_oracle as orc\n"
        "from p24 import synth\n"
        f"models, utils = ref_runtime.load({str(tmp_path / 'yolox_24p')!r})\n"
        "out = synth.make_head_outputs(1, 160, 80, seed=2); lab = synth.make_labels(1, [4], 6, 160, 80, seed=2, kind='smooth')\n"
        "xs, ys, ss = synth.make_grids(160)\n"
        "with ref_runtime.cuda0_shim(torch.device('cpu')):\n"
        "    r = models.Loss_Function(80).forward((xs, ys, ss, out.clone(), []), lab)\n"
        "o = orc.LossOracle(80).forward((xs, ys, ss, out.clone(), []), lab)\n"
        "assert all(torch.equal(r[i], o[i]) for i in range(4))\n"
        "p = synth.make_postprocess_input(1, 160, 80, seed=4)\n"
        "a, b = utils.postprocess(p.clone(), 80, 0.01, 0.5, False)[0], orc.postprocess(p.clone(), 80, 0.01, 0.5, False)[0]\n"
        "assert (a is None) == (b is None) and (a is None or torch.equal(a, b))\n"
        "print('STAGED_OK')\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "STAGED_OK" in res.stdout, res.stderr[-2000:]
