"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (CPU) in the build
container.  Run from the repo root:  ``python tests/tools/make_golden.py``.

The reference has no tests or golden vectors of its own, so these fixtures —
outputs of the reference itself on seeded inputs — are what pins the oracle
(``oracle/p24_oracle.py``) and, through it, the CUDA path.  While generating, the script also
asserts that the oracle restatement reproduces the reference bit for bit on this host.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)

from ref_loader import cuda0_shim, load_reference  # noqa: E402
from p24 import synth  # noqa: E402
from oracle import p24_oracle as orc  # noqa: E402
from oracle import p24_margins as mg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NC = 80


def regular_gon(cx, cy, r):
    k = np.arange(24, dtype=np.float64) * (15.0 * np.pi / 180.0)
    row = np.zeros(50, dtype=np.float64)
    row[0], row[1] = cx, cy
    row[2::2] = cx + r * np.cos(k)
    row[3::2] = cy + r * np.sin(k)
    return row.astype(np.float32)


def known_answers(ref_models, ref_utils):
    """SURVEY.md §8(c) known-answer table, regenerated from the reference code."""
    cases = [("identical", (100, 100, 20), (100, 100, 20)), ("contained", (100, 100, 20), (105, 100, 10)),
             ("partial_d_eq_r", (100, 100, 20), (120, 100, 20)), ("partial", (100, 100, 30), (125, 100, 10)),
             ("near_tangent_inner", (100, 100, 20), (109.9, 100, 10)), ("tangent_outer", (100, 100, 20), (130, 100, 10)),
             ("disjoint", (100, 100, 20), (200, 100, 10)), ("far", (100, 100, 20), (600, 500, 8))]
    gt = torch.from_numpy(np.stack([regular_gon(*c[1]) for c in cases]))
    pd = torch.tensor([[c[2][0], c[2][1]] + [c[2][2]] * 24 for c in cases], dtype=torch.float32)
    iou = ref_models.IOUloss()
    loss24, _ = iou.forward(pd, gt)
    cx, cy, rg = orc._gt_radii(gt)
    inter, dist = iou.circle_inter(cx, cy, rg, pd[:, 0], pd[:, 1], pd[:, 2:])
    pair = torch.stack([ref_utils.bboxes_iou(gt[i:i + 1], pd[i:i + 1])[0, 0] for i in range(len(cases))])
    o24, _ = orc.iou_loss_forward(pd, gt)
    assert torch.equal(o24, loss24)
    # pts_in_poly probes on the (100,100,r=20) 24-gon
    probes = torch.tensor([[100, 100], [119, 100], [119.9, 100], [114, 114], [120.5, 100], [121, 100],
                           [100, 125], [115, 115]], dtype=torch.float32)
    lf = ref_models.Loss_Function(NC)
    with cuda0_shim("cpu"):
        inside = lf.pts_in_poly(gt[0:1], probes[:, 0], probes[:, 1])
    assert torch.equal(inside, orc.pts_in_poly(gt[0:1], probes[:, 0], probes[:, 1]))
    sums = mg.angle_sums(gt[0:1].numpy(), probes[:, 0].numpy(), probes[:, 1].numpy())
    cxk, cyk = orc.spiral_coefficients()
    np.savez_compressed(os.path.join(GOLD, "known_answers.npz"), names=np.array([c[0] for c in cases]),
                        gt=gt.numpy(), pred=pd.numpy(), loss24=loss24.numpy(), inter=inter.numpy(),
                        dist=dist.numpy(), pair_iou=pair.numpy(), probes=probes.numpy(),
                        inside=inside.numpy(), angle_sums64=sums, coef_x=cxk.numpy(), coef_y=cyk.numpy())
    print("known_answers: loss24 ray0", loss24[:, 0].tolist())


def loss_case(ref_models, name, img_size, counts, max_labels, seed, kind, steps=2, tiny=None,
              radius_range=(0.08, 0.3), need_cert=True):
    B = len(counts)
    out = synth.make_head_outputs(B, img_size, NC, seed=seed)
    lab = synth.make_labels(B, counts, max_labels, img_size, NC, seed=seed, kind=kind, radius_range=radius_range)
    if tiny is not None:  # shrink one GT to a ~1 px polygon (penalised-regime edge case)
        b, g = tiny
        c = lab[b, g, 1:3].clone()
        lab[b, g, 3:] = (lab[b, g, 3:].view(24, 2) - c) .mul(1.0 / 64.0).add(c).view(-1)
    xs, ys, ss = synth.make_grids(img_size)
    ref = ref_models.Loss_Function(NC)
    mine = orc.LossOracle(NC)
    rec = dict(outputs=out.numpy(), labels=lab.numpy(), img_size=np.int64(img_size), steps=np.int64(steps))
    for s in range(steps):
        with cuda0_shim("cpu"):
            r = ref.forward((xs, ys, ss, out.clone(), []), lab)
        o = mine.forward((xs, ys, ss, out.clone(), []), lab)
        for i in range(4):
            assert torch.equal(torch.as_tensor(r[i]), torch.as_tensor(o[i])), (name, s, i)
        assert r[5] == o[5]
        for a, b in zip(r[6], o[6]):
            assert torch.equal(a, b)
        rec[f"s{s}_loss"] = r[0].numpy()
        rec[f"s{s}_loss_iou_w"] = r[1].numpy()
        rec[f"s{s}_loss_obj"] = r[2].numpy()
        rec[f"s{s}_loss_cls"] = r[3].numpy()
        rec[f"s{s}_ratio"] = np.float64(r[5])
        rec[f"s{s}_draw_cx"] = r[6][0].numpy()
        rec[f"s{s}_draw_r"] = r[6][2].numpy()
        rec[f"s{s}_reg_w"] = r[6][3].numpy()
        rec[f"s{s}_obj_w"] = r[6][4].numpy()
        rec[f"s{s}_cls_w"] = r[6][5].numpy()
    # per-image assignment results straight from the reference's get_assignments
    bbox = out[:, :, :26]
    obj = out[:, :, 26].unsqueeze(-1)
    cls = out[:, :, 27:]
    X, Y, S = torch.cat(xs, 1), torch.cat(ys, 1), torch.cat(ss, 1)
    A = out.shape[1]
    fg_all = np.zeros((B, A), dtype=bool)
    matched_all = np.full((B, A), -1, dtype=np.int64)
    iou_all = np.zeros((B, A), dtype=np.float32)
    dynk = np.zeros((B, max_labels), dtype=np.int64)
    ok, res = mg.certify_batch(out, lab, xs, ys, ss, NC)
    if need_cert and not ok:
        return False
    for b in range(B):
        n = counts[b]
        if n == 0:
            continue
        with cuda0_shim("cpu"):
            mcls, fg, miou, midx, nfg = ref.get_assignments(b, n, A, lab[b, :n, 1:], lab[b, :n, 0], bbox[b],
                                                            S, X, Y, cls, bbox, obj)
        tr = mine.trace[b]
        assert torch.equal(fg, tr["fg_mask"]) and torch.equal(midx, tr["matched"]) and torch.equal(miou, tr["ious"])
        fg_all[b] = fg.numpy()
        matched_all[b, fg.numpy()] = midx.numpy()
        iou_all[b, fg.numpy()] = miou.numpy()
        dynk[b, :n] = np.array(tr["dyn_k"])
        if ok:
            assert np.array_equal(res[b]["fg"], fg.numpy()) and np.array_equal(res[b]["matched"], midx.numpy())
            assert list(res[b]["dyn_k"]) == tr["dyn_k"]
    margins = {k: min(r["margins"][k] for r in res if r is not None) for k in mg.DEFAULT_THRESHOLDS} if any(
        r is not None for r in res) else {}
    rec.update(fg_mask=fg_all, matched_gt=matched_all, pred_iou=iou_all, dyn_k=dynk,
               counts=np.array(counts, dtype=np.int64), certified=np.bool_(ok),
               margins=np.array([margins.get(k, np.inf) for k in mg.DEFAULT_THRESHOLDS]))
    np.savez_compressed(os.path.join(GOLD, f"loss_{name}.npz"), **rec)
    print(f"loss_{name}: certified={ok} num_fg={fg_all.sum(1).tolist()} loss={float(rec['s0_loss']):.6f} margins={margins}")
    return True


def post_case(ref_utils, name, img_size, B, seed, settings):
    p = synth.make_postprocess_input(B, img_size, NC, seed=seed)
    rec = dict(prediction=p.numpy(), img_size=np.int64(img_size))
    for si, (c, n, ag) in enumerate(settings):
        rec[f"cfg{si}"] = np.array([c, n, float(ag)])
        for i in range(B):
            r = ref_utils.postprocess(p[i:i + 1].clone(), NC, c, n, ag)[0]
            o = orc.postprocess(p[i:i + 1].clone(), NC, c, n, ag)[0]
            assert (r is None) == (o is None) and (r is None or torch.equal(r, o))
            rec[f"cfg{si}_img{i}"] = np.zeros((0, 29), np.float32) if r is None else r.numpy()
        print(f"post_{name} cfg{si}: n={[rec[f'cfg{si}_img{i}'].shape[0] for i in range(B)]}")
    np.savez_compressed(os.path.join(GOLD, f"post_{name}.npz"), **rec)


def head_case(ref_models, name, img_size, B, seed):
    """Head decode (yolo_head_24p.py:212-256): raw per-level conv outputs -> the reference's decoded training buffer and
    inference prediction, by the reference's own methods called on a stand-in for the module state they read."""
    import types
    head = ref_models.YOLOXHead
    reg, obj, cls = synth.make_raw_levels(B, img_size, NC, seed=seed)
    strides = list(synth.STRIDES)
    me = types.SimpleNamespace(grids=[torch.zeros(1)] * len(strides), num_classes=NC, n_anchors=1)
    outs = [head.get_output_and_grid(me, torch.cat([reg[k], obj[k], cls[k]], 1), k, strides[k], "torch.FloatTensor")[0]
            for k in range(len(strides))]
    train = torch.cat(outs, 1)
    flat = [torch.cat([reg[k], obj[k].sigmoid(), cls[k].sigmoid()], 1) for k in range(len(strides))]
    me.hw, me.strides = [x.shape[-2:] for x in flat], strides
    infer = head.decode_outputs(me, torch.cat([x.flatten(start_dim=2) for x in flat], dim=2).permute(0, 2, 1),
                                dtype="torch.FloatTensor")
    assert torch.equal(orc.head_decode_train(reg, obj, cls, strides)[3], train)
    assert torch.equal(orc.head_decode_infer(reg, obj, cls, strides), infer)
    np.savez_compressed(os.path.join(GOLD, f"head_{name}.npz"), img_size=np.int64(img_size), batch=np.int64(B),
                        seed=np.int64(seed), train=train.numpy(), infer=infer.numpy())
    print(f"head_{name}: train {tuple(train.shape)} infer {tuple(infer.shape)}")


def pack_case():
    """Label packing (datasets/data_augment.py:131-174): the reference's TrainTransform on seeded ragged targets."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data_augment", "/root/reference/yolox_24p/datasets/data_augment.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    cases = [(7, (480, 640)), (60, (640, 427)), (1, (333, 500)), (0, (640, 640)), (50, (640, 640))]
    rec = {"counts": np.array([c[0] for c in cases]), "shapes": np.array([c[1] for c in cases])}
    flat = []
    for i, (n, (h, w)) in enumerate(cases):
        t = np.concatenate([rng.integers(0, 80, (n, 1)).astype(np.float64), rng.random((n, 50))], 1) if n else np.zeros((1, 0))
        _, want = mod.TrainTransform(max_labels=50)(np.zeros((h, w, 3), dtype=np.uint8), t.copy(), [640, 640])
        assert np.array_equal(orc.pack_labels(t, (h, w), (640, 640), 50), want)
        rec[f"labels{i}"] = want
        flat.append(t if n else np.zeros((0, 51)))
    rec["targets"] = np.concatenate(flat, 0)
    np.savez_compressed(os.path.join(GOLD, "pack_labels.npz"), **rec)
    print("pack_labels:", rec["counts"].tolist())


def main():
    if sys.argv[1:] == ["pack"]:
        pack_case()
        return
    if sys.argv[1:] == ["head"]:   # only the head-decode fixture (the others stay byte-identical)
        head_case(load_reference()[0], "s64", 64, 2, 21)
        return
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    ref_models, ref_utils = load_reference()
    known_answers(ref_models, ref_utils)
    # small (256 px -> 1344 anchors) certified cases; seeds searched until margin-safe
    for name, counts, kind, tiny in [("smooth_s256", [6, 5], "smooth", None), ("spiky_s256", [6, 8], "spiky", None),
                                     ("mixed_empty_s256", [4, 0, 3], "smooth", None),
                                     ("tiny_gt_s256", [3, 2], "smooth", (0, 1))]:
        for seed in range(100, 140):
            if loss_case(ref_models, name, 256, counts, 10, seed, kind, tiny=tiny, need_cert=tiny is None):
                break
        else:
            raise SystemExit(f"no certified seed for {name}")
    post_case(ref_utils, "s256", 256, 2, 3,
              [(0.25, 0.45, False), (0.01, 0.65, False), (0.01, 0.3, True), (0.99, 0.45, False)])
    head_case(ref_models, "s64", 64, 2, 21)
    pack_case()


if __name__ == "__main__":
    main()
