"""ORACLE (test infrastructure) — float64 numpy restatement of the SimOTA decisions, used to
measure how far every discrete decision of one input is from its threshold.

A fp32 implementation whose value errors are far below a decision's margin MUST take the same
decision as the reference; the parity tests therefore certify their seeded inputs as
"margin-safe" with this module before demanding bit-exact matched indices / dynamic-k counts
(SURVEY.md §7 hard part 1).  Formulas follow SURVEY.md Appendix A; reference lines are cited
per function (paths relative to ``/root/reference/yolox_24p``).

Only ``tests/`` (and ``bench.py``'s optional certification print) may import this file.
"""
from __future__ import annotations

import numpy as np

PI32 = float(np.float32(np.pi))           # torch.tensor(np.pi)
DEG32 = float(np.float32(180.0 / np.pi))  # rad2deg multiplier as fp32


def gt_geometry(gt50):
    """cx, cy, vertex arrays and per-ray radii (losses.py:91-108)."""
    gt50 = np.asarray(gt50, dtype=np.float64)
    cx, cy = gt50[:, 0], gt50[:, 1]
    vx, vy = gt50[:, 2::2], gt50[:, 3::2]
    rg = np.sqrt((vx - cx[:, None]) ** 2 + (vy - cy[:, None]) ** 2)
    return cx, cy, vx, vy, rg


def angle_sums(gt50, xc, yc):
    """Total unsigned angular variation in degrees, [G, A] (losses.py:566-588)."""
    _, _, vx, vy = gt_geometry(gt50)[:4]
    xc = np.asarray(xc, np.float64)
    yc = np.asarray(yc, np.float64)
    G = vx.shape[0]
    out = np.zeros((G, xc.shape[0]))
    for g in range(G):
        sx = vx[g][:, None] - xc[None, :]
        sy = vy[g][:, None] - yc[None, :]
        ex = np.roll(sx, -1, 0)
        ey = np.roll(sy, -1, 0)
        out[g] = (np.arctan2(np.abs(sx * ey - ex * sy), sx * ex + sy * ey) * DEG32).sum(0)
    return out


def center_deltas(gt50, xc, yc, st):
    """min over the four window distances, [G, A] (losses.py:523-542)."""
    gt50 = np.asarray(gt50, np.float64)
    gx, gy = gt50[:, 0:1], gt50[:, 1:2]
    r = 2.5 * st[None, :]
    return np.minimum(np.minimum(xc[None] - (gx - r), yc[None] - (gy - r)),
                      np.minimum((gx + r) - xc[None], (gy + r) - yc[None]))


def pair_values(gt50, pred26):
    """bboxes_iou in float64, [G, P] (boxes.py:166-243) + minimum branch margin."""
    cx, cy, _, _, rg = gt_geometry(gt50)
    pred26 = np.asarray(pred26, np.float64)
    G, P = rg.shape[0], pred26.shape[0]
    out = np.zeros((G, P))
    bmargin = np.inf
    for g in range(G):
        d = np.sqrt((cx[g] - pred26[:, 0]) ** 2 + (cy[g] - pred26[:, 1]) ** 2)[:, None]
        rgk = rg[g][None, :]
        rp = pred26[:, 2:]
        rmin, rmax = np.minimum(rgk, rp), np.maximum(rgk, rp)
        ac_min = np.clip((rmin**2 + d**2 - rmax**2) / (2 * rmin * d + 1e-8), -0.99, 0.99)
        ac_max = np.clip((rmax**2 + d**2 - rmin**2) / (2 * rmax * d + 1e-8), -0.99, 0.99)
        a_min, a_max = np.arccos(ac_min), np.arccos(ac_max)
        lens = a_min * rmin**2 + a_max * rmax**2 - rmin * d * np.sin(a_min)
        nested = np.abs(rgk - rp) >= d
        apart = d >= rgk + rp
        inter = np.where(apart, 0.0, np.where(nested, PI32 * rmin**2, lens))
        ag, ap = PI32 * rgk**2, PI32 * rp**2
        iou = inter / (ag + ap - inter + 1e-6)
        cl = np.where(nested, rmax, (rgk + rp + d) / 2)
        cs = PI32 * cl**2
        giou = iou - (cs - (ag + ap - inter)) / cs
        out[g] = (1 - giou).sum(1) / 24 / 2
        scale = np.maximum(d, 1.0)
        bmargin = min(bmargin, float(np.min(np.abs(np.abs(rgk - rp) - d) / scale)),
                      float(np.min(np.abs(d - (rgk + rp)) / scale)))
    return out, bmargin


def cls_cost(gt_classes, cls_logits, obj_logits):
    """Class cost, [G, P] (losses.py:399-416)."""
    cls_logits = np.asarray(cls_logits, np.float64)
    obj_logits = np.asarray(obj_logits, np.float64).reshape(-1, 1)
    p = np.sqrt(1 / (1 + np.exp(-cls_logits)) * (1 / (1 + np.exp(-obj_logits))))
    lp = np.maximum(np.log(p), -100.0)
    l1p = np.maximum(np.log1p(-p), -100.0)
    total = -l1p.sum(1)
    c = np.asarray(gt_classes).astype(np.int64)
    return total[None, :] + l1p[:, c].T - lp[:, c].T


def assign_image(gt50, gt_classes, out_img, xc, yc, st, num_classes=80):
    """One image's SimOTA in float64.  Returns decisions and the decision margins.

    margins: dict with
      in_box   min |angle_sum - 350| (degrees)
      center   min |window delta| (pixels)
      branch   min relative distance of a ray from a containment / disjointness switch
      dyn_k    min distance of a top-10 sum from an integer
      topk     min relative gap between the k-th and (k+1)-th cost of a GT
      argmin   min relative gap between best and second-best cost at a contested anchor
    """
    gt50 = np.asarray(gt50, np.float64)
    out_img = np.asarray(out_img, np.float64)
    xc = np.asarray(xc, np.float64)
    yc = np.asarray(yc, np.float64)
    st = np.asarray(st, np.float64)
    G = gt50.shape[0]
    ang = angle_sums(gt50, xc, yc)
    in_box = ang >= 350.0
    cd = center_deltas(gt50, xc, yc, st)
    in_ctr = cd > 0.0
    cand = in_box.any(0) | in_ctr.any(0)
    idx = np.nonzero(cand)[0]
    both = (in_box & in_ctr)[:, idx]
    pv, bmargin = pair_values(gt50, out_img[idx, :26])
    cc = cls_cost(gt_classes, out_img[idx, 27:27 + num_classes], out_img[idx, 26])
    base = cc + 3.0 * (-np.log(pv + 1e-8))
    # the +1e5 penalty is added in fp32 (spacing 1/128 at 1e5): emulate that quantisation so exact
    # fp32 ties between penalised entries show up as zero margins here too
    pen = (base.astype(np.float32) + np.float32(100000.0)).astype(np.float64)
    cost = np.where(both, base, pen)
    kc = min(10, idx.shape[0])
    top = -np.sort(-pv, axis=1)[:, :kc]
    sums = top.sum(1)
    dyn_k = np.maximum(np.trunc(sums).astype(np.int64), 1)
    m_dyn = float(np.min(np.minimum(sums - np.floor(sums), np.ceil(sums) - sums))) if G else np.inf
    match = np.zeros((G, idx.shape[0]), dtype=np.int64)
    m_topk = np.inf
    for g in range(G):
        order = np.argsort(cost[g], kind="stable")
        k = int(dyn_k[g])
        match[g, order[:k]] = 1
        if k < order.shape[0]:
            a, b = cost[g, order[k - 1]], cost[g, order[k]]
            m_topk = min(m_topk, float((b - a) / max(abs(a), 1e-12)))
    claims = match.sum(0)
    multi = claims > 1
    m_arg = np.inf
    if multi.any():
        sub = cost[:, multi]
        best = sub.argmin(0)
        if G > 1:
            srt = np.sort(sub, axis=0)
            m_arg = float(np.min((srt[1] - srt[0]) / np.maximum(np.abs(srt[0]), 1e-12)))
        match[:, multi] = 0
        match[best, np.nonzero(multi)[0]] = 1
    fg_in = match.sum(0) > 0
    fg = np.zeros(cand.shape[0], dtype=bool)
    fg[idx[fg_in]] = True
    matched = match[:, fg_in].argmax(0)
    margins = dict(in_box=float(np.min(np.abs(ang - 350.0))), center=float(np.min(np.abs(cd))),
                   branch=bmargin, dyn_k=m_dyn, topk=m_topk, argmin=m_arg)
    return dict(cand=cand, fg=fg, matched=matched, dyn_k=dyn_k, ious=(match * pv).sum(0)[fg_in],
                margins=margins, n_multi=int(multi.sum()))


# fp32 value errors observed on this path: angle sum ~2e-4 deg, window delta ~3e-5 px, top-10 sum
# ~1e-6, costs ~1e-6 relative.  Thresholds sit >= 10x above them.  ``branch`` is reported but not
# required: the containment / disjointness switches are evaluated with correctly rounded IEEE
# sub / mul / add / sqrt in the same order by every fp32 implementation, so they cannot flip.
DEFAULT_THRESHOLDS = dict(in_box=2e-3, center=1e-3, branch=0.0, dyn_k=1e-4, topk=1e-5, argmin=1e-5)


def certified(margins, thresholds=None):
    th = dict(DEFAULT_THRESHOLDS)
    if thresholds:
        th.update(thresholds)
    return all(margins[k] > th[k] for k in th if th[k] > 0)


def anchor_centres(x_shifts, y_shifts, strides):
    """xc, yc, stride per anchor (losses.py:505-516) from the head's 3-lists."""
    xs = np.concatenate([np.asarray(t.cpu()).reshape(-1) for t in x_shifts]).astype(np.float64)
    ys = np.concatenate([np.asarray(t.cpu()).reshape(-1) for t in y_shifts]).astype(np.float64)
    st = np.concatenate([np.asarray(t.cpu()).reshape(-1) for t in strides]).astype(np.float64)
    return xs * st + 0.5 * st, ys * st + 0.5 * st, st


def certify_batch(outputs, labels, x_shifts, y_shifts, strides, num_classes=80, thresholds=None):
    """Margins of a whole batch; returns (all_certified, per-image list)."""
    xc, yc, st = anchor_centres(x_shifts, y_shifts, strides)
    outputs = np.asarray(outputs.cpu(), np.float32)
    labels = np.asarray(labels.cpu(), np.float32)
    res, ok = [], True
    for b in range(outputs.shape[0]):
        n = int((labels[b].sum(1) > 0).sum())
        if n == 0:
            res.append(None)
            continue
        r = assign_image(labels[b, :n, 1:], labels[b, :n, 0], outputs[b], xc, yc, st, num_classes)
        ok = ok and certified(r["margins"], thresholds)
        res.append(r)
    return ok, res
