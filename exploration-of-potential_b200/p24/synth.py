"""Seeded synthetic inputs for the YOLOX-24p loss / SimOTA / postprocess path.

Shapes and distributions follow SURVEY.md §8(d).  Everything is generated with a CPU
``torch.Generator`` (so the same seed gives the same tensors on every host) and then moved
to the requested device.

The head-output layout is the one ``YOLOXHead.forward(train=True)`` produces
(reference ``yolox_24p/models/yolo_head_24p.py:167-197, 212-237``):
``outputs[B, A, 27 + nc]`` fp32, channels 0-1 decoded centre ``(raw + grid) * stride``,
2-25 decoded radii ``exp(raw) * stride``, 26 objectness logit, 27.. class logits; anchors
level-major (stride 8, 16, 32), row-major y then x.
Labels follow ``TrainTransform`` (``yolox_24p/datasets/data_augment.py:156-173``):
``labels[B, Lmax, 51] = [cls, cx, cy, x0, y0, ... x23, y23]`` absolute pixels, zero padded.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

STRIDES = (8, 16, 32)
N_RAYS = 24


def level_sizes(img_size: int, strides: Sequence[int] = STRIDES) -> List[Tuple[int, int]]:
    return [(img_size // s, img_size // s) for s in strides]


def make_grids(img_size: int = 640, strides: Sequence[int] = STRIDES, device="cpu"):
    """x_shifts, y_shifts, expanded_strides as 3-lists of ``[1, hw]`` fp32 tensors
    (what ``YOLOXHead.forward`` returns, ``yolo_head_24p.py:173-176, 222-230``)."""
    xs, ys, ss = [], [], []
    for (h, w), s in zip(level_sizes(img_size, strides), strides):
        yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, -1, 2).float()
        xs.append(grid[:, :, 0].contiguous().to(device))
        ys.append(grid[:, :, 1].contiguous().to(device))
        ss.append(torch.full((1, h * w), float(s)).to(device))
    return xs, ys, ss


def num_anchors(img_size: int = 640, strides: Sequence[int] = STRIDES) -> int:
    return sum(h * w for h, w in level_sizes(img_size, strides))


def make_head_outputs(batch: int, img_size: int = 640, num_classes: int = 80, seed: int = 0,
                      device="cpu", strides: Sequence[int] = STRIDES, prior_logit: float = -4.6,
                      raw_std: float = 0.5) -> torch.Tensor:
    """Random-init-like decoded head output ``[B, A, 27 + nc]`` (training layout, logits).

    ``device="cpu"`` (default) draws from a CPU generator: the same tensors on every host.  With a CUDA device the
    values are drawn and decoded there (large batches: bench.py configs[4]); same distributions, different numbers."""
    dev = torch.device(device)
    gen_dev = dev if dev.type == "cuda" else torch.device("cpu")
    g = torch.Generator(device=gen_dev).manual_seed(seed)
    A = num_anchors(img_size, strides)
    C = 27 + num_classes
    out = torch.randn(batch, A, C, generator=g, device=gen_dev) * raw_std
    out[:, :, 26:] += prior_logit
    xs, ys, ss = make_grids(img_size, strides)
    gx = torch.cat(xs, 1)[0].to(gen_dev)
    gy = torch.cat(ys, 1)[0].to(gen_dev)
    st = torch.cat(ss, 1)[0].to(gen_dev)
    raw01 = out[:, :, 0:2].clone()
    out[:, :, 0] = (raw01[:, :, 0] + gx) * st
    out[:, :, 1] = (raw01[:, :, 1] + gy) * st
    out[:, :, 2:26] = torch.exp(out[:, :, 2:26]) * st[None, :, None]
    return out.to(device)


def make_raw_levels(batch: int, img_size: int = 640, num_classes: int = 80, seed: int = 0, device="cpu",
                    strides: Sequence[int] = STRIDES, prior_logit: float = -4.6, raw_std: float = 0.5):
    """Random-init-like RAW conv outputs of the head (yolo_head_24p.py:160-164), per level:
    ``reg [B, 26, H, W]``, ``obj [B, 1, H, W]``, ``cls [B, nc, H, W]`` (same distributions as ``make_head_outputs``
    before its decode).  Returns (reg list, obj list, cls list)."""
    dev = torch.device(device)
    gen_dev = dev if dev.type == "cuda" else torch.device("cpu")
    g = torch.Generator(device=gen_dev).manual_seed(seed)
    reg, obj, cls = [], [], []
    for h, w in level_sizes(img_size, strides):
        reg.append((torch.randn(batch, 26, h, w, generator=g, device=gen_dev) * raw_std).to(device))
        obj.append((torch.randn(batch, 1, h, w, generator=g, device=gen_dev) * raw_std + prior_logit).to(device))
        cls.append((torch.randn(batch, num_classes, h, w, generator=g, device=gen_dev) * raw_std + prior_logit).to(device))
    return reg, obj, cls


def make_labels(batch: int, num_gt, max_labels: int = 50, img_size: int = 640, num_classes: int = 80,
                seed: int = 0, kind: str = "smooth", device="cpu", radius_range=(0.025, 0.2)) -> torch.Tensor:
    """``labels[B, Lmax, 51]``.  ``num_gt`` is an int or a per-image sequence.

    kind = "smooth": r_k = r0 (1 + sum_{m=1..3} a_m cos(m 15deg k + phi_m)), a_m ~ U(0, 0.2)
    kind = "spiky" : r_k ~ U(lo, hi) i.i.d. per ray
    """
    g = torch.Generator().manual_seed(seed)
    S = float(img_size)
    lo, hi = radius_range[0] * S, radius_range[1] * S
    counts = [num_gt] * batch if isinstance(num_gt, int) else list(num_gt)
    assert len(counts) == batch and max(counts + [0]) <= max_labels
    labels = torch.zeros(batch, max_labels, 51, dtype=torch.float64)
    k = torch.arange(N_RAYS, dtype=torch.float64)
    ang = k * (15.0 * math.pi / 180.0)
    for b, n in enumerate(counts):
        if n == 0:
            continue
        cx = (0.1 + 0.8 * torch.rand(n, generator=g, dtype=torch.float64)) * S
        cy = (0.1 + 0.8 * torch.rand(n, generator=g, dtype=torch.float64)) * S
        cls = torch.randint(0, num_classes, (n,), generator=g).double()
        if kind == "spiky":
            r = lo + (hi - lo) * torch.rand(n, N_RAYS, generator=g, dtype=torch.float64)
        elif kind == "smooth":
            r0 = lo + (hi - lo) * torch.rand(n, 1, generator=g, dtype=torch.float64)
            a = 0.2 * torch.rand(n, 3, generator=g, dtype=torch.float64)
            phi = 2 * math.pi * torch.rand(n, 3, generator=g, dtype=torch.float64)
            m = torch.arange(1, 4, dtype=torch.float64)
            wob = (a[:, :, None] * torch.cos(m[None, :, None] * ang[None, None, :] + phi[:, :, None])).sum(1)
            r = r0 * (1.0 + wob)
        else:
            raise ValueError(kind)
        px = cx[:, None] + r * torch.cos(ang)[None, :]
        py = cy[:, None] + r * torch.sin(ang)[None, :]
        labels[b, :n, 0] = cls
        labels[b, :n, 1] = cx
        labels[b, :n, 2] = cy
        labels[b, :n, 3::2] = px
        labels[b, :n, 4::2] = py
    return labels.float().to(device)


def make_postprocess_input(batch: int, img_size: int = 640, num_classes: int = 80, seed: int = 3,
                           device="cpu", strides: Sequence[int] = STRIDES, hot_frac: float = 0.03) -> torch.Tensor:
    """Decoded inference-layout prediction ``[B, A, 27 + nc]`` with obj/cls already in (0,1)
    (``yolo_head_24p.py:191``).  SURVEY.md §8(d) config 4."""
    g = torch.Generator().manual_seed(seed)
    A = num_anchors(img_size, strides)
    C = 27 + num_classes
    p = torch.empty(batch, A, C)
    p[:, :, 0:2] = torch.rand(batch, A, 2, generator=g) * float(img_size)
    p[:, :, 2:26] = 4.0 + 60.0 * torch.rand(batch, A, 24, generator=g)
    p[:, :, 26] = torch.rand(batch, A, generator=g)
    p[:, :, 27:] = 0.2 * torch.rand(batch, A, num_classes, generator=g)
    hot = torch.rand(batch, A, generator=g) < hot_frac
    hot_cls = torch.randint(0, num_classes, (batch, A), generator=g)
    hot_val = 0.5 + 0.5 * torch.rand(batch, A, generator=g)
    bi, ai = hot.nonzero(as_tuple=True)
    p[bi, ai, 27 + hot_cls[bi, ai]] = hot_val[bi, ai]
    return p.to(device)
