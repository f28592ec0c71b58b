"""Host-side mirror of the tail of ``YOLOXHead.forward(train=True)`` (``models/yolo_head_24p.py:143-199, 212-237``): what
stands between the head's prediction convs and ``Loss_Function.forward``.

Two ways to hand the conv outputs to the loss:

* ``train_outputs(reg, obj, cls, strides)``            the FUSED way: builds only the three grid lists (cached per
  shape, no per-step kernels) and wraps the conv outputs in ``RawLevels``; the loss kernels decode on load.
* ``train_outputs(reg, obj, cls, strides, fused=False)`` the reference's way with torch ops: cat -> view -> permute ->
  reshape -> centre / radius decode -> cat over the levels (several passes over the [B, A, 27 + nc] buffer); kept for
  comparison (bench.py ``train_raw`` block) and for callers that need the decoded buffer itself.

Both return the 5-tuple ``(x_shifts, y_shifts, expanded_strides, outputs, origin_preds)`` of the reference.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from .engine import RawLevels

_grid_cache: Dict[tuple, Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]] = {}


def level_grids(shapes: Sequence[Tuple[int, int]], strides: Sequence[float], device):
    """x_shifts / y_shifts / expanded_strides for levels of ``(H, W)`` cells (yolo_head_24p.py:171-174, 222-230); one
    entry per (shapes, strides, device), shared by every step."""
    key = (tuple(map(tuple, shapes)), tuple(float(s) for s in strides), str(device))
    hit = _grid_cache.get(key)
    if hit is None:
        xs, ys, ss = [], [], []
        for (h, w), s in zip(shapes, strides):
            yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
            grid = torch.stack((xv, yv), 2).view(1, -1, 2).to(device=device, dtype=torch.float32)
            xs.append(grid[:, :, 0].contiguous())
            ys.append(grid[:, :, 1].contiguous())
            ss.append(torch.full((1, h * w), float(s), device=device))
        hit = (xs, ys, ss)
        _grid_cache[key] = hit
    return hit


def train_outputs(reg: Sequence[torch.Tensor], obj: Sequence[torch.Tensor], cls: Sequence[torch.Tensor],
                  strides: Sequence[float] = (8, 16, 32), fused: bool = True):
    shapes = [tuple(r.shape[-2:]) for r in reg]
    xs, ys, ss = level_grids(shapes, strides, reg[0].device)
    if fused:
        return xs, ys, ss, RawLevels(reg, obj, cls), []
    outs = []
    for r, o, c, s, gx, gy in zip(reg, obj, cls, strides, xs, ys):
        out = torch.cat([r, o, c], 1)                                          # yolo_head_24p.py:165
        b, ch, h, w = out.shape
        out = out.view(b, 1, ch, h, w).permute(0, 1, 3, 4, 2).reshape(b, h * w, -1)   # :226-229
        grid = torch.stack((gx, gy), 2)
        out[..., :2] = (out[..., :2] + grid) * s                               # :233
        out[..., 2:26] = torch.exp(out[..., 2:26]) * s                         # :235
        outs.append(out)
    return xs, ys, ss, torch.cat(outs, 1), []


def infer_outputs(reg: Sequence[torch.Tensor], obj: Sequence[torch.Tensor], cls: Sequence[torch.Tensor],
                  strides: Sequence[float] = (8, 16, 32), fused: bool = True):
    """The tail of ``YOLOXHead.forward(train=False)`` (``yolo_head_24p.py:191, 201-211``) + ``decode_outputs``
    (``:239-256``).  ``fused=True``: the raw conv outputs wrapped for ``p24.boxes.postprocess`` (sigmoid and decode happen
    inside its filter pass, the decoded ``[B, A, 27 + nc]`` prediction is never written); ``fused=False``: the reference's
    torch ops, returning that prediction."""
    if fused:
        return RawLevels(reg, obj, cls, strides=strides)
    outs = [torch.cat([r, o.sigmoid(), c.sigmoid()], 1) for r, o, c in zip(reg, obj, cls)]
    shapes = [tuple(x.shape[-2:]) for x in outs]
    outputs = torch.cat([x.flatten(start_dim=2) for x in outs], dim=2).permute(0, 2, 1)
    xs, ys, ss = level_grids(shapes, strides, outputs.device)
    grids = torch.stack((torch.cat(xs, 1), torch.cat(ys, 1)), 2)
    st = torch.cat(ss, 1).unsqueeze(-1)
    outputs[..., :2] = (outputs[..., :2] + grids) * st
    outputs[..., 2:26] = torch.exp(outputs[..., 2:26]) * st
    return outputs
