"""Host-side mirror of the LABEL half of the dataset transform (``datasets/data_augment.py:131-174``,
``datasets/coco24p.py:78-131``): the ``[max_labels, 51]`` wire format the loss consumes, packed for a whole batch by
one CUDA kernel instead of per image in numpy inside the DataLoader workers.

The image half (cv2 resize + pad, ``data_augment.py:96-128``) is dataset I/O and stays where it is.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib
from .engine import _stream_ptr


class TrainTransform:
    """``TrainTransform(max_labels=50)`` of the reference, label side, batched on the GPU.

    ``pack(targets, image_hw, input_dim, device)``
        targets   sequence (one per image) of float64 arrays ``[n, 51]`` = the label-file rows
                  ``[cls, cx, cy, 24 x (x, y)]`` normalised to [0, 1]; an image without labels is an array with
                  ``shape[1] == 0`` (what ``pull_item`` yields for an empty file) or ``[0, 51]``
        image_hw  sequence of (height, width) of the resized, unpadded images the reference transform receives
        input_dim (in_h, in_w) of the padded network input
    returns ``labels [B, max_labels, 51]`` fp32 on ``device`` (and ``nlabel [B]`` int32 with ``return_counts=True``),
    bit-identical to stacking the reference's per-image ``padded_labels``.
    """

    def __init__(self, max_labels: int = 50, flip_prob: float = 0.5):
        self.max_labels = max_labels
        self.flip_prob = flip_prob  # (kept for signature parity: the 24p transform never flips, data_augment.py:131-174)

    def pack(self, targets: Sequence[np.ndarray], image_hw: Sequence[Tuple[int, int]], input_dim: Tuple[int, int],
             device="cuda", return_counts: bool = False):
        lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.P24Error("label packing runs on a CUDA device (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        B = len(targets)
        if B == 0 or len(image_hw) != B:
            raise IndexError("one (height, width) per image is required")
        rows, offsets = [], [0]
        for t in targets:
            t = np.asarray(t, dtype=np.float64)
            if t.ndim == 1:
                t = t[np.newaxis, :]          # coco24p.py:84-85
            if t.shape[1] == 0:               # data_augment.py:141-144: an empty label file
                t = np.zeros((0, 51))
            if t.shape[1] != 51:
                raise IndexError("a target row is [cls, cx, cy, 24 x (x, y)]: 51 values")
            rows.append(t)
            offsets.append(offsets[-1] + t.shape[0])
        flat = np.ascontiguousarray(np.concatenate(rows, 0)) if offsets[-1] else np.zeros((0, 51))
        # one small pinned staging buffer per call: targets | offsets | shapes
        d_t = torch.from_numpy(flat).to(dev, non_blocking=True) if offsets[-1] else None
        d_o = torch.tensor(offsets, dtype=torch.int32).to(dev, non_blocking=True)
        d_s = torch.tensor([[int(h), int(w)] for h, w in image_hw], dtype=torch.int32).to(dev, non_blocking=True)
        labels = torch.empty((B, self.max_labels, 51), dtype=torch.float32, device=dev)
        nlabel = torch.empty(B, dtype=torch.int32, device=dev) if return_counts else None
        with torch.cuda.device(dev):
            _lib.check(lib.p24_pack_labels(d_t.data_ptr() if d_t is not None else None, d_o.data_ptr(), d_s.data_ptr(),
                                           int(input_dim[0]), int(input_dim[1]), B, self.max_labels, labels.data_ptr(),
                                           nlabel.data_ptr() if nlabel is not None else None, _stream_ptr(dev)),
                       "p24_pack_labels")
        return (labels, nlabel) if return_counts else labels
