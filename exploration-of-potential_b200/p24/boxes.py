"""Drop-in replacements for the hot-path functions of the reference's ``utils/boxes.py`` (YOLOX-24p), backed by
libp24_b200:

  postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False)   utils/boxes.py:29
  bboxes_iou(bboxes_a[G,50], bboxes_b[P,26]) -> [G, P]                                       utils/boxes.py:166
  circle_inter(c_gtx, c_gty, gt_r, c_pdx, c_pdy, pd_r) -> ([G*P,24], [G*P,24])               utils/boxes.py:102

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .engine import _check_cuda_f32, _stream_ptr

_WS = {}


_COEF = {}


def spiral_coefficients(device="cpu"):
    """theta_k * cos(theta_k), theta_k * sin(theta_k) exactly as the reference evaluates them (boxes.py:30-33): fp32
    torch ops ON THE DEVICE OF THE PREDICTION (sic: not cos / sin).  Evaluated once per device and kept on the host
    (the C ABI takes them as host constants); a CUDA device costs one read-back the first time."""
    key = str(torch.device(device))
    if key not in _COEF:
        theta = torch.tensor(15 * np.pi / 180, device=device)
        th = torch.arange(24, device=device) * theta
        _COEF[key] = ((th * torch.cos(th)).cpu().contiguous(), (th * torch.sin(th)).cpu().contiguous())
    return _COEF[key]


def polar_coefficients(device="cpu"):
    """cos(theta_k), sin(theta_k), theta_k = k * 15 degrees: the decode the reference DRAWS with (``show_24p.py:326,
    346-348``) and the 24-point representation means; fp32 torch ops on the given device, cached like the spiral ones."""
    key = "polar:" + str(torch.device(device))
    if key not in _COEF:
        theta = torch.tensor(15 * np.pi / 180, device=device)
        th = torch.arange(24, device=device) * theta
        _COEF[key] = (torch.cos(th).cpu().contiguous(), torch.sin(th).cpu().contiguous())
    return _COEF[key]


def decode_polygons(rows, device=None):
    """Vertices ``[n, 24, 2]`` of detection rows ``[n, >= 26]`` (``[cx, cy, r0..r23, ...]``): ``cx + r_k cos(theta_k)``,
    ``cy + r_k sin(theta_k)`` -- the drawing-side decode of ``show_24p.py:344-348`` without its integer truncation."""
    dev = rows.device if device is None else device
    cx, cy = polar_coefficients(dev)
    cx, cy = cx.to(rows.device), cy.to(rows.device)
    return torch.stack((rows[:, 2:26] * cx + rows[:, 0:1], rows[:, 2:26] * cy + rows[:, 1:2]), dim=2)


def _workspace(key, nbytes, device):
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes + 256 or buf.device != device:
        buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf, (buf.data_ptr() + 255) & ~255


def postprocess_raw(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False, want_rects=False,
                    coef=None):
    """Device-resident result of the whole batch: (cand_count[B], det_count[B], det_rows[B, A, 29], keep_idx[B, A],
    rects[B, A, 4] | None).  No host synchronisation.  ``coef``: (coef_x[24], coef_y[24]) host fp32 tensors overriding the
    coefficients evaluated on the prediction's device (tests against CPU-made fixtures pass the CPU ones)."""
    lib = _lib.load()
    from .engine import RawLevels
    raw = isinstance(prediction, RawLevels)
    if raw:
        # the head's raw conv outputs (p24.head.infer_outputs): sigmoid + decode happen inside the filter pass
        if prediction.num_classes != num_classes:
            raise IndexError("raw cls tensors must have num_classes channels")
        planes, plane_bs = prediction.planes()
        lv, nlev = prediction.level_table()
        B, A, dev = prediction.batch, prediction.num_anchors, prediction.device
    else:
        _check_cuda_f32(prediction, "prediction")
        if prediction.dim() != 3 or prediction.shape[2] != 27 + num_classes:
            raise IndexError("prediction must be [B, A, 27 + num_classes]")
        if prediction.stride(2) != 1:
            prediction = prediction.contiguous()
        B, A, _ = prediction.shape
        dev = prediction.device
    cand = torch.empty(B, dtype=torch.int32, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    rows = torch.empty((B, A, 29), dtype=torch.float32, device=dev)
    keep = torch.empty((B, A), dtype=torch.int32, device=dev)
    rects = torch.empty((B, A, 4), dtype=torch.float32, device=dev) if want_rects else None
    nbytes = lib.p24_postprocess_workspace_bytes(B, A)
    _, ws_ptr = _workspace(("post", B, A, str(dev)), nbytes, dev)
    cx, cy = spiral_coefficients(dev) if coef is None else coef
    cxp = cx.numpy().ctypes.data_as(C.POINTER(C.c_float))
    cyp = cy.numpy().ctypes.data_as(C.POINTER(C.c_float))
    # the reference compares fp32 tensors with Python floats: the scalars are rounded to fp32 (boxes.py:55)
    tail = (B, A, num_classes, cxp, cyp, float(np.float32(conf_thre)), float(np.float32(nms_thre)),
            1 if class_agnostic else 0, cand.data_ptr(), cnt.data_ptr(), rows.data_ptr(),
            keep.data_ptr(), rects.data_ptr() if rects is not None else None, ws_ptr, nbytes, _stream_ptr(dev))
    with torch.cuda.device(dev):
        if raw:
            n = len(planes)
            code = lib.p24_postprocess_raw((C.c_void_p * n)(*[t.data_ptr() for t in planes]), (C.c_int64 * n)(*plane_bs),
                                           lv, nlev, *tail)
        else:
            code = lib.p24_postprocess(prediction.data_ptr(), prediction.stride(0), prediction.stride(1), *tail)
    _lib.check(code, "p24_postprocess_raw" if raw else "p24_postprocess")
    return cand, cnt, rows, keep, rects


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False, decode="spiral"):
    """``utils.boxes.postprocess``: list of B entries, ``Tensor[n_i, 29]`` (rows ``[cx, cy, r0..r23, obj, class_conf,
    class_pred]`` in NMS order) or ``None``.  For B >= 2 the reference itself raises (boxes.py:64-65); this returns
    the per-image result for each image.

    ``decode="spiral"`` (default) is the reference's rectangle: points ``r_k * theta_k cos(theta_k)`` (boxes.py:32-33, a
    known defect kept for parity).  ``decode="polar"`` is the OPT-IN corrected variant (it changes the results): the
    rectangle of the true polygon ``r_k cos(theta_k), r_k sin(theta_k)`` the reference draws (show_24p.py:346-348)."""
    from .engine import RawLevels
    if not isinstance(prediction, RawLevels):
        if prediction.shape[0] == 0:
            return []
        if prediction.shape[1] == 0:
            return [None for _ in range(len(prediction))]
    if decode not in ("spiral", "polar"):
        raise ValueError("decode must be 'spiral' (the reference) or 'polar' (true polygon rectangle)")
    coef = None
    if decode == "polar":
        coef = polar_coefficients(prediction.device)
    _, cnt, rows, _, _ = postprocess_raw(prediction, num_classes, conf_thre, nms_thre, class_agnostic, coef=coef)
    counts = cnt.tolist()  # the one host read (the reference returns a Python list)
    return [rows[i, :n].clone() if n else None for i, n in enumerate(counts)]


def bboxes_iou(bboxes_a, bboxes_b):
    """``utils.boxes.bboxes_iou``: gt ``[G, 50]`` x pred ``[P, 26]`` -> ``[G, P]`` mean ray loss / 2."""
    if bboxes_b.shape[1] != 26 or bboxes_a.shape[1] != 50:
        raise IndexError
    lib = _lib.load()
    a = bboxes_a.reshape(-1, 50).float()
    b = bboxes_b.reshape(-1, 26)
    _check_cuda_f32(a, "bboxes_a")
    _check_cuda_f32(b, "bboxes_b")
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    G, P = a.shape[0], b.shape[0]
    out = torch.empty((G, P), dtype=torch.float32, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(lib.p24_pair_iou(a.data_ptr(), a.stride(0), G, b.data_ptr(), b.stride(0), P, out.data_ptr(),
                                    _stream_ptr(b.device)), "p24_pair_iou")
    return out


def circle_inter(c_gtx, c_gty, gt_r, c_pdx, c_pdy, pd_r):
    """``utils.boxes.circle_inter`` (pairwise): ``([G*P, 24], [G*P, 24])`` intersection areas and centre distances."""
    from .losses import IOUloss
    G, P = c_gtx.shape[0], c_pdx.shape[0]
    gx = c_gtx.reshape(G, 1).repeat_interleave(P, 0).reshape(-1)
    gy = c_gty.reshape(G, 1).repeat_interleave(P, 0).reshape(-1)
    px = c_pdx.reshape(P, 1).repeat(G, 1).reshape(-1)
    py = c_pdy.reshape(P, 1).repeat(G, 1).reshape(-1)
    return IOUloss().circle_inter(gx, gy, gt_r.repeat_interleave(P, 0), px, py, pd_r.repeat(G, 1))
