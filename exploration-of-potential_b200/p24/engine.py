"""Host side of the fused SimOTA + loss-sum path: buffer management and the ctypes calls.

PyTorch is used for device memory and streams only; all arithmetic runs in libp24_b200
(``csrc/p24_simota.cu``).  Nothing here synchronises with the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence, Tuple

import ctypes as C

import torch

from . import lib as _lib

F_NO_PRUNE = 1   # evaluate every polygon angle sum exactly (self-check of the geometric pruning)
F_NO_FILTER = 2  # dynamic k from the exact value of every (GT, candidate) pair (self-check of the far-pair filter)
F_ALL_ROWS = 4   # every label row is a GT (per-image API)
F_NO_PDL = 8     # plain stream-ordered launches
F_EARLY_PREP = 16  # back-to-back steps on resident inputs: k_prep runs beside the previous step's last kernel


class RawLevels:
    """The head's raw per-level conv outputs (``yolo_head_24p.py:160-164``), handed to the loss instead of the decoded
    ``[B, A, 27 + nc]`` buffer: ``reg[k] [B, 26, H, W]``, ``obj[k] [B, 1, H, W]``, ``cls[k] [B, nc, H, W]`` for every
    level k.  The kernels decode on load (``get_output_and_grid``, ``yolo_head_24p.py:233-235``); the cat / view /
    permute / reshape copies of ``YOLOXHead.forward(train=True)`` never happen.  Channel slices of one concatenated
    ``[B, 27 + nc, H, W]`` tensor are fine (only the H x W planes must be dense)."""

    def __init__(self, reg, obj, cls, strides=None):
        self.reg, self.obj, self.cls = list(reg), list(obj), list(cls)
        # (the loss takes the strides from the grid lists of the 5-tuple; the postprocess needs them here)
        self.strides = None if strides is None else [float(s) for s in strides]
        if not (len(self.reg) == len(self.obj) == len(self.cls)) or not self.reg:
            raise IndexError("RawLevels needs reg / obj / cls tensors for the same levels")
        for k, (r, o, c) in enumerate(zip(self.reg, self.obj, self.cls)):
            if r.dim() != 4 or r.shape[1] != 26 or o.shape[1] != 1 or r.shape[2:] != o.shape[2:] or r.shape[2:] != c.shape[2:] \
                    or not (r.shape[0] == o.shape[0] == c.shape[0]):
                raise IndexError(f"level {k}: expected reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W]")

    @property
    def device(self):
        return self.reg[0].device

    @property
    def batch(self):
        return self.reg[0].shape[0]

    @property
    def num_classes(self):
        return self.cls[0].shape[1]

    @property
    def num_anchors(self):
        return sum(r.shape[2] * r.shape[3] for r in self.reg)

    def requires_grad(self):
        return any(t.requires_grad for lst in (self.reg, self.obj, self.cls) for t in lst)

    def level_table(self):
        """HOST (anchor offset, W, H, stride bits) per level, as the C ABI takes it."""
        import struct
        if self.strides is None or len(self.strides) != len(self.reg):
            raise IndexError("RawLevels needs one stride per level here")
        lv, off = [], 0
        for r, s in zip(self.reg, self.strides):
            H, W = r.shape[2], r.shape[3]
            lv += [off, W, H, struct.unpack("<i", struct.pack("<f", s))[0]]
            off += H * W
        return (C.c_int32 * len(lv))(*lv), len(self.reg)

    def planes(self):
        """(tensor list in ABI order, batch strides): every tensor with dense H x W planes and channel stride H * W."""
        out, bs = [], []
        for lst in (self.reg, self.obj, self.cls):
            for t in lst:
                _check_cuda_f32(t, "raw level tensor")
                H, W = t.shape[2], t.shape[3]
                if t.stride(3) != 1 or t.stride(2) != W or (t.shape[1] > 1 and t.stride(1) != H * W):
                    t = t.contiguous()
                out.append(t)
                bs.append(t.stride(0))
        return out, bs


@dataclass
class Assignment:
    """Device-resident result of one batch (no host sync has happened)."""
    fg_mask: torch.Tensor      # [B, A] uint8
    matched_gt: torch.Tensor   # [B, A] int32, -1 = background
    pred_iou: torch.Tensor     # [B, A] fp32
    num_fg: torch.Tensor       # [B] int32
    num_gt: torch.Tensor       # [B] int32
    dyn_k: torch.Tensor        # [B, Lmax] int32
    sums28: torch.Tensor | None  # [28] fp32


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _check_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.P24Error(f"{name} must be a CUDA tensor: the p24 path has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")


class GridCache:
    """Concatenated x_shifts / y_shifts / expanded_strides (losses.py:193-195) and the level structure of the grid
    (anchor offset, width, height per level), cached per 3-list.  Building an entry reads the grids back once
    (one host sync per new grid, none afterwards)."""

    def __init__(self):
        self._key = None
        self._val = None

    @staticmethod
    def _levels(gx: torch.Tensor, gy: torch.Tensor, gs: torch.Tensor):
        """Level table of a grid laid out as the head builds it (yolo_head_24p.py:222-230): a new level starts
        wherever the stride value changes; inside a level the anchors are the cells of a W x H grid, row-major."""
        x, y, st = gx.cpu(), gy.cpu(), gs.cpu()
        A = x.numel()
        cuts = [0] + (torch.nonzero(st[1:] != st[:-1]).flatten() + 1).tolist() + [A]
        levels = []
        for off, end in zip(cuts[:-1], cuts[1:]):
            n = end - off
            W = int(x[off:end].max().item()) + 1
            H = int(y[off:end].max().item()) + 1
            idx = torch.arange(n)
            ok = (W * H == n and torch.equal(x[off:end], (idx % W).to(x.dtype))
                  and torch.equal(y[off:end], (idx // W).to(y.dtype)))
            if not ok:
                raise _lib.P24Error("x_shifts / y_shifts / expanded_strides do not form per-level row-major grids "
                                    "(yolo_head_24p.py:222-230): unsupported anchor layout")
            import struct
            stride_bits = struct.unpack("<i", struct.pack("<f", float(st[off])))[0]
            if not bool((st[off:end] == st[off]).all()):
                raise _lib.P24Error("expanded_strides is not constant within a level")
            levels.append((off, W, H, stride_bits))
        if len(levels) > 4:
            raise _lib.P24Error(f"{len(levels)} feature levels: at most 4 are supported")
        arr = (C.c_int32 * (4 * len(levels)))(*[v for lv in levels for v in lv])
        return arr, len(levels)

    def get(self, x_shifts, y_shifts, strides, device):
        if torch.is_tensor(x_shifts):
            x_shifts, y_shifts, strides = [x_shifts], [y_shifts], [strides]
        # shapes are part of the key: a grid re-created for another H x W with the same numel usually gets the same address
        key = tuple([(t.data_ptr(), tuple(t.shape), t._version, str(t.device)) for lst in (x_shifts, y_shifts, strides)
                     for t in lst])
        if key != self._key:
            cat = [torch.cat([t.reshape(1, -1) for t in lst], 1).reshape(-1).to(device=device, dtype=torch.float32)
                   .contiguous() for lst in (x_shifts, y_shifts, strides)]
            lv, nlev = self._levels(*cat)
            self._key, self._val = key, (cat[0], cat[1], cat[2], lv, nlev)
        return self._val


class SimOTAEngine:
    """Owns the workspace and output buffers for one (B, A, Lmax) shape on one device."""

    def __init__(self):
        self._bufs: Dict[tuple, dict] = {}
        self.grids = GridCache()
        # reuse_buffers: the Assignment of a shape is allocated once and overwritten by every call (less host work per
        # step; the caller must be done with one step's results before it enqueues the next)
        self.reuse_buffers = False
        self._out_cache: Dict[tuple, Assignment] = {}

    def _buffers(self, B, A, Lmax, device):
        key = (B, A, Lmax, str(device))
        buf = self._bufs.get(key)
        if buf is None:
            lib = _lib.load()
            nbytes = lib.p24_workspace_bytes(B, A, Lmax)
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            ptr = (ws.data_ptr() + 255) & ~255
            with torch.cuda.device(device):
                _lib.check(lib.p24_workspace_init(ptr, nbytes, _stream_ptr(device)), "p24_workspace_init")
            buf = dict(ws=ws, ptr=ptr, nbytes=nbytes)
            self._bufs[key] = buf
        return buf

    def run(self, outputs: torch.Tensor, labels: torch.Tensor, x_shifts, y_shifts, strides, num_classes: int,
            flags: int = 0, want_sums: bool = True, out: Assignment | None = None,
            finalize: tuple | None = None, comm=None) -> Assignment:
        """Enqueue the kernel chain.  ``finalize=(state26, result54, weights27)`` fuses the normalisation and
        re-weighting into the last kernel.  ``comm`` (``p24.dist.PeerComm``): the 28 sums are all-reduced over peer
        memory inside the last kernel first (one process per GPU, images sharded by rank)."""
        lib = _lib.load()
        raw = isinstance(outputs, RawLevels)
        _check_cuda_f32(labels, "labels")
        if raw:
            if outputs.num_classes != num_classes:
                raise IndexError("raw cls tensors must have num_classes channels")
            planes, plane_bs = outputs.planes()
            B, A, dev = outputs.batch, outputs.num_anchors, outputs.device
        else:
            _check_cuda_f32(outputs, "outputs")
            if outputs.dim() != 3 or outputs.shape[2] != 27 + num_classes:
                raise IndexError("outputs must be [B, A, 27 + num_classes]")
            if outputs.stride(2) != 1:
                outputs = outputs.contiguous()
            B, A, _ = outputs.shape
            dev = outputs.device
        if labels.dim() != 3 or labels.shape[2] != 51 or labels.shape[0] != B:
            raise IndexError("labels must be [B, Lmax, 51]")
        if labels.stride(2) != 1:
            labels = labels.contiguous()
        Lmax = labels.shape[1]
        gx, gy, gs, lv, nlev = self.grids.get(x_shifts, y_shifts, strides, dev)
        if gx.numel() != A:
            raise IndexError("grid length does not match the number of anchors")
        if out is None and self.reuse_buffers:
            out = self._out_cache.get((B, A, Lmax, want_sums, dev))
        if out is None:
            out = Assignment(
                fg_mask=torch.empty((B, A), dtype=torch.uint8, device=dev),
                matched_gt=torch.empty((B, A), dtype=torch.int32, device=dev),
                pred_iou=torch.empty((B, A), dtype=torch.float32, device=dev),
                num_fg=torch.empty((B,), dtype=torch.int32, device=dev),
                num_gt=torch.empty((B,), dtype=torch.int32, device=dev),
                dyn_k=torch.empty((B, max(Lmax, 1)), dtype=torch.int32, device=dev),
                sums28=torch.empty((28,), dtype=torch.float32, device=dev) if want_sums else None,
            )
        if self.reuse_buffers:
            self._out_cache[(B, A, Lmax, want_sums, dev)] = out
        if Lmax == 0:  # no label rows at all: everything is background
            out.fg_mask.zero_(); out.matched_gt.fill_(-1); out.pred_iou.zero_()
            out.num_fg.zero_(); out.num_gt.zero_(); out.dyn_k.zero_()
            labels = labels.new_zeros((B, 1, 51))
            Lmax = 1
        buf = self._buffers(B, A, Lmax, dev)
        ws_ptr = buf["ptr"]
        fin = [t.data_ptr() for t in finalize] if finalize is not None else [None, None, None]
        if raw:
            if len(outputs.reg) != nlev:
                raise IndexError("RawLevels and the grids describe different numbers of levels")
            for k, r in enumerate(outputs.reg):
                if (r.shape[3], r.shape[2]) != (lv[4 * k + 1], lv[4 * k + 2]):
                    raise IndexError(f"level {k}: the raw tensors are {r.shape[2]}x{r.shape[3]}, the grid is not")
            head = ((C.c_void_p * len(planes))(*[t.data_ptr() for t in planes]), (C.c_int64 * len(planes))(*plane_bs))
            entry, what = lib.p24_simota_loss_batch_raw, "p24_simota_loss_batch_raw"
        else:
            head = (outputs.data_ptr(), outputs.stride(0), outputs.stride(1))
            entry, what = lib.p24_simota_loss_batch, "p24_simota_loss_batch"
        args = head + (B, A, num_classes,
                labels.data_ptr(), labels.stride(0), labels.stride(1), Lmax,
                gx.data_ptr(), gy.data_ptr(), gs.data_ptr(), lv, nlev,
                out.fg_mask.data_ptr(), out.matched_gt.data_ptr(), out.pred_iou.data_ptr(),
                out.num_fg.data_ptr(), out.num_gt.data_ptr(), out.dyn_k.data_ptr(),
                out.sums28.data_ptr() if out.sums28 is not None else None,
                fin[0], fin[1], fin[2],
                ws_ptr, buf["nbytes"], flags,
                comm.pointers if comm is not None else None, comm.rank if comm is not None else 0,
                comm.nranks if comm is not None else 1, comm.next_epoch() if comm is not None else 0,
                _stream_ptr(dev))
        if torch.cuda.current_device() == dev.index:
            code = entry(*args)
        else:  # the launches go to the tensors' device
            with torch.cuda.device(dev):
                code = entry(*args)
        _lib.check(code, what)
        if comm is not None:
            # the collect half of the fused all-reduce, on the comm's side stream (sums28 / the finalize buffers are valid
            # once comm.wait() has ordered the consumer's stream behind it)
            comm.finish(out.sums28, *(finalize if finalize is not None else (None, None, None)), ws_ptr, B, A, Lmax)
        return out

    def read_status(self):
        """Sticky status words of every workspace this engine has used (one small D2H copy + stream sync each):
        list of (key, [err_bits, n_brute, n_spill, list_max, exchange_wait_cycles, list_sum, gts, n_exact]); the
        counters restart with every read, the error bits are sticky."""
        lib = _lib.load()
        out = []
        for key, buf in self._bufs.items():
            B, A, Lmax, dev = key
            dev = torch.device(dev)
            st = (C.c_int32 * 8)()
            with torch.cuda.device(dev):
                _lib.check(lib.p24_read_status(buf["ptr"], B, A, Lmax, st, _stream_ptr(dev)), "p24_read_status")
            out.append((key, list(st)))
        return out

    def finalize(self, sums28: torch.Tensor, state26: torch.Tensor, result54: torch.Tensor | None = None,
                 weights27: torch.Tensor | None = None):
        lib = _lib.load()
        dev = sums28.device
        if result54 is None:
            result54 = torch.empty(54, dtype=torch.float32, device=dev)
        if weights27 is None:
            weights27 = torch.empty(27, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            code = lib.p24_loss_finalize(sums28.data_ptr(), state26.data_ptr(), result54.data_ptr(),
                                         weights27.data_ptr(), _stream_ptr(dev))
        _lib.check(code, "p24_loss_finalize")
        return result54, weights27
