"""Drop-in replacements for the reference's ``models/losses.py`` classes (YOLOX-24p), backed by the
hand-written sm_100a kernels of libp24_b200.

Same names, argument meaning, return layout and error behaviour as the reference
(``/root/reference/yolox_24p/models/losses.py``):

  IOUloss(reduction="none")                  losses.py:14-157
  Loss_Function(num_classes)                 losses.py:159-604
      .forward(outputs_train, labels)        losses.py:175   -> 7-tuple
      .get_assignments(...)                  losses.py:360   -> (classes, fg_mask, ious, gt_inds, num_fg)
      .dynamic_k_matching(...)               losses.py:444

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib as _lib
from .engine import Assignment, RawLevels, SimOTAEngine, F_ALL_ROWS, F_EARLY_PREP, _check_cuda_f32, _stream_ptr


def _rows(t: torch.Tensor, width: int, name: str) -> torch.Tensor:
    t = t.reshape(-1, width)
    _check_cuda_f32(t, name)
    return t if t.stride(1) == 1 else t.contiguous()


class _IouLossFn(torch.autograd.Function):
    """loss24 = IOUloss.forward(pred, target)[0] with the hand-written backward (p24_iou_loss_bwd)."""

    @staticmethod
    def forward(ctx, pred, target):
        lib = _lib.load()
        p, t = _rows(pred.detach(), 26, "pred"), _rows(target.detach().float(), 50, "target")
        n = p.shape[0]
        out = torch.empty((n, 24), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(lib.p24_iou_loss_fwd(p.data_ptr(), p.stride(0), t.data_ptr(), t.stride(0), n, out.data_ptr(),
                                            _stream_ptr(p.device)), "p24_iou_loss_fwd")
        ctx.save_for_backward(p, t)
        return out

    @staticmethod
    def backward(ctx, grad24):
        lib = _lib.load()
        p, t = ctx.saved_tensors
        n = p.shape[0]
        g = grad24.contiguous().float()
        gp = torch.empty((n, 26), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(lib.p24_iou_loss_bwd(p.data_ptr(), p.stride(0), t.data_ptr(), t.stride(0), g.data_ptr(), n,
                                            gp.data_ptr(), _stream_ptr(p.device)), "p24_iou_loss_bwd")
        return gp, None


class IOUloss(nn.Module):
    """24-ray concentric-circle GIoU loss (reference ``models/losses.py:14-157``)."""

    def __init__(self, reduction="none"):
        super().__init__()
        self.reduction = reduction

    def circle_inter(self, c_gtx, c_gty, gt_r, c_pdx, c_pdy, pd_r):
        """Element-wise over N pairs -> (res_inter [N, 24], dist [N, 24])   (losses.py:23-78)."""
        lib = _lib.load()
        n = gt_r.shape[0]
        res = torch.zeros_like(gt_r, dtype=torch.float32)
        if n == 0 or pd_r.shape[0] == 0:  # losses.py:40-42
            dist = torch.sqrt((c_gtx - c_pdx) ** 2 + (c_gty - c_pdy) ** 2).unsqueeze(1).repeat(1, 24)
            return res, dist
        gr, pr = _rows(gt_r.float(), 24, "gt_r"), _rows(pd_r.float(), 24, "pd_r")
        gx, gy = c_gtx.float().contiguous(), c_gty.float().contiguous()
        px, py = c_pdx.float().contiguous(), c_pdy.float().contiguous()
        dist = torch.empty((n, 24), dtype=torch.float32, device=gr.device)
        with torch.cuda.device(gr.device):
            _lib.check(lib.p24_circle_inter_fwd(gx.data_ptr(), gy.data_ptr(), gr.data_ptr(), gr.stride(0), px.data_ptr(),
                                                py.data_ptr(), pr.data_ptr(), pr.stride(0), n, res.data_ptr(),
                                                dist.data_ptr(), _stream_ptr(gr.device)), "p24_circle_inter_fwd")
        return res, dist

    def forward(self, pred, target):
        """pred [N, 26], target [N, 50] -> (loss24 [N, 24] un-reduced 1 - giou per ray, [pd_cx, pd_cy, pd_r])."""
        if pred.shape[1] != 26 or target.shape[1] != 50:
            raise IndexError
        pred = pred.view(-1, 26)
        target = target.view(-1, 50)
        pcx, pcy, pr = pred[:, 0].to(torch.float), pred[:, 1].to(torch.float), pred[:, 2:]
        if pred.shape[0] == 0 or target.shape[0] == 0:  # losses.py:111-115
            return pred.new_zeros(1, 24), [pcx.new_zeros(1, 24), pcy.new_zeros(1, 24), pr.new_zeros(1, 24)]
        return _IouLossFn.apply(pred, target), [pcx, pcy, pr]


class _LossFn(torch.autograd.Function):
    """result54 = the whole fused loss forward; backward = p24_loss_bwd scaled by d(loss)."""

    @staticmethod
    def forward(ctx, outputs, labels, owner, x_shifts, y_shifts, strides):
        result54, weights27, asg = owner.forward_async((x_shifts, y_shifts, strides, outputs.detach(), []), labels)
        owner.wait_results()
        ctx.owner_nc = owner.num_classes
        ctx.asg = asg
        ctx.save_for_backward(outputs.detach(), labels, weights27)
        return result54

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        outputs, labels, w27 = ctx.saved_tensors
        asg = ctx.asg
        if outputs.stride(2) != 1:
            outputs = outputs.contiguous()
        lab = labels if labels.stride(2) == 1 else labels.contiguous()
        B, A, C = outputs.shape
        gout = torch.empty((B, A, C), dtype=torch.float32, device=outputs.device)
        scale = g[0:1].contiguous().float()  # only d/d(loss) is propagated (the training script backpropagates the loss)
        with torch.cuda.device(outputs.device):
            _lib.check(lib.p24_loss_bwd(outputs.data_ptr(), outputs.stride(0), outputs.stride(1), B, A, ctx.owner_nc,
                                        lab.data_ptr(), lab.stride(0), lab.stride(1), asg.fg_mask.data_ptr(),
                                        asg.matched_gt.data_ptr(), asg.pred_iou.data_ptr(), w27.data_ptr(),
                                        scale.data_ptr(), gout.data_ptr(), _stream_ptr(outputs.device)), "p24_loss_bwd")
        return gout, None, None, None, None, None


class _RawLossFn(torch.autograd.Function):
    """The same for the head's raw per-level conv outputs (``RawLevels``): backward = p24_loss_bwd_raw, one gradient
    tensor per conv output (the chain rule through the decode of yolo_head_24p.py:233-235 is applied in the kernel)."""

    @staticmethod
    def forward(ctx, labels, owner, x_shifts, y_shifts, strides, nlev, *planes):
        raw = RawLevels([t.detach() for t in planes[:nlev]], [t.detach() for t in planes[nlev:2 * nlev]],
                        [t.detach() for t in planes[2 * nlev:]])
        result54, weights27, asg = owner.forward_async((x_shifts, y_shifts, strides, raw, []), labels)
        owner.wait_results()
        ctx.owner, ctx.asg, ctx.raw, ctx.grid = owner, asg, raw, (x_shifts, y_shifts, strides)
        ctx.save_for_backward(labels, weights27)
        return result54

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        lib = _lib.load()
        labels, w27 = ctx.saved_tensors
        asg, raw = ctx.asg, ctx.raw
        dev = raw.device
        lab = labels if labels.stride(2) == 1 else labels.contiguous()
        planes, bs = raw.planes()
        grads = [torch.empty(t.shape, dtype=torch.float32, device=dev) for t in planes]
        _, _, _, lv, nlev = ctx.owner._engine.grids.get(*ctx.grid, dev)
        B, A = asg.fg_mask.shape
        scale = g[0:1].contiguous().float()
        n = len(planes)
        with torch.cuda.device(dev):
            _lib.check(lib.p24_loss_bwd_raw((C.c_void_p * n)(*[t.data_ptr() for t in planes]), (C.c_int64 * n)(*bs),
                                            (C.c_void_p * n)(*[t.data_ptr() for t in grads]), lv, nlev, B, A,
                                            ctx.owner.num_classes, lab.data_ptr(), lab.stride(0), lab.stride(1),
                                            asg.fg_mask.data_ptr(), asg.matched_gt.data_ptr(), asg.pred_iou.data_ptr(),
                                            w27.data_ptr(), scale.data_ptr(), _stream_ptr(dev)), "p24_loss_bwd_raw")
        return (None, None, None, None, None, None) + tuple(grads)


class Loss_Function(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes
        self.use_l1 = False  # never enabled by the 24p scripts (losses.py:163); the L1 branch is not provided
        self.iou_loss = IOUloss(reduction="none")
        self.bcewithlog_loss = nn.BCEWithLogitsLoss(reduction="none")
        # stateful re-weighting memory (losses.py:170-172); kept as Python/torch state on the object and,
        # like the reference, not checkpointed
        self.last_iou_loss = 1.0
        self.last_obj_loss = 1.0
        self.last_cls_loss = 1.0
        self._state26 = None
        self._engine = SimOTAEngine()
        self.last_assignment: Assignment | None = None
        # set to a torch.distributed group to shard the batch by image across GPUs: the only collective is the
        # SUM all-reduce of the 28 loss sums
        self.process_group = None
        # reuse_buffers = True: result / assignment tensors are allocated once per shape and overwritten by the next
        # forward (less host work per step); the default hands out fresh tensors like the reference does
        self.reuse_buffers = False
        self._res_cache = {}
        # pipelined = True: the caller guarantees that the inputs of every forward were complete before the previous
        # forward was enqueued (loss steps back to back on resident batches): the preparation kernel of a step then runs
        # beside the last kernel of the step before it (P24_F_EARLY_PREP).  Off by default: a head that writes `outputs`
        # right before the call is the normal case.
        self.pipelined = False

    def _result_buffers(self, device):
        if self.reuse_buffers:
            buf = self._res_cache.get(device)
            if buf is None:
                buf = (torch.empty(54, dtype=torch.float32, device=device), torch.empty(27, dtype=torch.float32, device=device))
                self._res_cache[device] = buf
            return buf
        return (torch.empty(54, dtype=torch.float32, device=device), torch.empty(27, dtype=torch.float32, device=device))

    # -- state ---------------------------------------------------------------------------------
    def _state(self, device):
        if self._state26 is None or self._state26.device != device:
            st = torch.ones(26, dtype=torch.float32, device=device)
            if torch.is_tensor(self.last_iou_loss):
                st[:24] = self.last_iou_loss.to(device)
                st[24] = self.last_obj_loss.to(device)
                st[25] = self.last_cls_loss.to(device)
            self._state26 = st
        return self._state26

    # -- the hot path ----------------------------------------------------------------------------
    def forward_async(self, outputs_train, labels, flags: int = 0):
        """Enqueue the whole loss forward; returns (result54, weights27, Assignment) device tensors.

        result54 = [loss, reg_w*loss_iou (24), loss_obj, loss_cls, num_fg/max(num_gts,1), reg_w (24), obj_w, cls_w]
        No host synchronisation happens here.
        """
        if self.use_l1:
            raise NotImplementedError("use_l1 is never enabled by the 24p scripts (losses.py:163)")
        x_shifts, y_shifts, expanded_strides, outputs = outputs_train[:4]
        # (outputs: the decoded [B, A, 27 + nc] buffer of the reference's head, or p24.engine.RawLevels: the head's raw
        # conv outputs, decoded inside the kernels)
        state = self._state(outputs.device)
        self._engine.reuse_buffers = self.reuse_buffers
        if self.pipelined:
            flags |= F_EARLY_PREP
        if self.process_group is None:
            # single GPU: the last CTA of the chain applies the normalisation and re-weighting itself
            result54, weights27 = self._result_buffers(outputs.device)
            asg = self._engine.run(outputs, labels, x_shifts, y_shifts, expanded_strides, self.num_classes,
                                   flags=flags, finalize=(state, result54, weights27))
        elif getattr(self, "peer_comm", None) is not None:
            # several GPUs of one box: the last CTA all-reduces the 28 sums over peer memory, then finalizes
            result54, weights27 = self._result_buffers(outputs.device)
            asg = self._engine.run(outputs, labels, x_shifts, y_shifts, expanded_strides, self.num_classes,
                                   flags=flags, finalize=(state, result54, weights27), comm=self.peer_comm)
        else:
            import torch.distributed as dist
            asg = self._engine.run(outputs, labels, x_shifts, y_shifts, expanded_strides, self.num_classes, flags=flags)
            dist.all_reduce(asg.sums28, op=dist.ReduceOp.SUM, group=self.process_group)
            result54, weights27 = self._engine.finalize(asg.sums28, state)
        self.last_assignment = asg
        # expose the state like the reference does (views: no copy, no sync)
        self.last_iou_loss = state[:24]
        self.last_obj_loss = state[24]
        self.last_cls_loss = state[25]
        return result54, weights27, asg

    # -- bookkeeping of the asynchronous path ---------------------------------------------------------------
    def wait_results(self):
        """Order the current stream behind everything ``forward_async`` has enqueued.  With the fused peer all-reduce
        the normalisation runs on a side stream (``p24.dist.PeerComm``): result54 / weights27 / the state are valid for
        the current stream only after this call (``forward`` makes it before it reads the result)."""
        comm = getattr(self, "peer_comm", None)
        if comm is not None:
            comm.wait()

    def check_errors(self):
        """Raise ``P24Error`` when a kernel reported an internal error (sticky bits in the workspace: window-list
        overflow, peer time-out).  Costs one small D2H copy and a stream synchronisation per workspace: ``forward``
        does it together with its own result read, ``forward_async`` leaves it to the caller."""
        for key, st in self._engine.read_status():
            if st[0]:
                raise _lib.P24Error(f"p24 kernels reported error bits {st[0]:#x} on workspace {key} "
                                    "(P24_ERR_* of include/p24.h: 1 window list overflow, 4 peer time-out)")

    def read_status(self):
        """One read of the kernels' status words (they restart with every read): the rare-path counters since the
        previous read -- GTs whose dynamic k needed exact pair values / the brute-force evaluation, GTs that spilled into
        the penalised regime, the longest and the mean top-10 candidate list -- and the microseconds between the
        first and the last rank's contribution to the last fused all-reduce arriving at this rank (rank skew plus link
        latency; clock64 cycles at the nominal 1965 MHz).  Raises ``P24Error`` on error bits."""
        st = self._engine.read_status()
        for key, s in st:
            if s[0]:
                raise _lib.P24Error(f"p24 kernels reported error bits {s[0]:#x} on workspace {key} "
                                    "(P24_ERR_* of include/p24.h: 1 window list overflow, 4 peer time-out)")
        st = [s for _, s in st]
        gts = sum(s[6] for s in st)
        return {"brute_force_gts": sum(s[1] for s in st), "exact_gts": sum(s[7] for s in st),
                "spill_gts": sum(s[2] for s in st), "gts": gts,
                "list_max": max([s[3] for s in st] + [0]), "list_mean": (sum(s[5] for s in st) / gts) if gts else 0.0,
                "list_capacity": 4096, "exchange_wait_us": max([s[4] for s in st] + [0]) / 1965.0}

    def path_stats(self):
        return self.read_status()

    def forward(self, outputs_train, labels):
        outputs = outputs_train[3]
        if isinstance(outputs, RawLevels):
            return self._forward_raw(outputs_train, labels)
        if torch.is_grad_enabled() and outputs.requires_grad:
            r = _LossFn.apply(outputs, labels, self, outputs_train[0], outputs_train[1], outputs_train[2])
            asg = self.last_assignment
        else:
            r, _, asg = self.forward_async(outputs_train, labels)
        self.wait_results()
        fg = asg.fg_mask.view(-1).bool()
        rows = outputs.reshape(-1, outputs.shape[-1])[fg]
        if rows.shape[0] == 0:  # losses.py:111-115
            draw = [rows.new_zeros(1, 24), rows.new_zeros(1, 24), rows.new_zeros(1, 24)]
        else:
            draw = [rows[:, 0], rows[:, 1], rows[:, 2:26]]
        draw += [r[28:52], r[52], r[53]]
        ratio = float(r[27].detach())  # one D2H read per step (the reference returns a Python float here)
        self.check_errors()            # ... and the kernels' error bits with it (the reference raises on its failures)
        return (r[0], r[1:25], r[25], r[26], 0.0, ratio, draw)

    def _forward_raw(self, outputs_train, labels):
        """``forward`` on the head's raw conv outputs (``p24.engine.RawLevels`` in place of the decoded buffer).  Same
        7-tuple; the drawing entries (decoded centres / radii of the foreground anchors) are gathered from the raw planes.
        Differentiable w.r.t. every conv output (``_RawLossFn``)."""
        raw = outputs_train[3]
        if torch.is_grad_enabled() and raw.requires_grad():
            r = _RawLossFn.apply(labels, self, outputs_train[0], outputs_train[1], outputs_train[2], len(raw.reg),
                                 *raw.reg, *raw.obj, *raw.cls)
            asg = self.last_assignment
        else:
            r, _, asg = self.forward_async(outputs_train, labels)
        self.wait_results()
        fg = asg.fg_mask.view(-1).bool()
        nfg = int(fg.sum())
        if nfg == 0:  # losses.py:111-115
            z = r.new_zeros(1, 24)
            draw = [z, z.clone(), z.clone()]
        else:
            # decode the foreground anchors only (yolo_head_24p.py:233-235)
            B, A = asg.fg_mask.shape
            idx = fg.nonzero().flatten()
            bi, ai = idx // A, idx % A
            rows = []
            off = 0
            gx, gy, gs, _, _ = self._engine.grids.get(outputs_train[0], outputs_train[1], outputs_train[2], raw.device)
            for reg in raw.reg:
                n = reg.shape[2] * reg.shape[3]
                m = (ai >= off) & (ai < off + n)
                rows.append((m, reg.detach().flatten(2)[bi[m], :, ai[m] - off]))
                off += n
            dec = r.new_empty(nfg, 26)
            for m, v in rows:
                dec[m] = v
            dec[:, 0] = (dec[:, 0] + gx[ai]) * gs[ai]
            dec[:, 1] = (dec[:, 1] + gy[ai]) * gs[ai]
            dec[:, 2:] = torch.exp(dec[:, 2:]) * gs[ai][:, None]
            draw = [dec[:, 0], dec[:, 1], dec[:, 2:26]]
        draw += [r[28:52], r[52], r[53]]
        ratio = float(r[27].detach())
        self.check_errors()
        return (r[0], r[1:25], r[25], r[26], 0.0, ratio, draw)

    # -- per-image API (losses.py:359-442) -----------------------------------------------------------
    @torch.no_grad()
    def get_assignments(self, batch_idx, num_gt, total_num_anchors, gt_bboxes_per_image, gt_classes,
                        bboxes_preds_per_image, expanded_strides, x_shifts, y_shifts, cls_preds, bbox_preds,
                        obj_preds):
        """Same contract as the reference; the three prediction views must be slices of one
        ``[B, A, 27+nc]`` buffer (as ``Loss_Function.forward`` makes them, losses.py:185-187)."""
        nc = cls_preds.shape[-1]
        base = bbox_preds[batch_idx]
        A = base.shape[0]
        row_stride = base.stride(0)
        ok = (cls_preds[batch_idx].data_ptr() == base.data_ptr() + 27 * 4 and row_stride >= 27 + nc
              and obj_preds[batch_idx].data_ptr() == base.data_ptr() + 26 * 4)
        if ok:
            image = torch.as_strided(base, (1, A, 27 + nc), (A * row_stride, row_stride, 1))
        else:  # independent tensors: assemble the head layout once
            image = torch.cat([bbox_preds[batch_idx], obj_preds[batch_idx].reshape(A, 1), cls_preds[batch_idx]],
                              1).unsqueeze(0)
        labels = torch.cat([gt_classes.reshape(-1, 1).float(), gt_bboxes_per_image.float()], 1)[:num_gt].unsqueeze(0)
        asg = self._engine.run(image, labels.contiguous(), x_shifts, y_shifts, expanded_strides, nc, want_sums=False,
                               flags=F_ALL_ROWS)
        fg_mask = asg.fg_mask[0].bool()
        matched = asg.matched_gt[0][fg_mask].long()
        ious = asg.pred_iou[0][fg_mask]
        num_fg = int(asg.num_fg[0])
        self.last_assignment = asg
        return gt_classes[matched], fg_mask, ious, matched, num_fg

    # -- losses.py:444-494 on materialised matrices ------------------------------------------------------
    @torch.no_grad()
    def dynamic_k_matching(self, cost, pair_wise_ious, gt_classes, num_gt, fg_mask):
        """Same contract as the reference: returns (num_fg, gt_matched_classes, pred_ious_this_matching,
        matched_gt_inds) and updates ``fg_mask`` in place (``fg_mask[fg_mask.clone()] = fg_mask_inboxes``)."""
        lib = _lib.load()
        cost = cost.float().contiguous()
        ious = pair_wise_ious.float().contiguous()
        _check_cuda_f32(cost, "cost")
        G, P = cost.shape
        dev = cost.device
        fg_in = torch.empty(P, dtype=torch.uint8, device=dev)
        matched = torch.empty(P, dtype=torch.int32, device=dev)
        miou = torch.empty(P, dtype=torch.float32, device=dev)
        dyn_k = torch.empty(G, dtype=torch.int32, device=dev)
        nfg = torch.empty(1, dtype=torch.int32, device=dev)
        nbytes = lib.p24_dynamic_k_workspace_bytes(G, P)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.p24_dynamic_k_matching(cost.data_ptr(), ious.data_ptr(), G, P, fg_in.data_ptr(),
                                                  matched.data_ptr(), miou.data_ptr(), dyn_k.data_ptr(), nfg.data_ptr(),
                                                  (ws.data_ptr() + 255) & ~255, nbytes, _stream_ptr(dev)),
                       "p24_dynamic_k_matching")
        inb = fg_in.bool()
        fg_mask[fg_mask.clone()] = inb
        idx = matched[inb].long()
        self.last_dynamic_ks = dyn_k
        return int(nfg), gt_classes[idx], miou[inb], idx
