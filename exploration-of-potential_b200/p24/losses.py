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
from .engine import Assignment, SimOTAEngine, F_ALL_ROWS


class Loss_Function(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes
        self.use_l1 = False  # never enabled by the 24p scripts (losses.py:163); the L1 branch is not provided
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
        state = self._state(outputs.device)
        if self.process_group is None:
            # single GPU: the last CTA of the chain applies the normalisation and re-weighting itself
            result54 = torch.empty(54, dtype=torch.float32, device=outputs.device)
            weights27 = torch.empty(27, dtype=torch.float32, device=outputs.device)
            asg = self._engine.run(outputs, labels, x_shifts, y_shifts, expanded_strides, self.num_classes,
                                   flags=flags, finalize=(state, result54, weights27))
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

    def forward(self, outputs_train, labels):
        outputs = outputs_train[3]
        result54, weights27, asg = self.forward_async(outputs_train, labels)
        r = result54
        fg = asg.fg_mask.view(-1).bool()
        rows = outputs.reshape(-1, outputs.shape[-1])[fg]
        if rows.shape[0] == 0:  # losses.py:111-115
            draw = [rows.new_zeros(1, 24), rows.new_zeros(1, 24), rows.new_zeros(1, 24)]
        else:
            draw = [rows[:, 0], rows[:, 1], rows[:, 2:26]]
        draw += [r[28:52], r[52], r[53]]
        ratio = float(r[27])  # one D2H read per step (the reference returns a Python float here)
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
