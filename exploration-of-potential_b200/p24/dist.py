"""Multi-GPU sharding of the loss path: the batch shards by image across the GPUs of one box, one
process per GPU; the only collective is the SUM all-reduce of the 28 loss sums (24 per-ray IoU sums, obj BCE, cls BCE,
num_fg, num_gt).  Assignments are strictly per image, so no other data crosses GPUs; the postprocess needs no collective
at all.

Two ways to do the all-reduce:
  * ``PeerComm`` (default on GPUs of one box): every rank owns a small mailbox in device memory that its peers map
    through CUDA IPC; the last CTA of the kernel chain stores its sums into every peer's mailbox over NVLink / NVSwitch
    (the publish half is part of the compute kernel); a one-warp collect kernel on a side stream waits for the peers'
    flags, adds the contributions in rank order and applies the normalisation, while the next step's chain already runs
    (nothing in it depends on the global sums).  No NCCL launch.
  * NCCL ``all_reduce`` of the 28-float vector on the compute stream followed by ``p24_loss_finalize`` (fallback when
    the mailboxes cannot be mapped; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``total`` images for ``rank`` (earlier ranks take the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(outputs: torch.Tensor, labels: torch.Tensor, rank: int, world: int):
    lo, hi = shard_range(outputs.shape[0], rank, world)
    return outputs[lo:hi], labels[lo:hi]


def allreduce_sums(sums28: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the 28-float vector (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    if sums28.numel() != 28:
        raise ValueError("expected the 28 loss sums")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums28, op=dist.ReduceOp.SUM, group=group)
    return sums28


class PeerComm:
    """Mailboxes of the fused all-reduce (``include/p24.h``: p24_comm_*), one per rank, mapped into this process."""

    def __init__(self, group=None):
        import ctypes as C
        from . import lib as _lib
        lib = _lib.load()
        self._lib = lib
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        if self.nranks > 16:
            raise RuntimeError("PeerComm serves at most 16 ranks (one box)")
        own = C.c_void_p()
        _lib.check(lib.p24_comm_alloc(C.byref(own)), "p24_comm_alloc")
        self._own = own
        handle = C.create_string_buffer(64)
        _lib.check(lib.p24_comm_export(own, handle), "p24_comm_export")
        handles = [None] * self.nranks
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._peers = []
        ptrs = (C.c_void_p * self.nranks)()
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs[r] = own.value
            else:
                peer = C.c_void_p()
                _lib.check(lib.p24_comm_import(C.create_string_buffer(h, 64), C.byref(peer)), "p24_comm_import")
                self._peers.append(peer)
                ptrs[r] = peer.value
        self.pointers = ptrs
        self._epoch = 0
        # the collect half of the exchange (p24_comm_finish) runs on a side stream with no stream dependency on the chain (it
        # waits for a flag of the chain's last CTA): the next step's kernels do not depend on the global sums, and no event
        # sits between two steps on the compute stream (that would break their programmatic overlap)
        self.side = torch.cuda.Stream()
        self._fin_ev = torch.cuda.Event()
        self._dirty = False   # a collect kernel has been enqueued that the compute stream is not ordered behind yet
        dist.barrier(group=group)  # every mailbox is mapped everywhere before the first kernel writes to it

    def close(self):
        """Unmap the peers' mailboxes and free the own one (call on every rank, after the last step has completed)."""
        if self._own is None:
            return
        torch.cuda.synchronize()
        for peer in self._peers:
            self._lib.p24_comm_close(peer)
        self._peers = []
        self._lib.p24_comm_free(self._own)
        self._own = None

    def next_epoch(self) -> int:
        self._epoch = (self._epoch + 1) & 0xFFFFFFFF or 1
        return self._epoch

    def finish(self, sums28, state26, result54, weights27, ws_ptr, B, A, Lmax):
        """Enqueue the collect kernel of the current epoch on the side stream (call it once per step, right after the
        chain has been enqueued)."""
        from . import lib as _lib
        code = self._lib.p24_comm_finish(self._own, self.nranks, self._epoch, sums28.data_ptr(),
                                         state26.data_ptr() if state26 is not None else None,
                                         result54.data_ptr() if result54 is not None else None,
                                         weights27.data_ptr() if weights27 is not None else None,
                                         ws_ptr, B, A, Lmax, self.side.cuda_stream)
        _lib.check(code, "p24_comm_finish")
        self._dirty = True

    def wait(self):
        """Order the current stream behind every collect kernel enqueued so far (results are then safe to read)."""
        if self._dirty:
            self._fin_ev.record(self.side)
            torch.cuda.current_stream().wait_event(self._fin_ev)
            self._dirty = False


def attach(loss_function, group=None, peer: bool = True):
    """Make ``Loss_Function.forward`` shard-aware: the 28 sums are all-reduced over ``group`` (default WORLD) before the
    normalisation, so every rank computes the same loss / weights / state.  ``peer=True`` uses the fused peer-memory
    all-reduce when the ranks' mailboxes can be mapped (GPUs of one box) and says so on stderr when it cannot."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    loss_function.process_group = group if group is not None else dist.group.WORLD
    loss_function.peer_comm = None
    if peer and torch.cuda.is_available() and dist.get_backend(group) == "nccl":
        ok = torch.ones(1, device="cuda")
        comm = None
        try:
            comm = PeerComm(group)
        except Exception as exc:  # IPC not permitted / no peer access: every rank must take the same path
            import sys
            print(f"p24.dist: peer-memory all-reduce unavailable ({exc}); using NCCL", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        loss_function.peer_comm = comm if float(ok) > 0 else None
    return loss_function
