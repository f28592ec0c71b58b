"""World-size-2 gloo test (CPU) of the multi-GPU contract: images shard by rank, the 28 loss sums are all-reduced, and
every rank's finalize gives the loss of the unsharded batch.  The per-shard sums come from the oracle here (there is no
GPU in this container); the GPU path uses the same p24.dist helpers with NCCL."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "exploration-of-potential_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import p24_oracle as orc
    from p24 import dist as p24_dist
    from p24 import synth
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        size = 256
        out = synth.make_head_outputs(3, size, 80, seed=100)
        lab = synth.make_labels(3, [6, 0, 5], 10, size, 80, seed=100, kind="smooth")
        xs, ys, ss = synth.make_grids(size)
        o_sh, l_sh = p24_dist.shard_batch(out, lab, rank, world)
        shard = orc.LossOracle(80)
        sums = shard.sums28((xs, ys, ss, o_sh.clone(), []), l_sh)
        p24_dist.allreduce_sums(sums)
        r = shard.finalize(sums[:24], sums[24], sums[25], float(sums[26]), float(sums[27]), [])
        full = orc.LossOracle(80)
        want = full.forward((xs, ys, ss, out.clone(), []), lab)
        q.put((rank, float(r[0]), float(want[0]), r[5], want[5], float((r[1] - want[1]).abs().max())))
    finally:
        dist.destroy_process_group()


def test_sharded_sums_allreduce_reproduces_unsharded_loss():
    from p24 import dist as p24_dist
    assert [p24_dist.shard_range(160, r, 8) for r in (0, 7)] == [(0, 20), (140, 160)]
    assert [p24_dist.shard_range(5, r, 2) for r in (0, 1)] == [(0, 3), (3, 5)]
    with pytest.raises(ValueError):
        p24_dist.shard_range(4, 2, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, loss, want, ratio, want_ratio, diff in res:
        assert loss == pytest.approx(want, rel=1e-5)
        assert ratio == pytest.approx(want_ratio, rel=1e-6)
        assert diff < 1e-5
    assert res[0][1] == res[1][1]  # identical on every rank
