"""torchrun --nproc-per-node N tests/tools/dist_check.py : the sharded loss (fused peer-memory all-reduce, and NCCL) on N
GPUs equals the loss of the whole batch on one GPU (1e-5 relative), and is bit-identical on every rank."""
import os
import sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from p24 import dist as p24_dist, synth
from p24.losses import Loss_Function

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
size, B = 640, 4 * world
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
fails = 0
full = Loss_Function(80)
lfs = {"peer": p24_dist.attach(Loss_Function(80), peer=True), "nccl": p24_dist.attach(Loss_Function(80), peer=False)}
assert lfs["nccl"].peer_comm is None
for step in range(6):  # several steps: the re-weighting state and the mailbox epochs advance
    out = synth.make_head_outputs(B, size, 80, seed=50 + step).to(dev)
    lab = synth.make_labels(B, [3 + (i * 5 + step) % 17 for i in range(B)], 50, size, 80, seed=50 + step, kind="smooth").to(dev)
    want = full.forward((g[0], g[1], g[2], out, []), lab)
    o_sh, l_sh = p24_dist.shard_batch(out, lab, rank, world)
    for name, lf in lfs.items():
        got = lf.forward((g[0], g[1], g[2], o_sh, []), l_sh)
        rel = abs(float(got[0]) - float(want[0])) / abs(float(want[0]))
        t = torch.tensor([float(got[0])], device=dev, dtype=torch.float64)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = rel < 1e-5 and float(lo) == float(hi) and abs(got[5] - want[5]) < 1e-6
        fails += 0 if ok else 1
        if rank == 0:
            print(f"step {step} {name:5s} fused={lf.peer_comm is not None} loss {float(got[0]):.6f} full-batch {float(want[0]):.6f} rel {rel:.2e} "
                  f"same on all ranks {float(lo) == float(hi)} {'ok' if ok else 'FAIL'}")
# back-to-back asynchronous steps: the collect kernels run on the side stream while the next chains are enqueued (the
# mailbox slot sets and the one-step run-ahead limit are exercised); peer and NCCL must end in the same state
a_peer, a_nccl = p24_dist.attach(Loss_Function(80), peer=True), p24_dist.attach(Loss_Function(80), peer=False)
a_peer.pipelined = True   # (the shards are resident before the first step: the steps of the fused path are pipelined)
outs = [synth.make_head_outputs(B, size, 80, seed=70 + i).to(dev) for i in range(3)]
labs = [synth.make_labels(B, 11, 50, size, 80, seed=70 + i, kind="smooth").to(dev) for i in range(3)]
res = {}
for name, lf in (("peer", a_peer), ("nccl", a_nccl)):
    for step in range(9):
        o_sh, l_sh = p24_dist.shard_batch(outs[step % 3], labs[step % 3], rank, world)
        r, _, _ = lf.forward_async((g[0], g[1], g[2], o_sh, []), l_sh)
    lf.wait_results()
    torch.cuda.synchronize()
    lf.check_errors()
    res[name] = r.clone()
rel = float(((res["peer"].double() - res["nccl"].double()).abs() / res["nccl"].double().abs().clamp_min(1e-9)).max())
gat = [torch.zeros_like(res["peer"]) for _ in range(world)]
dist.all_gather(gat, res["peer"])
same = all(torch.equal(gat[0], x) for x in gat)
ok = rel < 1e-5 and same
fails += 0 if ok else 1
if rank == 0:
    print(f"async x9: peer vs nccl rel {rel:.2e}, peer result bit-identical on all ranks {same} {'ok' if ok else 'FAIL'}")
a_peer.peer_comm and a_peer.peer_comm.close()
t = torch.tensor([fails], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("DIST_CHECK", "PASS" if int(t) == 0 else "FAIL", "peer path active:", lfs["peer"].peer_comm is not None)
torch.cuda.synchronize()
dist.barrier()
lfs["peer"].peer_comm and lfs["peer"].peer_comm.close()
dist.destroy_process_group()
sys.exit(0 if int(t) == 0 else 1)
