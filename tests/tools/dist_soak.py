"""torchrun --nproc-per-node N tests/tools/dist_soak.py [steps]: many thousand pipelined steps through the fused peer
all-reduce (mailbox slot sets, flow control, side-stream collect) -- hang / drift detection; the final state must equal the
NCCL path's after the same sequence."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from p24 import dist as p24_dist, synth  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
size, B = 640, 4 * world
xs, ys, ss = synth.make_grids(size, device=dev)
outs = [synth.make_head_outputs(B, size, 80, seed=50 + i).to(dev) for i in range(3)]
labs = [synth.make_labels(B, 9 + 5 * i, 50, size, 80, seed=50 + i, kind="smooth").to(dev) for i in range(3)]
shards = [p24_dist.shard_batch(o, l, rank, world) for o, l in zip(outs, labs)]
res = {}
for name, peer, n in (("peer", True, steps), ("nccl", False, min(steps, 600))):
    lf = p24_dist.attach(Loss_Function(80), peer=peer)
    lf.pipelined = peer
    lf.reuse_buffers = True
    for s in range(n):
        o, l = shards[s % 3]
        r = lf.forward_async((xs, ys, ss, o, []), l)[0]
        if s == min(steps, 600) - 1:
            lf.wait_results()
            res[name] = r.clone()
        if s % 1000 == 999:
            lf.wait_results()
            torch.cuda.synchronize()
    lf.wait_results()
    torch.cuda.synchronize()
    lf.check_errors()
    if rank == 0:
        print(f"{name}: {n} steps done, loss {float(r[0]):.6f}")
rel = float(((res["peer"] - res["nccl"]).abs() / (res["nccl"].abs() + 1e-12)).max())
flag = torch.tensor([1.0 if rel < 1e-5 else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"state after {min(steps, 600)} steps: peer vs nccl rel {rel:.2e}")
    print("DIST_SOAK", "PASS" if float(flag) == 1.0 else "FAIL")
dist.barrier()
dist.destroy_process_group()
