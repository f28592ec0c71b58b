"""Debug: 2 ranks, map mailboxes, one forward through the fused all-reduce, dump the mailboxes."""
import os, sys, time, ctypes as C
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import numpy as np
from p24 import dist as p24_dist, synth, lib as p24_lib
from p24.losses import Loss_Function
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
print(rank, "can access peer", torch.cuda.can_device_access_peer(rank, 1 - rank), flush=True)
t0 = time.time()
lf = p24_dist.attach(Loss_Function(80), peer=True)
print(rank, "attach done in", round(time.time() - t0, 2), "s; peer comm", lf.peer_comm is not None,
      [hex(p or 0) for p in lf.peer_comm.pointers] if lf.peer_comm else None, flush=True)
size, B = 320, 2
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
out = synth.make_head_outputs(B, size, 80, seed=5 + rank).to(dev)
lab = synth.make_labels(B, 4, 10, size, 80, seed=5 + rank, kind="smooth").to(dev)
from cuda import cudart
def dump(tag):
    n = p24_lib.load().p24_comm_mailbox_bytes()
    host = np.zeros(n // 4, dtype=np.float32)
    err, = cudart.cudaMemcpy(host.ctypes.data, lf.peer_comm.pointers[rank], n, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    h = host.reshape(2, 16, 64)
    for half in range(2):
        for r in range(world):
            print(rank, tag, "half", half, "slot", r, "flag", h[half, r, 32:33].view(np.uint32)[0], "sums[24:28]", h[half, r, 24:28], flush=True)
for step in range(3):
    t0 = time.time()
    r = lf.forward_async((g[0], g[1], g[2], out, []), lab)
    torch.cuda.synchronize()
    print(rank, "step", step, "took", round(time.time() - t0, 3), "s loss", float(r[0][0]), "sums", r[2].sums28[24:28].tolist(), flush=True)
    dump(f"after step {step}")
dist.destroy_process_group()
