"""GPU box: host time to enqueue one step (Loss_Function.forward_async) vs its GPU time."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import synth
from p24.losses import Loss_Function
dev = "cuda:0"
B, size = 20, 640
out = synth.make_head_outputs(B, size, 80, seed=1).to(dev)
lab = synth.make_labels(B, 20, 50, size, 80, seed=1, kind="smooth").to(dev)
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
for reuse in (False, True):
    lf = Loss_Function(80)
    if reuse:
        if not hasattr(lf, "reuse_buffers"):
            break
        lf.reuse_buffers = True
    for i in range(20):
        lf.forward_async((g[0], g[1], g[2], out, []), lab)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for i in range(n):
        lf.forward_async((g[0], g[1], g[2], out, []), lab)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"reuse_buffers={reuse}: host enqueue {1e6 * (t1 - t0) / n:.1f} us/step, total {1e6 * (t2 - t0) / n:.1f} us/step (GPU-bound if total >> enqueue)")
