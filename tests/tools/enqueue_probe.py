"""Debug: CPU enqueue time vs GPU time per step of the training-shaped workload (B=20), with and without PDL."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import synth
from p24.losses import Loss_Function
B, size, G, Lmax = 20, 640, 20, 50
dev = "cuda:0"
sets = [(synth.make_head_outputs(B, size, 80, seed=1 + 100 * i).to(dev),
         synth.make_labels(B, G, Lmax, size, 80, seed=1 + 100 * i, kind="smooth").to(dev)) for i in range(5)]
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
for flags, name in [(0, "pdl"), (8, "no-pdl")]:
    lf = Loss_Function(80)
    for i in range(5):
        lf.forward_async((g[0], g[1], g[2], sets[i % 5][0], []), sets[i % 5][1], flags=flags)
    torch.cuda.synchronize()
    for nset in (1, 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(100):
            lf.forward_async((g[0], g[1], g[2], sets[i % nset][0], []), sets[i % nset][1], flags=flags)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{name} nset={nset}: cpu enqueue {1e4 * (t1 - t0):.1f} us/step, gpu {10 * e0.elapsed_time(e1):.1f} us/step")
    # per-set GPU time with a sync between steps
    for k in range(5):
        ts = []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lf.forward_async((g[0], g[1], g[2], sets[k][0], []), sets[k][1], flags=flags); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"   set {k}: isolated step {min(ts):.1f} us (min of 5), dyn_k hist {torch.bincount(lf.last_assignment.dyn_k.flatten()).tolist()}")
