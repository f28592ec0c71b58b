"""Minimal driver for ncu: a few steps of the training-shaped workload (B=20, 640, 20 GT)."""
import os
import sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import torch
from p24 import synth
from p24.losses import Loss_Function

B, size, G, Lmax = 20, 640, 20, 50
if len(sys.argv) > 1:
    B, size, G, Lmax = [int(x) for x in sys.argv[1:5]]
dev = "cuda:0"
sets = []
for i in range(2):
    sets.append((synth.make_head_outputs(B, size, 80, seed=1 + 100 * i).to(dev),
                 synth.make_labels(B, G, Lmax, size, 80, seed=1 + 100 * i, kind="smooth").to(dev)))
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
for i in range(4):
    o, l = sets[i % 2]
    r = lf.forward_async((g[0], g[1], g[2], o, []), l)
torch.cuda.synchronize()
print("loss", float(r[0][0]))
