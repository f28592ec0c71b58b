"""Debug: eager vs CUDA-graph replay of the loss step (B=20)."""
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
lf = Loss_Function(80)
for i in range(5):
    lf.forward_async((g[0], g[1], g[2], sets[i][0], []), sets[i][1])
torch.cuda.synchronize()
def timeit(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("eager us/step", timeit(lambda i: lf.forward_async((g[0], g[1], g[2], sets[i % 5][0], []), sets[i % 5][1])))
graphs = []
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(5):
        lf.forward_async((g[0], g[1], g[2], sets[i][0], []), sets[i][1])
    torch.cuda.synchronize()
    for i in range(5):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            out = lf.forward_async((g[0], g[1], g[2], sets[i][0], []), sets[i][1])
        graphs.append((gr, out))
torch.cuda.synchronize()
print("graph us/step", timeit(lambda i: graphs[i % 5][0].replay()))
print("loss", float(graphs[0][1][0][0]))
