"""Debug tool (GPU box): per-kernel [first CTA start, last CTA end] of one step from the -DP24_TIMING build."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import lib as p24_lib
from p24 import synth
lib = p24_lib.load(os.path.join(ROOT, "exploration-of-potential_b200", "p24", "_lib", "libp24_timing.so"))
p24_lib._LIB = lib
from p24.losses import Loss_Function
B, size, G, Lmax = 20, 640, 20, 50
dev = "cuda:0"
sets = [(synth.make_head_outputs(B, size, 80, seed=1 + 100 * i).to(dev),
         synth.make_labels(B, G, Lmax, size, 80, seed=1 + 100 * i, kind="smooth").to(dev)) for i in range(5)]
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for i in range(10):
    lf.forward_async((g[0], g[1], g[2], sets[i % 5][0], []), sets[i % 5][1], flags=flags)
torch.cuda.synchronize()
buf = np.zeros((6, 4096, 20), dtype=np.uint64)
lib.p24_debug_read_timers.argtypes = [ctypes.c_void_p]
assert lib.p24_debug_read_timers(buf.ctypes.data) == 0
t = buf.astype(np.int64)
order = [(3, "k_gt_prep", 2), (0, "k_anchor_pass", 5), (1, "k_dyn_k", 5), (4, "k_window_eval", 2), (5, "k_select", 2), (2, "k_resolve_loss", 5)]
base = None
for k, nm, last in order:
    tt = t[k]
    ok = (tt[:, 0] > 0) & (tt[:, last] > 0)
    if not ok.any():
        print(nm, "no data"); continue
    st, en = tt[ok, 0].min(), tt[ok][:, 1:last + 1].max()
    if base is None:
        base = st
    work = tt[ok, last] - tt[ok, 1]
    print(f"{nm:16s} first start {(st - base) / 1e3:7.1f}  after-wait start {(tt[ok, 1].min() - base) / 1e3:7.1f}  last end {(en - base) / 1e3:7.1f}"
          f"   CTAs {int(ok.sum()):5d}  per-CTA work mean {work.mean() / 1e3:6.2f} max {work.max() / 1e3:6.2f} us")
