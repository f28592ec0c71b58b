"""Debug tool (GPU box): per-kernel CTA start / end distribution of one step from the -DP24_TIMING build
(nvcc ... -DP24_TIMING -o p24/_lib/libp24_timing.so).  Usage: timeline.py [flags] [B size G Lmax]"""
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
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if len(sys.argv) > 5:
    B, size, G, Lmax = [int(x) for x in sys.argv[2:6]]
dev = "cuda:0"
SEED0 = int(os.environ.get("P24_SEED0", "1"))  # bench.py rank r uses 1 + 1000 r
sets = [(synth.make_head_outputs(B, size, 80, seed=SEED0 + 100 * i).to(dev),
         synth.make_labels(B, G, Lmax, size, 80, seed=SEED0 + 100 * i, kind="smooth").to(dev)) for i in range(5)]
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
for i in range(10):
    lf.forward_async((g[0], g[1], g[2], sets[i % 5][0], []), sets[i % 5][1], flags=flags)
torch.cuda.synchronize()
buf = np.zeros((6, 4096, 20), dtype=np.uint64)
lib.p24_debug_read_timers.argtypes = [ctypes.c_void_p]
assert lib.p24_debug_read_timers(buf.ctypes.data) == 0
t = buf.astype(np.int64)
# (timer bank, name, slots in order)
order = [(3, "k_gt_prep", ["start", "end"]),
         (0, "k_pass/anchor", ["start", "wait", "recs+rows", "gt loop", "items", "end"]),
         (4, "k_pass/window", ["start", "end"]),
         (1, "k_match", ["start", "wait", "load", None, "bracket", "dyn_k", "select"]),
         (2, "k_resolve_loss", ["start", "wait", None, None, "entries", "partials"])]
base = None
pct = lambda x: "min %7.1f p10 %7.1f p50 %7.1f p90 %7.1f max %7.1f" % tuple(np.percentile(x, [0, 10, 50, 90, 100]))
for k, nm, slots in order:
    tt = t[k]
    last = max(i for i, s in enumerate(slots) if s)
    ok = (tt[:, 0] > 0) & (tt[:, last] > 0)
    if not ok.any():
        print(nm, "no data"); continue
    if base is None:
        base = tt[ok, 0].min()
    print(f"== {nm}: {int(ok.sum())} CTAs")
    print(f"   start  {pct((tt[ok, 0] - base) / 1e3)}")
    print(f"   end    {pct((tt[ok, last] - base) / 1e3)}")
    prev = 0
    for i in range(1, last + 1):
        if not slots[i]:
            continue
        v = ok & (tt[:, i] > 0) & (tt[:, prev] > 0)
        d = (tt[v, i] - tt[v, prev]) / 1e3
        if d.size:
            print(f"   {slots[i]:10s} mean {d.mean():7.2f}  {pct(d)}")
        prev = i
    if k == 1:
        slow = tt[:, 8]
        sl = ok & (slow != 0)
        print("   slow-path CTAs:", int(sl.sum()), "kinds", slow[sl].tolist())
        f32 = lambda x: np.array(x, dtype=np.uint64).astype(np.uint32).view(np.float32)
        for ci in np.nonzero(sl)[0]:
            print("      slow GT", int(ci), "L", f32(t[1, ci, 12]), "U", f32(t[1, ci, 13]), "tmax", f32(t[1, ci, 14]), "rgmax", f32(t[1, ci, 15]))
        for ci in np.nonzero(sl)[0]:
            r = t[1, ci]
            print(f"      slow GT {int(ci)}: segments {r[9]} candidates {r[10]} examined segments {r[11]} survivors {r[7]} | seg scan {(r[16]-r[4])/1e3:.1f} bounds {(r[17]-r[16])/1e3:.1f} "
                  f"threshold {(r[18]-r[17])/1e3:.1f} exact {(r[19]-r[18])/1e3:.1f} rest {(r[5]-r[19])/1e3:.1f} us")
        wid = f32(t[1][ok][:, 13]) - f32(t[1][ok][:, 12])
        print("   bracket width U-L quantiles (10/50/90/99 %):", np.percentile(wid, [10, 50, 90, 99]))
