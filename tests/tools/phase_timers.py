"""Debug tool (GPU box): per-CTA phase timestamps of the SimOTA chain from the -DP24_TIMING build.
Build first: nvcc ... -DP24_TIMING -o p24/_lib/libp24_timing.so (see tests/tools/README in DESIGN.md)."""
import ctypes
import os
import sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import torch
from p24 import lib as p24_lib
from p24 import synth
lib = p24_lib.load(os.path.join(ROOT, "exploration-of-potential_b200", "p24", "_lib", "libp24_timing.so"))
p24_lib._LIB = lib
from p24.losses import Loss_Function

B, size, G, Lmax = 20, 640, 20, 50
dev = "cuda:0"
sets = [(synth.make_head_outputs(B, size, 80, seed=1 + 100 * i).to(dev),
         synth.make_labels(B, G, Lmax, size, 80, seed=1 + 100 * i, kind="smooth").to(dev)) for i in range(3)]
xs, ys, ss = synth.make_grids(size)
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
sets += [(synth.make_head_outputs(B, size, 80, seed=1 + 100 * i).to(dev),
          synth.make_labels(B, G, Lmax, size, 80, seed=1 + 100 * i, kind="smooth").to(dev)) for i in range(3, 5)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for i in range(10):
    o, l = sets[(i % 5) if only is None else only]
    lf.forward_async((g[0], g[1], g[2], o, []), l)
torch.cuda.synchronize()
buf = np.zeros((6, 4096, 20), dtype=np.uint64)
lib.p24_debug_read_timers.argtypes = [ctypes.c_void_p]
assert lib.p24_debug_read_timers(buf.ctypes.data) == 0
t = buf.astype(np.int64)
names = {0: ("k_anchor_pass", 660, ["start", "pdl", "rows+recs", "pass1/2 gen", "items", "end"]),
         3: ("-", 0, []), 4: ("-", 0, []), 5: ("-", 0, []),
         1: ("k_dyn_k", 400, ["start", "pdl", "load", "-", "bracket", "dyn_k"]),
         2: ("k_resolve_loss", 320, ["start", "pdl", "-", "-", "entries", "partials", "last"])}
base = t[0, :660, 0].min()
for k, (nm, ncta, ph) in names.items():
    if not ncta:
        continue
    tt = t[k, :ncta, :len(ph)]
    ok = tt[:, len(ph) - 2] > 0
    print(f"== {nm}: first start {(tt[ok, 0].min() - base) / 1e3:.1f} us, last end {(tt[ok].max() - base) / 1e3:.1f} us after chain start")
    for i in range(1, len(ph)):
        if ph[i] == "-":
            continue
        j = i - 1
        while ph[j] == "-":
            j -= 1
        valid = ok & (tt[:, i] > 0) & (tt[:, j] > 0)
        d = (tt[valid, i] - tt[valid, j]) / 1e3
        if d.size:
            print(f"   {ph[i]:14s} mean {d.mean():7.2f} us  p50 {np.median(d):7.2f}  max {d.max():7.2f}  (n={d.size})")
    if k == 1:
        slow = t[1, :ncta, 8]
        print("   slow-path CTAs:", int((slow != 0).sum()), "of", ncta, "kinds", slow[slow != 0].tolist(),
              "nhit", t[1, :ncta, 9][slow != 0].tolist(), "nev", t[1, :ncta, 10][slow != 0].tolist(),
              "overflow", t[1, :ncta, 11][slow != 0].tolist())
        full = t[1, :ncta]
        for ci in np.nonzero(slow != 0)[0][:6]:
            r = full[ci]
            print("      slow CTA", int(ci), "dyn_k phases us: tau", (r[12] - r[4]) / 1e3, "scan", (r[13] - r[12]) / 1e3,
                  "ub", (r[14] - r[13]) / 1e3, "exact", (r[15] - r[14]) / 1e3, "select", (r[5] - r[15]) / 1e3,
                  "| ends at", (r[5] - base) / 1e3)
        tot = (tt[ok, 5] - tt[ok, 1]) / 1e3
        print(f"   CTA total mean {tot.mean():.2f} us max {tot.max():.2f} us; slow ones: {np.round(tot[slow[ok] != 0], 1).tolist()[:20]}")
