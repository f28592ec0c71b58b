"""Phase timeline of the training chain from the debug build (P24_TIMING=1 python tests/tools/timeline.py [workload]).
%globaltimer stamps written by the kernels at phase boundaries; prints where the time of k_pass / k_tail goes."""
import ctypes
import os
import sys

os.environ["P24_TIMING"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from p24 import lib as p24_lib, synth  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "train"
wl = bench.TRAIN_WORKLOADS[name]
B = wl["B"] or 20
dev = torch.device("cuda:0")
sets = []
for i in range(3):
    s = wl["seed"] + 100 * i
    sets.append((synth.make_head_outputs(B, wl["size"], 80, seed=s).to(dev),
                 synth.make_labels(B, wl["G"], wl["Lmax"], wl["size"], 80, seed=s, kind=wl["kind"]).to(dev)))
xs, ys, ss = synth.make_grids(wl["size"])
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
lf.reuse_buffers = True
lf.pipelined = bool(int(os.environ.get("P24_PIPELINED", "1")))
lib = p24_lib.load()
for i in range(6):
    lf.forward_async((g[0], g[1], g[2], sets[i % 3][0], []), sets[i % 3][1])
torch.cuda.synchronize()
ROWS, SLOTS = 8192, 16
buf = np.zeros((3, ROWS, SLOTS), dtype=np.uint64)
fn = lib.p24_debug_read_timers
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
A = sum((wl["size"] // s) ** 2 for s in (8, 16, 32))
tiles = (A + 255) // 256
kp = buf[1].astype(np.float64)
kt = buf[2].astype(np.float64)
t0 = kp[6000:6000 + 1024, 0]
t0 = t0[t0 > 0].min()


def us(x):
    return (x - t0) / 1e3


def stat(label, d):
    d = np.asarray(d) / 1e3
    if d.size:
        print(f"   {label:18s} mean {d.mean():7.2f}  p50 {np.median(d):7.2f}  max {d.max():7.2f}  (n={d.size})")


k0 = buf[0].astype(np.float64)
rows0 = np.concatenate([k0[b * 64:b * 64 + (wl["G"] + 1) // 2] for b in range(B)])
p0 = rows0[:, 0].min()
print(f"== k_prep: first CTA start {us(p0):.1f} us, last end {us(rows0[:, 7].max()):.1f} us (k_pass past its pdl_wait at 0)")
_st, _en = np.sort(us(rows0[:, 0])), np.sort(us(rows0[:, 7]))
print("   CTA work start percentiles (0/25/50/75/100):", " ".join(f"{np.percentile(_st, q):.1f}" for q in (0, 25, 50, 75, 100)))
print("   CTA work end   percentiles (0/25/50/75/100):", " ".join(f"{np.percentile(_en, q):.1f}" for q in (0, 25, 50, 75, 100)))
for lab, a, b_ in [("count labels", 0, 1), ("records+pairs", 1, 2), ("image barrier", 2, 3), ("stage records", 3, 4),
                   ("seed points", 4, 5), ("dedupe+rank+sync", 5, 6), ("T + far2", 6, 7), ("CTA total", 0, 7)]:
    stat(lab, rows0[:, b_] - rows0[:, a])
print(f"== k_pass: CTAs past pdl_wait at 0 us; last CTA end {us(kp[6000:7024, 1].max()):.1f} us")
tl = kp[:min(B * tiles, 4095)]
print(f"   tiles: first start {us(tl[:, 0].min()):.1f}  last end {us(tl[:, 7].max()):.1f}")
if tl.shape[0] == B * tiles:
    per = (tl[:, 7] - tl[:, 0]).reshape(B, tiles) / 1e3
    st_ = ((tl[:, 0]).reshape(B, tiles) - t0) / 1e3
    print("   tile total by tile-in-image (mean us):", " ".join(f"{v:.0f}" for v in per.mean(0)))
    print("   tile start by tile-in-image (mean us):", " ".join(f"{v:.0f}" for v in st_.mean(0)))
    late = np.argsort(-(tl[:, 7]))[:8]
    print("   last tiles to end (image, tile, start, end):", [(int(k // tiles), int(k % tiles), round(us(tl[k, 0]), 1), round(us(tl[k, 7]), 1)) for k in late])
fq = kp[4096:6000]
fq = fq[fq[:, 0] > 0]
if fq.shape[0]:
    print(f"   far-queue chunks: n={fq.shape[0]} first start {us(fq[:, 0].min()):.1f} last end {us(fq[:, 7].max()):.1f}")
    stat("far chunk", fq[:, 7] - fq[:, 0])
for lab, a, b_ in [("pre", 0, 1), ("recs+rows", 1, 3), ("gt loop+lists", 3, 4), ("poly items", 4, 5), ("cand sync", 5, 2),
                   ("wait seeds", 2, 8), ("far list", 8, 6), ("far items+out", 6, 7), ("tile total", 0, 7)]:
    stat(lab, tl[:, b_] - tl[:, a])
wn = kp[3000:4095]
ok = (wn[:, 0] > 0) & (wn[:, 7] > 0)
if ok.any():
    print(f"   window chunks: first start {us(wn[ok, 0].min()):.1f}  last end {us(wn[ok, 7].max()):.1f}")
    stat("window chunk", wn[ok, 7] - wn[ok, 0])
c = kt[:B * 8]
tt0 = c[:, 0].min()
print(f"== k_tail: first CTA start {us(tt0):.1f} us, last end {us(c[:, 9].max()):.1f} us")
for lab, a, b_ in [("stage recs", 0, 1), ("phase 1 (own GTs)", 1, 2), ("cluster barrier", 2, 3), ("rare", 3, 4), ("phase 2", 4, 5),
                   ("phase 3a/b", 5, 6), ("phase 3c", 6, 7), ("reduce wait", 7, 8), ("atomics+ticket", 8, 9), ("CTA total", 0, 9)]:
    stat(lab, c[:, b_] - c[:, a])
gt = np.concatenate([kt[1024 + b * 64:1024 + b * 64 + wl["G"]] for b in range(B)]) if wl["G"] <= 64 else kt[1024:1024 + 64]
for lab, a, b_ in [("lcount RT", 0, 1), ("list pass A", 1, 2), ("threshold", 2, 3), ("pass B + exact", 3, 4), ("  B: compaction", 3, 10),
                   ("  B: rows + 1st eval", 10, 11), ("  B: rest", 11, 4), ("sum", 4, 5),
                   ("claims", 5, 6), ("GT total", 0, 6)]:
    stat(lab, gt[:, b_] - gt[:, a])
print(f"   survivors per GT mean {gt[:, 8].mean():.1f} max {gt[:, 8].max():.0f}; list length mean {gt[:, 9].mean():.1f} max {gt[:, 9].max():.0f}")
print(f"   GT start offsets from kernel start: mean {((gt[:, 0] - tt0) / 1e3).mean():.2f} max {((gt[:, 0] - tt0) / 1e3).max():.2f}")
