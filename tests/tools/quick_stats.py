"""Quick look at one workload on the GPU: per-kernel CUDA-event times, rare-path counters, step time.
Usage: python tests/tools/quick_stats.py [train|train_spiky|crowded|hires] [steps] [raw]
(raw: feed the head's raw per-level conv outputs, p24.engine.RawLevels, instead of the decoded buffer)"""
import ctypes
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from p24 import lib as p24_lib, synth  # noqa: E402
from p24.engine import RawLevels  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "train"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wl = bench.TRAIN_WORKLOADS[name]
B = wl["B"] or 20
dev = torch.device("cuda:0")
RAW = len(sys.argv) > 3 and sys.argv[3] == "raw"
sets = []
for i in range(5 if wl["size"] == 640 else 2):
    s = wl["seed"] + 100 * i
    if RAW:
        r, o, c = synth.make_raw_levels(B, wl["size"], 80, seed=s)
        head = RawLevels([t.to(dev) for t in r], [t.to(dev) for t in o], [t.to(dev) for t in c])
    else:
        head = synth.make_head_outputs(B, wl["size"], 80, seed=s).to(dev)
    sets.append((head,
                 synth.make_labels(B, wl["G"], wl["Lmax"], wl["size"], 80, seed=s, kind=wl["kind"]).to(dev)))
xs, ys, ss = synth.make_grids(wl["size"])
g = [[t.to(dev) for t in l] for l in (xs, ys, ss)]
lf = Loss_Function(80)
lf.reuse_buffers = True
lf.pipelined = bool(int(os.environ.get("P24_PIPELINED", "1")))
lib = p24_lib.load()
for i in range(10):
    lf.forward_async((g[0], g[1], g[2], sets[i % len(sets)][0], []), sets[i % len(sets)][1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    lf.forward_async((g[0], g[1], g[2], sets[i % len(sets)][0], []), sets[i % len(sets)][1])
e1.record()
torch.cuda.synchronize()
print(f"{name}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step  ({B * steps / e0.elapsed_time(e1) * 1e3:.0f} img/s)")
lib.p24_profile_enable(1)
buf = (ctypes.c_float * 8)()
acc = [0.0] * 8
for i in range(steps):
    lf.forward_async((g[0], g[1], g[2], sets[i % len(sets)][0], []), sets[i % len(sets)][1])
    lib.p24_profile_read(buf)
    for k in range(8):
        acc[k] += buf[k]
lib.p24_profile_enable(0)
print("kernel us:", {n: round(a / steps * 1e3, 1) for n, a in zip(["k_prep", "k_pass", "k_tail"], acc)})
print("stats:", lf.path_stats())
