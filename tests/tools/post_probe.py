"""Per-step timing of the postprocess with and without a synchronisation between the steps."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import synth, boxes
conf, nms, ag = (float(sys.argv[1]), float(sys.argv[2]), bool(int(sys.argv[3]))) if len(sys.argv) > 3 else (0.01, 0.65, False)
dev = "cuda:0"
sets = [synth.make_postprocess_input(64, 640, 80, seed=3 + 100 * i).to(dev) for i in range(2)]
for i in range(5):
    r = boxes.postprocess_raw(sets[i % 2], 80, conf, nms, ag)
torch.cuda.synchronize()
ts = []
for i in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = boxes.postprocess_raw(sets[i % 2], 80, conf, nms, ag)
    e1.record()
    torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1) * 1e3))
print("synced steps us:", ts)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    r = boxes.postprocess_raw(sets[i % 2], 80, conf, nms, ag)
e1.record()
torch.cuda.synchronize()
print("back-to-back us/step:", round(e0.elapsed_time(e1) / 20 * 1e3))
import time
t0 = time.perf_counter()
for i in range(20):
    r = boxes.postprocess_raw(sets[i % 2], 80, conf, nms, ag)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue us/step:", round((t1 - t0) / 20 * 1e6), " total wall us/step:", round((time.perf_counter() - t0) / 20 * 1e6))
