"""GPU box: timing of the postprocess (BASELINE.json configs[3]: batch 64 at 640x640) vs the oracle on the host CPU."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import synth, boxes
from oracle import p24_oracle as orc
dev = "cuda:0"
B = 64
sets = [synth.make_postprocess_input(B, 640, 80, seed=3 + i) for i in range(3)]   # 3 x 230 MB > L2
dsets = [s.to(dev) for s in sets]
for conf, nms, ag, name in [(0.25, 0.45, False, "conf 0.25 / nms 0.45 batched (~164 pre-NMS/img)"),
                            (0.01, 0.3, True, "conf 0.01 / nms 0.3 class-agnostic (show_24p.py settings)"),
                            (0.01, 0.65, False, "conf 0.01 / nms 0.65 batched (~8k pre-NMS/img)")]:
    for i in range(3):
        boxes.postprocess_raw(dsets[i], 80, conf, nms, ag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 12
    e0.record()
    for i in range(n):
        r = boxes.postprocess_raw(dsets[i % 3], 80, conf, nms, ag)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = B * 8400 * 107 * 4 / (ms * 1e-3) / 1e9
    t0 = time.perf_counter()
    for i in range(8):
        orc.postprocess_image(sets[0][i], 80, conf, nms, ag)
    cpu = 8 / (time.perf_counter() - t0)
    print(f"{name}: {ms * 1e3:.1f} us / batch of {B} -> {B / ms * 1e3:.0f} img/s, {gbs:.0f} GB/s algorithmic "
          f"({gbs / 6549.8 * 100:.1f} % of measured HBM peak); kept {r[1].float().mean().item():.0f}/img; CPU oracle {cpu:.0f} img/s")
