"""Minimal driver for ncu: the postprocess on BASELINE.json configs[3] (batch 64 at 640x640).  Usage: post_step.py [conf nms agnostic]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200")); sys.path.insert(0, ROOT)
import torch
from p24 import synth, boxes
conf, nms, ag = (float(sys.argv[1]), float(sys.argv[2]), bool(int(sys.argv[3]))) if len(sys.argv) > 3 else (0.25, 0.45, False)
dev = "cuda:0"
sets = [synth.make_postprocess_input(64, 640, 80, seed=3 + i).to(dev) for i in range(2)]
for i in range(4):
    r = boxes.postprocess_raw(sets[i % 2], 80, conf, nms, ag)
torch.cuda.synchronize()
print("kept per image", r[1].float().mean().item())
