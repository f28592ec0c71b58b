"""A small tour of every kernel for compute-sanitizer (memcheck / racecheck): the training chain on decoded rows and on raw
head planes (pipelined steps), backward, the postprocess at the three settings (decoded and raw input), label packing.
Usage: python tests/tools/sanitize_step.py   (or under compute-sanitizer --tool memcheck where the pool allows it)"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from p24 import boxes, head, synth  # noqa: E402
from p24.data import TrainTransform  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402

dev = "cuda:0"
size = 320
B = 3
xs, ys, ss = synth.make_grids(size, device=dev)
lf = Loss_Function(80)
lf.pipelined = True
for step in range(3):
    out = synth.make_head_outputs(B, size, 80, seed=step).to(dev)
    lab = synth.make_labels(B, [7, 0, 30], 40, size, 80, seed=step, kind="spiky" if step == 1 else "smooth").to(dev)
    r, _, a = lf.forward_async((xs, ys, ss, out, []), lab)
torch.cuda.synchronize()
print("rows loss", float(r[0]), a.num_fg.tolist())
reg, obj, cls = synth.make_raw_levels(B, size, 80, seed=5, device=dev)
leaves = [[t.clone().requires_grad_(True) for t in lst] for lst in (reg, obj, cls)]
res = Loss_Function(80).forward(head.train_outputs(*leaves, synth.STRIDES), lab)
res[0].backward()
print("raw loss", float(res[0]), float(leaves[0][0].grad.abs().sum()))
o2 = out.clone().requires_grad_(True)
Loss_Function(80).forward((xs, ys, ss, o2, []), lab)[0].backward()
for t in obj + cls:
    t += 4.5
for conf, nms, ag in [(0.25, 0.45, False), (0.01, 0.3, True), (0.01, 0.65, False)]:
    a1 = boxes.postprocess(head.infer_outputs(reg, obj, cls, synth.STRIDES), 80, conf, nms, ag)
    a2 = boxes.postprocess(head.infer_outputs(reg, obj, cls, synth.STRIDES, fused=False), 80, conf, nms, ag)
    same = all((x is None) == (y is None) and (x is None or torch.equal(x, y)) for x, y in zip(a1, a2))
    print("post", conf, nms, ag, [None if x is None else x.shape[0] for x in a1], "raw == decoded:", same)
rng = np.random.default_rng(0)
t = [np.concatenate([rng.integers(0, 80, (n, 1)).astype(np.float64), rng.random((n, 50))], 1) if n else np.zeros((1, 0))
     for n in (5, 0, 60)]
print("pack", TrainTransform(50).pack(t, [(300, 320)] * 3, (320, 320), dev).sum().item())
