"""Soak run of the pipelined training chain: 300 steps compared bit for bit with the plain chain, then many thousand
steps back to back (hang / error-bit detection).  Usage: python tests/tools/soak.py [steps]"""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from p24 import synth  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = "cuda:0"
xs, ys, ss = synth.make_grids(640, device=dev)
sets = []
for i, (counts, kind) in enumerate([([20] * 20, "smooth"), ([3, 0, 50, 7, 1, 12] * 3 + [20, 20], "spiky"), ([33] * 20, "smooth"),
                                    ([50] * 20, "spiky")]):
    sets.append((synth.make_head_outputs(20, 640, 80, seed=90 + i).to(dev),
                 synth.make_labels(20, counts, 50, 640, 80, seed=90 + i, kind=kind).to(dev)))
plain, piped = Loss_Function(80), Loss_Function(80)
piped.pipelined = True
piped.reuse_buffers = True
want = []
for s in range(300):
    o, l = sets[s % 4]
    want.append(plain.forward_async((xs, ys, ss, o, []), l)[0].clone())
torch.cuda.synchronize()
got = []
for s in range(300):
    o, l = sets[s % 4]
    got.append(piped.forward_async((xs, ys, ss, o, []), l)[0].clone())
torch.cuda.synchronize()
bad = [s for s in range(300) if not torch.equal(want[s], got[s])]
print("300 pipelined steps vs plain: mismatching steps", bad[:10], "OK" if not bad else "FAIL")
t0 = time.perf_counter()
for s in range(steps):
    o, l = sets[s % 4]
    r = piped.forward_async((xs, ys, ss, o, []), l)[0]
    if s % 2000 == 1999:
        torch.cuda.synchronize()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
piped.check_errors()
print(f"{steps} steps back to back: {dt / steps * 1e6:.1f} us/step, last loss {float(r[0]):.6f}, status {piped.read_status()}")
print("SOAK", "PASS" if not bad and torch.isfinite(r).all() else "FAIL")
