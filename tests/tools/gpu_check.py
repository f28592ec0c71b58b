"""Diagnostic run on a GPU box: product path vs the oracle (on CUDA and on CPU), with mismatch details.
Usage: python tests/tools/gpu_check.py [--size 640] [--batch 2] [--gt 20] [--kind smooth] [--seed 0]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)

from p24 import synth  # noqa: E402
from p24.losses import Loss_Function  # noqa: E402
from p24 import engine as eng  # noqa: E402
from oracle import p24_oracle as orc  # noqa: E402


def compare(asg, trace, tag):
    B = len(trace)
    bad = 0
    for b in range(B):
        tr = trace[b]
        fg_o = tr["fg_mask"].cpu().numpy()
        fg_m = asg.fg_mask[b].bool().cpu().numpy()
        nd = int((fg_o != fg_m).sum())
        msg = f"[{tag}] img {b}: num_gt={tr['num_gt']} fg oracle={fg_o.sum()} mine={fg_m.sum()} fg_diff={nd}"
        if tr["num_gt"]:
            both = fg_o & fg_m
            mo = np.full(fg_o.shape, -1, np.int64)
            mo[fg_o] = tr["matched"].cpu().numpy()
            mm = asg.matched_gt[b].cpu().numpy()
            md = int((mo[both] != mm[both]).sum())
            io = np.zeros(fg_o.shape, np.float32)
            io[fg_o] = tr["ious"].cpu().numpy()
            im = asg.pred_iou[b].cpu().numpy()
            rel = np.abs(io[both] - im[both]) / np.maximum(np.abs(io[both]), 1e-12)
            dk_o = np.array(tr["dyn_k"])
            dk_m = asg.dyn_k[b, :tr["num_gt"]].cpu().numpy()
            msg += f" matched_diff={md} iou_maxrel={rel.max() if rel.size else 0:.2e} dyn_k_diff={int((dk_o != dk_m).sum())} dyn_k={dk_m.tolist()[:8]} vs {dk_o.tolist()[:8]}"
            bad += md + int((dk_o != dk_m).sum())
        bad += nd
        print(msg)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--gt", type=int, default=20)
    ap.add_argument("--lmax", type=int, default=50)
    ap.add_argument("--kind", default="smooth")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-oracle", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    out = synth.make_head_outputs(args.batch, args.size, 80, seed=args.seed)
    lab = synth.make_labels(args.batch, args.gt, max(args.lmax, args.gt), args.size, 80, seed=args.seed, kind=args.kind)
    xs, ys, ss = synth.make_grids(args.size)
    outd, labd = out.to(dev), lab.to(dev)
    xsd, ysd, ssd = [t.to(dev) for t in xs], [t.to(dev) for t in ys], [t.to(dev) for t in ss]

    lf = Loss_Function(80)
    total_bad = 0
    for flags, name in [(0, "default"), (eng.F_NO_PRUNE | eng.F_NO_FILTER, "no-prune/no-filter")]:
        lf2 = Loss_Function(80)
        torch.cuda.synchronize()
        t0 = time.time()
        res, w, asg = lf2.forward_async((xsd, ysd, ssd, outd, []), labd, flags=flags)
        torch.cuda.synchronize()
        print(f"== product path ({name}) first call {1e3 * (time.time() - t0):.2f} ms; loss={float(res[0]):.6f} "
              f"num_fg={asg.num_fg.tolist()} sums[24:28]={asg.sums28[24:].tolist()}")
        if flags == 0:
            base = (res.clone(), asg)
        else:
            same = (torch.equal(base[1].fg_mask, asg.fg_mask) and torch.equal(base[1].matched_gt, asg.matched_gt)
                    and torch.equal(base[1].dyn_k, asg.dyn_k) and torch.equal(base[1].pred_iou, asg.pred_iou))
            print("   pruned == unpruned:", same)
            total_bad += 0 if same else 1
    res, asg = base
    # oracle on the same GPU (torch CUDA eager)
    o = orc.LossOracle(80)
    t0 = time.time()
    r = o.forward((xsd, ysd, ssd, outd.clone(), []), labd)
    torch.cuda.synchronize()
    print(f"== oracle on CUDA: {time.time() - t0:.2f} s loss={float(r[0]):.6f}")
    total_bad += compare(asg, o.trace, "cuda-oracle")
    print("   loss rel diff", abs(float(r[0]) - float(res[0])) / abs(float(r[0])),
          "obj", abs(float(r[2]) - float(res[25])) / abs(float(r[2])),
          "cls", abs(float(r[3]) - float(res[26])) / max(abs(float(r[3])), 1e-12),
          "iou24 max", float(((r[1] - res[1:25]).abs() / r[1].abs().clamp_min(1e-12)).max()))
    if args.cpu_oracle:
        o2 = orc.LossOracle(80)
        t0 = time.time()
        r2 = o2.forward((xs, ys, ss, out.clone(), []), lab)
        print(f"== oracle on CPU: {time.time() - t0:.2f} s loss={float(r2[0]):.6f}")
        total_bad += compare(asg, o2.trace, "cpu-oracle")
        print("   loss rel diff", abs(float(r2[0]) - float(res[0])) / abs(float(r2[0])))
    # timing
    for _ in range(3):
        lf.forward_async((xsd, ysd, ssd, outd, []), labd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lf.forward_async((xsd, ysd, ssd, outd, []), labd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"== timing: {ms * 1e3:.1f} us / step, {args.batch / ms * 1e3:.0f} img/s")
    print("TOTAL_BAD", total_bad)


if __name__ == "__main__":
    main()
