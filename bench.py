#!/usr/bin/env python
"""bench.py — 24p loss + SimOTA images/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|crowded|hires]

A step = one pass of the hot path (fused SimOTA assignment + loss sums + re-weighting, i.e. the
reference's ``Loss_Function.forward``) over one batch of synthetic head outputs.  Default workload =
BASELINE.json configs[1]: batch 20 per GPU at 640x640 (8400 anchors), 20 GT/img, 80 classes.

  value     images/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e       the same metric through the public ``Loss_Function.forward`` with HOST (pinned) inputs:
            H2D of the head output + labels and the D2H read of the loss inside the timed region
  roofline  the dominant kernel's algorithmic bytes / its live CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle restatement of the reference (torch CPU eager, all host threads) on a bounded sample

Multi-GPU (torchrun, one rank per GPU): images shard by rank (weak scaling, 20 images per GPU); the only
collective is the 28-float all-reduce of the loss sums (NCCL), between the sums kernel and the finalize
kernel on the compute stream.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "exploration-of-potential_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (per-GPU batch, image size, GT per image, Lmax, label kind, seed)
    "train": (20, 640, 20, 50, "smooth", 1),     # BASELINE.json configs[1]
    "crowded": (20, 640, 100, 100, "smooth", 2),  # configs[2]
    "hires": (20, 1280, 20, 50, "smooth", 4),    # configs[4], per-GPU share at 8 GPUs
}
METRIC = "24p loss+SimOTA images/s @640, 20 GT/img"
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes_per_image(A, C, Lmax):
    # head output read once + labels + fg_mask + matched_gt + pred_iou writes
    return A * C * 4 + Lmax * 51 * 4 + A * 1 + A * 4 * 2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(workload, rank, n_sets, device):
    from p24 import synth
    B, size, G, Lmax, kind, seed = WORKLOADS[workload]
    sets = []
    for i in range(n_sets):
        s = seed + 1000 * rank + 100 * i
        out = synth.make_head_outputs(B, size, 80, seed=s)
        lab = synth.make_labels(B, G, Lmax, size, 80, seed=s, kind=kind)
        sets.append((out, lab))
    xs, ys, ss = synth.make_grids(size)
    return sets, (xs, ys, ss)


def cpu_reference_run(workload, steps, warmup, sample_images=None):
    """The reference's CPU implementation of the path (oracle restatement: torch CPU eager, same ATen ops in the
    same order as the reference, pinned bit-for-bit to it in the build container) on all host threads."""
    from oracle import p24_oracle as orc
    from p24 import synth
    B, size, G, Lmax, kind, seed = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    nb = sample_images or B
    out = synth.make_head_outputs(nb, size, 80, seed=seed)
    lab = synth.make_labels(nb, G, Lmax, size, 80, seed=seed, kind=kind)
    xs, ys, ss = synth.make_grids(size)
    o = orc.LossOracle(80)
    for _ in range(warmup):
        o.forward((xs, ys, ss, out.clone(), []), lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.forward((xs, ys, ss, out.clone(), []), lab)
    dt = time.perf_counter() - t0
    return nb * steps / dt, dt / steps * 1e3, cores, f"{steps} x Loss_Function.forward on a batch of {nb} images ({size}x{size}, {G} GT/img)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused peer-memory all-reduce inside the last kernel (default) or NCCL + finalize kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, size, G, Lmax, kind, seed = WORKLOADS[args.workload]
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    config = {"workload": f"configs[1] training-shaped: batch {B}/GPU at {size}x{size} ({A} anchors), {G} GT/img, 80 classes"
              if args.workload == "train" else f"{args.workload}: batch {B}/GPU at {size}x{size} ({A} anchors), {G} GT/img",
              "per_gpu_batch": B, "global_batch": B * world, "anchors": A, "gt_per_image": G, "label_kind": kind,
              "sharding": "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 3)
        v, ms, cores, sample = cpu_reference_run(args.workload, steps, 1)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
                          "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the p24 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from p24 import lib as p24_lib
    from p24.losses import Loss_Function
    lib = p24_lib.load()

    img_bytes = A * 107 * 4
    n_sets = max(2, -(-int(2.2 * L2_BYTES) // (B * img_bytes)))  # rotate over > 2x L2 of distinct inputs
    # (P24_SEED_RANK: debug aid, the inputs another rank would get)
    sets, (xs, ys, ss) = make_inputs(args.workload, int(os.environ.get("P24_SEED_RANK", rank)), n_sets, dev)
    dsets = [(o.to(dev), l.to(dev)) for o, l in sets]
    gx, gy, gs = [t.to(dev) for t in xs], [t.to(dev) for t in ys], [t.to(dev) for t in ss]
    config["l2"] = f"inputs rotate over {n_sets} distinct batches ({n_sets * B * img_bytes / 2**20:.0f} MiB > 126 MiB L2)"
    from p24 import dist as p24_dist
    lf = Loss_Function(80)
    lf.reuse_buffers = not os.environ.get("P24_NO_REUSE")  # result tensors allocated once and overwritten per step (public option: less host work)
    if world > 1 and os.environ.get("P24_DEBUG_NO_EXCHANGE"):
        config["sharding"] = "DEBUG: ranks run independently (no all-reduce)"
    elif world > 1:
        p24_dist.attach(lf, peer=(args.allreduce == "peer"))
        fused = lf.peer_comm is not None
        config["sharding"] = (f"images sharded over {world} GPU(s); 28-float all-reduce per step: " +
                              ("fused into the last kernel over NVLink peer memory" if fused else "NCCL + finalize kernel"))

    def step(i):
        o, l = dsets[i % n_sets]
        return lf.forward_async((gx, gy, gs, o, []), l)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("P24_NO_CLOCK_SAMPLER"):
        sampler.start()
    t_load = time.perf_counter()
    for i in range(args.warmup):
        step(i)
    # keep the GPU under this load until nvidia-smi has had time to sample it (the timed region itself lasts only
    # milliseconds): extra untimed warm-up steps
    while True:
        for i in range(20):
            step(i)
        torch.cuda.synchronize()
        more = torch.tensor([1.0 if time.perf_counter() - t_load < 0.6 else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(more, op=dist.ReduceOp.MAX)  # every rank runs the same number of steps
        if float(more) == 0.0:
            break
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    for i in range(args.steps):
        res = step(i)
    t_host = (time.perf_counter() - t_host) / args.steps * 1e3  # host time to enqueue one step on this rank
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    host_ms = [t_host]
    if world > 1:
        th = torch.tensor([t_host], device=dev)
        tg = [torch.zeros_like(th) for _ in range(world)]
        dist.all_gather(tg, th)
        host_ms = [float(x) for x in tg]
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    value = B * world / (ms_step * 1e-3)

    # ---- per-kernel durations (second pass over the same steps, CUDA events recorded inside the C call on the
    # launching stream) -> roofline of the dominant kernel ------------------------------------------------------
    import ctypes
    lib.p24_profile_enable(1)
    acc = [0.0] * 6
    buf = (ctypes.c_float * 6)()
    for i in range(args.steps):
        step(i)
        p24_lib.check(lib.p24_profile_read(buf), "p24_profile_read")
        for k in range(6):
            acc[k] += buf[k]
    lib.p24_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    names = ["k_pass", "k_match", "k_resolve_loss"]
    kern_ms = [a / args.steps for a in acc[:len(names)]]
    rank_kernel_ms = None
    if world > 1:
        tk = torch.tensor(kern_ms + [ms_total / args.steps], device=dev)
        tg = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(tg, tk)
        rank_kernel_ms = [[round(float(v), 4) for v in t] for t in tg]  # per rank: k_pass, k_match, k_resolve_loss, step
    top = max(range(len(names)), key=lambda k: kern_ms[k])
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes_per_image(A, 107, Lmax) * B
    achieved = alg / (kern_ms[top] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get(names[top])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": names[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg,
                "kernel_ms": dict(zip(names, kern_ms)),
                "whole_step_frac": (alg / (ms_step * 1e-3) / 1e9) / peak}

    # ---- e2e: public API, host (pinned) inputs, H2D + D2H inside the timed region --------------------------------
    e2e = None
    if not args.no_e2e:
        hsets = [(o.pin_memory(), l.pin_memory()) for o, l in sets[:2]]
        d_out = torch.empty_like(dsets[0][0])
        d_lab = torch.empty_like(dsets[0][1])
        lf2 = Loss_Function(80)
        if world > 1:
            p24_dist.attach(lf2, peer=(args.allreduce == "peer"))

        def e2e_step(i):
            ho, hl = hsets[i % 2]
            d_out.copy_(ho, non_blocking=True)
            d_lab.copy_(hl, non_blocking=True)
            r = lf2.forward((gx, gy, gs, d_out, []), d_lab)
            return float(r[0])  # D2H read of the loss

        ke = max(3, min(args.steps, 20))
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * ke / float(tt[0]), "unit": "images/s",
               "h2d_bytes_per_step": (sets[0][0].numel() + sets[0][1].numel()) * 4, "d2h_bytes_per_step": 8,
               "steps": ke, "api": "Loss_Function.forward(outputs_train, labels)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores, sample = cpu_reference_run(args.workload, 2, 1)
        cpu_baseline = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        # k_gt_prep, k_pass, k_match, k_resolve_loss (+ k_finalize after an NCCL all-reduce)
        launches_per_step = 4 if (world == 1 or getattr(lf, "peer_comm", None) is not None or lf.process_group is None) else 5
        print(json.dumps({"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                          "roofline": roofline, "cpu_baseline": cpu_baseline,
                          "host_enqueue_ms_per_step": [round(x, 4) for x in host_ms],
                          "per_rank_kernel_and_step_ms": rank_kernel_ms,
                          "loss_check": float(res[0][0])}))
    if world > 1:
        barrier()
        for f in (lf, locals().get("lf2")):
            if f is not None and getattr(f, "peer_comm", None) is not None:
                f.peer_comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
