#!/usr/bin/env python
"""bench.py — 24p loss + SimOTA images/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--only train|...]

A step = one pass of the hot path (fused SimOTA assignment + loss sums + re-weighting, i.e. the reference's
``Loss_Function.forward``) over one batch of synthetic head outputs.  Headline workload = BASELINE.json configs[1]:
batch 20 per GPU at 640x640 (8400 anchors), 20 GT/img, 80 classes, smooth labels.

  value         images/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks; the steps are
                pipelined (``Loss_Function.pipelined``: legal because the batches are resident before the loop starts;
                ``config.pipelined``, P24_NO_PIPELINE=1 switches it off)
  e2e           the same metric through the public ``Loss_Function.forward`` with HOST (pinned) inputs:
                H2D of the head output + labels (prefetched on a copy stream while the previous step computes) and the
                D2H read of the loss inside the timed region
  roofline      the dominant kernel's algorithmic bytes / its live CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU code (oracle/_ref, kind "reference"; else the oracle port) on a bounded sample

The rest of the north-star path is timed after the headline with fewer steps and reported as extra keys of the same
line, each with its own value / roofline / e2e / cpu_baseline:
  "train_spiky"  configs[1] with the spiky (i.i.d. radius) labels            SURVEY.md 8(d)
  "crowded"      configs[2]: batch 20, 100 GT/img
  "hires"        configs[4]: 1280x1280 (33600 anchors), GLOBAL batch 160 sharded over the GPUs (strong scaling)
  "postprocess"  configs[3]: batch 64 per GPU through utils.boxes.postprocess at the three settings of SURVEY.md 8(d)
  "train_raw"    configs[1] fed with the head's RAW conv outputs (decode fused into the kernels) vs torch decode + loss
  "post_raw"     configs[3] fed with the head's RAW conv outputs (sigmoid + decode fused) vs torch decode + postprocess

Multi-GPU (torchrun, one rank per GPU): images shard by rank (weak scaling, 20 images per GPU for the headline); the
only exchange is the SUM all-reduce of the 28 loss sums, fused into the kernel chain over NVLink peer memory (NCCL +
finalize kernel with ``--allreduce nccl``).  ``allreduce_check`` compares both on the first batch.
"""
from __future__ import annotations

import argparse
import ctypes
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

TRAIN_WORKLOADS = {
    # name: per-GPU batch (None: global batch / world), image size, GT per image, Lmax, label kind, seed, description
    "train": dict(B=20, size=640, G=20, Lmax=50, kind="smooth", seed=1, cfg="configs[1] training-shaped"),
    "train_spiky": dict(B=20, size=640, G=20, Lmax=50, kind="spiky", seed=1, cfg="configs[1], spiky labels"),
    "crowded": dict(B=20, size=640, G=100, Lmax=100, kind="smooth", seed=2, cfg="configs[2] crowded"),
    "hires": dict(B=None, global_B=160, size=1280, G=20, Lmax=50, kind="smooth", seed=4, cfg="configs[4] high-res"),
}
POST_SETTINGS = [
    ("conf0.25_nms0.45_batched", 0.25, 0.45, False),
    ("conf0.01_nms0.3_agnostic", 0.01, 0.3, True),     # show_24p.py:301
    ("conf0.01_nms0.65_batched", 0.01, 0.65, False),   # evaluator settings, ~8k candidates / image
]
METRIC = "24p loss+SimOTA images/s @640, 20 GT/img"
L2_BYTES = 126 * 1024 * 1024
TRAIN_STAGES = ["k_prep", "k_pass", "k_tail"]      # p24_profile_read slots 0..2
POST_STAGES = {4: "k_post_filter", 5: "k_post_nms"}  # slots 4, 5


def algorithmic_bytes_per_image(A, C, Lmax):
    # head output read once + labels + fg_mask + matched_gt + pred_iou writes
    return A * C * 4 + Lmax * 51 * 4 + A * 1 + A * 4 * 2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code when oracle/_ref is present, else the oracle port
# ----------------------------------------------------------------------------------------------------------------
def _cpu_impl():
    """(kind, make_loss_function, postprocess_per_image, shim)"""
    import contextlib
    from oracle import ref_runtime
    if ref_runtime.available():
        models, utils = ref_runtime.load()

        def post(p, conf, nms, agn):
            # the reference's batched call raises for B >= 2 (boxes.py:64-65): image by image
            return [utils.postprocess(p[i:i + 1], 80, conf, nms, agn)[0] for i in range(p.shape[0])]
        return "reference", (lambda: models.Loss_Function(80)), post, (lambda: ref_runtime.cuda0_shim(torch.device("cpu")))
    from oracle import p24_oracle as orc
    return "port", (lambda: orc.LossOracle(80)), (lambda p, conf, nms, agn: orc.postprocess(p, 80, conf, nms, agn)), \
        contextlib.nullcontext


def cpu_train_run(wl, steps, warmup, sample_images):
    from p24 import synth
    kind, make_lf, _, shim = _cpu_impl()
    cores = host_cores()
    torch.set_num_threads(cores)
    nb = sample_images
    out = synth.make_head_outputs(nb, wl["size"], 80, seed=wl["seed"])
    lab = synth.make_labels(nb, wl["G"], wl["Lmax"], wl["size"], 80, seed=wl["seed"], kind=wl["kind"])
    xs, ys, ss = synth.make_grids(wl["size"])
    lf = make_lf()
    with shim():
        for _ in range(warmup):
            lf.forward((xs, ys, ss, out.clone(), []), lab)
        t0 = time.perf_counter()
        for _ in range(steps):
            lf.forward((xs, ys, ss, out.clone(), []), lab)
        dt = time.perf_counter() - t0
    return {"value": nb * steps / dt, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"{steps} x Loss_Function.forward on a batch of {nb} images "
                      f"({wl['size']}x{wl['size']}, {wl['G']} GT/img, {wl['kind']} labels), torch CPU eager",
            "ms_per_step": dt / steps * 1e3}


def cpu_post_run(conf, nms, agn, nimg):
    from p24 import synth
    kind, _, post, _ = _cpu_impl()
    cores = host_cores()
    torch.set_num_threads(cores)
    p = synth.make_postprocess_input(nimg, 640, 80, seed=3)
    post(p[:2], conf, nms, agn)
    t0 = time.perf_counter()
    post(p, conf, nms, agn)
    dt = time.perf_counter() - t0
    return {"value": nimg / dt, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"utils.boxes.postprocess image by image over {nimg} images (640x640), torch CPU eager"}


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def barrier(ctx):
    if ctx.world > 1:
        ctx.dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ctx, x):
    t = torch.tensor([x], device=ctx.dev, dtype=torch.float64)
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t[0])


def gather_ranks(ctx, vals):
    t = torch.tensor(vals, device=ctx.dev, dtype=torch.float64)
    if ctx.world == 1:
        return [[float(v) for v in t]]
    g = [torch.zeros_like(t) for _ in range(ctx.world)]
    ctx.dist.all_gather(g, t)
    return [[float(v) for v in x] for x in g]


def make_train_inputs(ctx, wl, B, n_sets):
    from p24 import synth
    sets = []
    seed_rank = int(os.environ.get("P24_SEED_RANK", ctx.rank))  # debug aid: the inputs another rank would get
    for i in range(n_sets):
        s = wl["seed"] + 1000 * seed_rank + 100 * i
        if B * wl["size"] * wl["size"] > 40 * 640 * 640:
            # large batches (configs[4]): generated on the device, the host copy would take minutes
            out = synth.make_head_outputs(B, wl["size"], 80, seed=s, device=ctx.dev)
        else:
            out = synth.make_head_outputs(B, wl["size"], 80, seed=s)
        lab = synth.make_labels(B, wl["G"], wl["Lmax"], wl["size"], 80, seed=s, kind=wl["kind"])
        sets.append((out, lab))
    return sets, synth.make_grids(wl["size"])


def new_loss_function(ctx, allreduce):
    from p24 import dist as p24_dist
    from p24.losses import Loss_Function
    lf = Loss_Function(80)
    lf.reuse_buffers = not os.environ.get("P24_NO_REUSE")  # result tensors allocated once per shape (public option)
    if ctx.world > 1 and not os.environ.get("P24_DEBUG_NO_EXCHANGE"):
        p24_dist.attach(lf, peer=(allreduce == "peer"))
    ctx.lfs.append(lf)
    return lf


def bench_train(ctx, name, steps, warmup, allreduce, want_e2e=True, want_cpu=True, sampler=None, e2e_steps=None):
    """One training-path workload -> result dict (rank 0 gets the complete one)."""
    wl = TRAIN_WORKLOADS[name]
    B = wl["B"] if wl["B"] is not None else max(1, wl["global_B"] // ctx.world)
    size, G, Lmax = wl["size"], wl["G"], wl["Lmax"]
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    img_bytes = A * 107 * 4
    n_sets = max(2, -(-int(2.2 * L2_BYTES) // (B * img_bytes)))  # rotate over > 2x L2 of distinct inputs
    sets, (xs, ys, ss) = make_train_inputs(ctx, wl, B, n_sets)
    dsets = [(o.to(ctx.dev), l.to(ctx.dev)) for o, l in sets]
    gx, gy, gs = [t.to(ctx.dev) for t in xs], [t.to(ctx.dev) for t in ys], [t.to(ctx.dev) for t in ss]
    lf = new_loss_function(ctx, allreduce)
    # the batches are resident in HBM before the first step is enqueued: the steps may be pipelined (the preparation
    # kernel of a step beside the last kernel of the step before it, Loss_Function.pipelined / P24_F_EARLY_PREP).  The e2e
    # leg below copies its inputs inside the loop and does not set it.
    lf.pipelined = not os.environ.get("P24_NO_PIPELINE")
    fused = getattr(lf, "peer_comm", None) is not None
    config = {"workload": f"{wl['cfg']}: batch {B}/GPU at {size}x{size} ({A} anchors), {G} GT/img, 80 classes",
              "per_gpu_batch": B, "global_batch": B * ctx.world, "anchors": A, "gt_per_image": G,
              "label_kind": wl["kind"],
              "pipelined": bool(lf.pipelined),
              "l2": f"inputs rotate over {n_sets} distinct batches ({n_sets * B * img_bytes / 2**20:.0f} MiB > 126 MiB L2)",
              "sharding": "single GPU" if ctx.world == 1 else
              (f"images sharded over {ctx.world} GPU(s); 28-float all-reduce per step: " +
               ("fused into the kernel chain over NVLink peer memory" if fused else "NCCL + finalize kernel"))}

    def step(i):
        o, l = dsets[i % n_sets]
        return lf.forward_async((gx, gy, gs, o, []), l)

    check = None
    if ctx.world > 1 and name == "train":
        check = allreduce_check(ctx, dsets[0], (gx, gy, gs))
    t_load = time.perf_counter()
    for i in range(warmup):
        step(i)
    if sampler is not None:
        # keep the GPU under this load until nvidia-smi has had time to sample it (the timed region itself lasts only
        # milliseconds): extra untimed warm-up steps, the same number on every rank
        while True:
            for i in range(20):
                step(i)
            torch.cuda.synchronize()
            if max_over_ranks(ctx, 1.0 if time.perf_counter() - t_load < 0.6 else 0.0) == 0.0:
                break
    barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(ctx)
    e0.record()
    t_host = time.perf_counter()
    res = None
    for i in range(steps):
        res = step(i)
    lf.wait_results()   # results produced on a side stream (multi-GPU finalize) are ordered before the end mark
    t_host = (time.perf_counter() - t_host) / steps * 1e3  # host time to enqueue one step on this rank
    e1.record()
    barrier(ctx)
    ms_total = e0.elapsed_time(e1)
    ms_step = max_over_ranks(ctx, ms_total) / steps
    value = B * ctx.world / (ms_step * 1e-3)
    lf.check_errors()
    loss_check = float(res[0][0])
    # the same loop without the pipelining promise (what a training loop sees when the head writes `outputs` right before
    # the loss): reported beside the headline, never as the headline
    plain = None
    if lf.pipelined:
        lf.pipelined = False
        for i in range(3):
            step(i)
        barrier(ctx)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(steps):
            step(i)
        lf.wait_results()
        p1.record()
        barrier(ctx)
        ms_plain = max_over_ranks(ctx, p0.elapsed_time(p1)) / steps
        plain = {"ms_per_step": ms_plain, "value": B * ctx.world / (ms_plain * 1e-3)}
        lf.pipelined = True

    # ---- per-kernel durations (second pass over the same steps, CUDA events recorded inside the C call on the
    # launching stream, plain stream order) -> roofline of the dominant kernel -------------------------------------
    from p24 import lib as p24_lib
    lib = ctx.lib
    lib.p24_profile_enable(1)
    acc = [0.0] * 8
    buf = (ctypes.c_float * 8)()
    ksteps = max(3, min(steps, 30))
    for i in range(ksteps):
        step(i)
        p24_lib.check(lib.p24_profile_read(buf), "p24_profile_read")
        for k in range(8):
            acc[k] += buf[k]
    lf.wait_results()
    lib.p24_profile_enable(0)
    kern_ms = [a / ksteps for a in acc[:len(TRAIN_STAGES)]]
    status = lf.read_status()   # (one read: the counters restart with it)
    per_rank = gather_ranks(ctx, kern_ms + [ms_total / steps, t_host, status["exchange_wait_us"]])
    top = max(range(len(TRAIN_STAGES)), key=lambda k: kern_ms[k])
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes_per_image(A, 107, Lmax) * B
    achieved = alg / (kern_ms[top] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(name, {}).get(TRAIN_STAGES[top])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": TRAIN_STAGES[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "kernel_ms": dict(zip(TRAIN_STAGES, kern_ms)),
                "whole_step_frac": (alg / (ms_step * 1e-3) / 1e9) / peak}
    stats = {k: v for k, v in status.items() if k != "exchange_wait_us"}

    # ---- e2e: public API, host (pinned) inputs, H2D + D2H inside the timed region ---------------------------------
    e2e = None
    if want_e2e:
        hsets = [(o.cpu().pin_memory(), l.cpu().pin_memory()) for o, l in sets[:2]]
        # two device buffers and a copy stream: the H2D copy of step i + 1 is in flight while step i computes (a data
        # loader's prefetch); every step's copy, compute and D2H read lie inside the timed region
        d_bufs = [(torch.empty_like(dsets[0][0]), torch.empty_like(dsets[0][1])) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=ctx.dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        lf2 = new_loss_function(ctx, allreduce)

        def enqueue_copy(i):
            ho, hl = hsets[i % 2]
            with torch.cuda.stream(copy_stream):
                d_bufs[i % 2][0].copy_(ho, non_blocking=True)
                d_bufs[i % 2][1].copy_(hl, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def e2e_step(i, last):
            if not last:
                enqueue_copy(i + 1)   # (its buffer was last read by step i - 1, whose result has been read back)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            r = lf2.forward((gx, gy, gs, d_bufs[i % 2][0], []), d_bufs[i % 2][1])
            return float(r[0])  # D2H read of the loss

        ke = e2e_steps or max(3, min(steps, 20))
        enqueue_copy(0)
        for i in range(3):
            e2e_step(i, i == 2)
        torch.cuda.synchronize()
        barrier(ctx)
        t0 = time.perf_counter()
        enqueue_copy(3)               # (the first timed step's copy is inside the timed region like all the others)
        for i in range(3, 3 + ke):
            e2e_step(i, i == 2 + ke)
        torch.cuda.synchronize()
        dt = max_over_ranks(ctx, time.perf_counter() - t0)
        e2e = {"value": B * ctx.world * ke / dt, "unit": "images/s",
               "h2d_bytes_per_step": (sets[0][0].numel() + sets[0][1].numel()) * 4, "d2h_bytes_per_step": 8,
               "steps": ke, "api": "Loss_Function.forward(outputs_train, labels)",
               "pipeline": "H2D of step i+1 on a copy stream while step i computes (2 device buffers)"}
        del hsets, d_bufs
    cpu = None
    if want_cpu and ctx.rank == 0 and ctx.world == 1:
        sample = {"train": 20, "train_spiky": 20, "crowded": 2, "hires": 2}[name]
        cpu = cpu_train_run(wl, 2 if sample >= 20 else 1, 1, sample)
    del dsets, sets
    torch.cuda.empty_cache()
    launches = len(TRAIN_STAGES) + (1 if ctx.world > 1 else 0)
    return {"value": value, "unit": "images/s", "ms_per_step": ms_step, "steps": steps, "config": config,
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": launches * steps,
            "host_enqueue_ms_per_step": round(t_host, 4),
            "per_rank_ms": {"columns": TRAIN_STAGES + ["step", "host_enqueue", "exchange_wait_us"],
                            "rows": [[round(v, 4) for v in r] for r in per_rank]},
            "slow_path": stats, "loss_check": loss_check, "allreduce_check": check, "unpipelined": plain}


def bench_head_fusion(ctx, steps, warmup):
    """the loss on the head's RAW per-level conv outputs (decode fused into the kernels) against the
    reference's way (torch cat / permute / decode passes, then the loss on the decoded buffer).  configs[1] shapes."""
    from p24 import head as p24_head
    from p24 import synth
    wl = TRAIN_WORKLOADS["train"]
    B, size, G, Lmax = wl["B"], wl["size"], wl["G"], wl["Lmax"]
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    n_sets = max(2, -(-int(2.2 * L2_BYTES) // (B * A * 107 * 4)))
    raws, labs = [], []
    for i in range(n_sets):
        r, o, c = synth.make_raw_levels(B, size, 80, seed=wl["seed"] + 100 * i, device=ctx.dev)
        raws.append((r, o, c))
        labs.append(synth.make_labels(B, G, Lmax, size, 80, seed=wl["seed"] + 100 * i, kind=wl["kind"]).to(ctx.dev))
    out = {}
    for mode in ("fused", "unfused"):
        lf = new_loss_function(ctx, "peer")
        # (fused: the raw batches are resident, the steps may be pipelined like the headline loop; unfused: torch kernels
        # write the decoded buffer between the steps, so they may not)
        lf.pipelined = mode == "fused" and not os.environ.get("P24_NO_PIPELINE")

        def step(i):
            r, o, c = raws[i % n_sets]
            tup = p24_head.train_outputs(r, o, c, synth.STRIDES, fused=(mode == "fused"))
            return lf.forward_async(tup, labs[i % n_sets])

        for i in range(warmup + 5):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            res = step(i)
        e1.record()
        torch.cuda.synchronize()
        lf.check_errors()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"ms_per_step": ms, "value": B / (ms * 1e-3), "loss_check": float(res[0][0])}
    alg = algorithmic_bytes_per_image(A, 107, Lmax) * B
    peak, _ = measured_peak()
    return {"metric": "train_step_images_per_sec_from_raw_head_outputs", "unit": "images/s",
            "config": {"workload": f"configs[1] shapes, inputs = raw per-level conv outputs reg/obj/cls [B, C, H, W] "
                                   f"(batch {B}, {A} anchors, {G} GT/img)",
                       "l2": f"inputs rotate over {n_sets} distinct batches"},
            "value": out["fused"]["value"], "ms_per_step": out["fused"]["ms_per_step"],
            "unfused_torch_decode_then_loss": out["unfused"],
            "speedup_vs_unfused": out["unfused"]["ms_per_step"] / out["fused"]["ms_per_step"],
            "loss_check_equal": out["fused"]["loss_check"] == out["unfused"]["loss_check"],
            "whole_step_frac_of_hbm_peak": (alg / (out["fused"]["ms_per_step"] * 1e-3) / 1e9) / peak,
            "algorithmic_bytes_per_step": alg}


def bench_post_fusion(ctx, steps, warmup):
    """SURVEY.md 8f row 2, inference side: postprocess on the head's RAW conv outputs (sigmoid + decode inside the
    filter pass) against the reference's way (torch sigmoid / cat / permute / decode passes, then the postprocess on the
    decoded prediction).  configs[3] shapes (batch 64 at 640x640), conf 0.25 / nms 0.45 batched."""
    from p24 import boxes as p24_boxes
    from p24 import head as p24_head
    from p24 import synth
    B, size = 64, 640
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    raws = []
    for i in range(2):  # 2 x 230 MB > 126 MiB L2
        r, o, c = synth.make_raw_levels(B, size, 80, seed=3 + 100 * i, device=ctx.dev)
        for t in o + c:
            t += 3.6   # (scores reach the 0.25 threshold for a few hundred anchors per image)
        raws.append((r, o, c))
    out = {}
    for mode in ("fused", "unfused"):
        def step(i):
            r, o, c = raws[i % 2]
            return p24_boxes.postprocess_raw(p24_head.infer_outputs(r, o, c, synth.STRIDES, fused=(mode == "fused")),
                                             80, 0.25, 0.45, False)
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            res = step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"ms_per_step": ms, "value": B / (ms * 1e-3), "kept_check": int(res[1].sum())}
        if mode == "fused":
            from p24 import lib as p24_lib
            ctx.lib.p24_profile_enable(1)
            buf = (ctypes.c_float * 8)()
            acc = {k: 0.0 for k in POST_STAGES}
            for i in range(5):
                step(i)
                p24_lib.check(ctx.lib.p24_profile_read(buf), "p24_profile_read")
                for k in POST_STAGES:
                    acc[k] += buf[k]
            ctx.lib.p24_profile_enable(0)
            out[mode]["kernel_ms"] = {POST_STAGES[k].replace("filter", "filter_raw"): v / 5 for k, v in acc.items()}
    alg = A * 107 * 4 * B
    peak, _ = measured_peak()
    return {"metric": "postprocess_images_per_sec_from_raw_head_outputs", "unit": "images/s",
            "config": {"workload": f"configs[3] shapes, inputs = raw per-level conv outputs (batch {B}, {A} anchors), "
                                   "conf 0.25 / nms 0.45 batched", "l2": "inputs rotate over 2 distinct batches (460 MB)"},
            "value": out["fused"]["value"], "ms_per_step": out["fused"]["ms_per_step"],
            "kernel_ms": out["fused"]["kernel_ms"],
            "unfused_torch_decode_then_postprocess": out["unfused"],
            "speedup_vs_unfused": out["unfused"]["ms_per_step"] / out["fused"]["ms_per_step"],
            "kept_equal": out["fused"]["kept_check"] == out["unfused"]["kept_check"],
            "whole_step_frac_of_hbm_peak": (alg / (out["fused"]["ms_per_step"] * 1e-3) / 1e9) / peak,
            "algorithmic_bytes_per_step": alg}


def allreduce_check(ctx, dset, grids):
    """N > 1: the fused peer-memory exchange against NCCL + finalize on the same shard (first batch): the loss must be
    bit-identical on all ranks in both modes, and the two modes must agree to fp32 rounding of the 28-float sums."""
    from p24 import dist as p24_dist
    from p24.losses import Loss_Function
    gx, gy, gs = grids
    out = {}
    for mode in ("peer", "nccl"):
        lf = Loss_Function(80)
        p24_dist.attach(lf, peer=(mode == "peer"))
        ctx.lfs.append(lf)
        r, _, _ = lf.forward_async((gx, gy, gs, dset[0], []), dset[1])
        lf.wait_results()
        torch.cuda.synchronize()
        lf.check_errors()
        mine = r[:28].detach().clone()
        g = [torch.zeros_like(mine) for _ in range(ctx.world)]
        ctx.dist.all_gather(g, mine)
        out[mode] = (mine, all(torch.equal(g[0], x) for x in g), getattr(lf, "peer_comm", None) is not None)
    a, b = out["peer"][0].double(), out["nccl"][0].double()
    rel = float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())
    return {"bit_identical_across_ranks": bool(out["peer"][1] and out["nccl"][1]), "fused_vs_nccl_rel": rel,
            "fused_path_active": bool(out["peer"][2])}


def bench_post(ctx, label, conf, nms, agn, steps, warmup, want_e2e=True, want_cpu=True):
    from p24 import boxes as p24_boxes
    from p24 import lib as p24_lib
    from p24 import synth
    B, size = 64, 640
    A = sum((size // s) ** 2 for s in (8, 16, 32))
    n_sets = 2  # 2 x 230 MB > 126 MiB L2
    hsets = [synth.make_postprocess_input(B, size, 80, seed=3 + 1000 * ctx.rank + 100 * i) for i in range(n_sets)]
    dsets = [p.to(ctx.dev) for p in hsets]

    def step(i):
        return p24_boxes.postprocess_raw(dsets[i % n_sets], 80, conf, nms, agn)

    for i in range(warmup):
        step(i)
    barrier(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        r = step(i)
    e1.record()
    barrier(ctx)
    ms_step = max_over_ranks(ctx, e0.elapsed_time(e1)) / steps
    lib = ctx.lib
    lib.p24_profile_enable(1)
    buf = (ctypes.c_float * 8)()
    acc = {k: 0.0 for k in POST_STAGES}
    ks = max(3, min(steps, 10))
    for i in range(ks):
        step(i)
        p24_lib.check(lib.p24_profile_read(buf), "p24_profile_read")
        for k in POST_STAGES:
            acc[k] += buf[k]
    lib.p24_profile_enable(0)
    kern_ms = {POST_STAGES[k]: v / ks for k, v in acc.items()}
    peak, peak_src = measured_peak()
    alg = A * 107 * 4 * B  # the prediction read once
    top = max(kern_ms, key=kern_ms.get)
    roofline = {"bound": "hbm", "kernel": top, "achieved": alg / (kern_ms[top] * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": alg / (kern_ms[top] * 1e-3) / 1e9 / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": kern_ms,
                "filter_frac": alg / (kern_ms["k_post_filter"] * 1e-3) / 1e9 / peak,
                "whole_step_frac": alg / (ms_step * 1e-3) / 1e9 / peak}
    cand, cnt = r[0], r[1]
    e2e = None
    if want_e2e:
        pinned = [p.pin_memory() for p in hsets]
        d_in = torch.empty_like(dsets[0])

        def e2e_step(i):
            d_in.copy_(pinned[i % n_sets], non_blocking=True)
            return p24_boxes.postprocess(d_in, 80, conf, nms, agn)  # reads the per-image counts back (D2H)

        for i in range(2):
            e2e_step(i)
        barrier(ctx)
        ke = max(3, min(steps, 10))
        t0 = time.perf_counter()
        for i in range(ke):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = max_over_ranks(ctx, time.perf_counter() - t0)
        e2e = {"value": B * ctx.world * ke / dt, "unit": "images/s", "h2d_bytes_per_step": hsets[0].numel() * 4,
               "d2h_bytes_per_step": 4 * B, "steps": ke, "api": "p24.boxes.postprocess(prediction, 80, conf, nms, agnostic)"}
    cpu = None
    if want_cpu and ctx.rank == 0 and ctx.world == 1:
        cpu = cpu_post_run(conf, nms, agn, 64 if conf > 0.1 else 16)
    out = {"settings": {"conf_thre": conf, "nms_thre": nms, "class_agnostic": agn},
           "value": B * ctx.world / (ms_step * 1e-3), "unit": "images/s", "ms_per_step": ms_step, "steps": steps,
           "config": {"workload": f"configs[3] inference postprocess: batch {B}/GPU at {size}x{size} ({A} anchors)",
                      "per_gpu_batch": B, "mean_candidates_per_image": float(cand.float().mean()),
                      "mean_kept_per_image": float(cnt.float().mean())},
           "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": 2 * steps}
    del dsets, hsets
    torch.cuda.empty_cache()
    return out


def reference_arm(args, rank):
    if rank != 0:
        return
    wl = TRAIN_WORKLOADS["train"]
    A = sum((wl["size"] // s) ** 2 for s in (8, 16, 32))
    steps = min(args.steps, 3)
    cpu = cpu_train_run(wl, steps, 1, wl["B"])
    config = {"workload": f"{wl['cfg']}: batch {wl['B']}/GPU at {wl['size']}x{wl['size']} ({A} anchors), "
                          f"{wl['G']} GT/img, 80 classes",
              "per_gpu_batch": wl["B"], "global_batch": wl["B"], "anchors": A, "gt_per_image": wl["G"],
              "label_kind": wl["kind"]}
    v = cpu["value"]
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
                      "steps": steps, "warmup": 1, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                      "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
                      "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--only", default=None, help="comma list of blocks to run beside nothing else: "
                    "train,train_spiky,crowded,hires,train_raw,postprocess,post_raw (default: headline + all extras)")
    ap.add_argument("--workload", default=None, help="(compat) same as --only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused peer-memory all-reduce inside the kernel chain (default) or NCCL + finalize kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the p24 path has no CPU fallback")
    ctx = Ctx()
    ctx.rank, ctx.world = rank, world
    torch.cuda.set_device(local_rank)
    ctx.dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    ctx.dist = dist
    if world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    from p24 import lib as p24_lib
    ctx.lib = p24_lib.load()
    ctx.lfs = []
    only = args.only or args.workload
    blocks = only.split(",") if only else (["train"] if args.no_extras else
                                           ["train", "train_spiky", "crowded", "hires", "train_raw", "postprocess", "post_raw"])
    want_cpu, want_e2e = not args.no_cpu_baseline, not args.no_e2e
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("P24_NO_CLOCK_SAMPLER"):
        sampler.start()
    results = {}
    head_name = blocks[0] if blocks[0] in TRAIN_WORKLOADS else None
    clocks = None
    for bi, bname in enumerate(blocks):
        if bi == 1 and head_name is None:
            # (no training block first: the clocks were sampled under the first block; nvidia-smi polling stops here, it
            # can stall launches for milliseconds and the remaining blocks are short)
            clocks = sampler.stop() if rank == 0 else None
        if bname in TRAIN_WORKLOADS:
            head = bname == head_name
            k = args.steps if head else max(5, args.steps // (4 if bname == "hires" else 2))
            results[bname] = bench_train(ctx, bname, k, args.warmup, args.allreduce, want_e2e, want_cpu,
                                         sampler=sampler if head else None,
                                         e2e_steps=None if bname != "hires" else 3)
            if head:
                clocks = sampler.stop() if rank == 0 else None
        elif bname == "train_raw":
            if world == 1:   # (a single-GPU comparison: the sharded path is the same chain)
                results["train_raw"] = bench_head_fusion(ctx, max(5, args.steps // 2), args.warmup)
        elif bname == "post_raw":
            if world == 1:
                results["post_raw"] = bench_post_fusion(ctx, max(5, args.steps // 2), args.warmup)
        elif bname == "postprocess":
            k = max(5, args.steps // 2)
            results["postprocess"] = {lab: bench_post(ctx, lab, c, n, a, k, args.warmup, want_e2e, want_cpu)
                                      for lab, c, n, a in POST_SETTINGS}
        else:
            raise SystemExit(f"unknown block {bname}")
    if head_name is None and clocks is None:
        clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        if head_name is not None:
            h = results.pop(head_name)
            line = {"metric": METRIC, "value": h["value"], "unit": "images/s", "n_gpus": world, "steps": h["steps"],
                    "warmup": args.warmup, "ms_per_step": h["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": h["config"], "clocks": clocks, "e2e": h["e2e"], "gpu_launches": h["gpu_launches"],
                    "roofline": h["roofline"], "cpu_baseline": h["cpu_baseline"],
                    "host_enqueue_ms_per_step": h["host_enqueue_ms_per_step"], "per_rank_ms": h["per_rank_ms"],
                    "slow_path": h["slow_path"], "loss_check": h["loss_check"], "unpipelined": h["unpipelined"],
                    "allreduce_check": h["allreduce_check"]}
        else:
            line = {"metric": METRIC, "n_gpus": world, "warmup": args.warmup, "clocks": clocks}
        for k, v in results.items():
            if k == "hires":
                v["scaling"] = "strong (global batch 160 sharded over the GPUs)"
            line[k] = v
        print(json.dumps(line))
    if world > 1:
        barrier(ctx)
        for f in ctx.lfs:
            if getattr(f, "peer_comm", None) is not None:
                f.peer_comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
