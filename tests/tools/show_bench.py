"""Print the essentials of a bench.py JSON line.  Usage: python tests/tools/show_bench.py <file>"""
import json
import sys

line = [x for x in open(sys.argv[1]).read().split("\n") if x.startswith("{")][-1]
d = json.loads(line)
if "value" in d:
    print(f"headline: {d['value']:.0f} {d['unit']}  {d['ms_per_step'] * 1e3:.1f} us/step  n_gpus={d['n_gpus']}  steps={d['steps']}")
    print("  e2e:", d.get("e2e"))
    r = d["roofline"]
    print(f"  roofline: {r['kernel']} {r['achieved']:.0f}/{r['peak']:.0f} GB/s frac {r['frac']:.3f} traffic {r['traffic']} "
          f"kernel_ms {r['kernel_ms']} whole_step_frac {r['whole_step_frac']:.3f}")
    print("  cpu_baseline:", d.get("cpu_baseline"))
    print("  clocks:", d.get("clocks"), " launches:", d.get("gpu_launches"), " host_enqueue_ms:", d.get("host_enqueue_ms_per_step"))
    print("  slow_path:", d.get("slow_path"), " allreduce_check:", d.get("allreduce_check"))
    print("  per_rank_ms:", d.get("per_rank_ms"))
    print("  unpipelined:", d.get("unpipelined"))
for k in ("train_spiky", "crowded", "hires"):
    if k in d:
        v = d[k]
        print(f"{k}: {v['value']:.0f} img/s {v['ms_per_step'] * 1e3:.1f} us  kernels {v['roofline']['kernel_ms']}  slow {v['slow_path']}  "
              f"e2e {v['e2e']['value'] if v.get('e2e') else None}")
for k in ("train_raw", "post_raw"):
    if k in d:
        v = d[k]
        un = v.get("unfused_torch_decode_then_loss") or v.get("unfused_torch_decode_then_postprocess")
        print(f"{k}: {v['value']:.0f} img/s {v['ms_per_step'] * 1e3:.1f} us  unfused {un['ms_per_step'] * 1e3:.1f} us  x{v['speedup_vs_unfused']:.2f}  "
              f"{v.get('kernel_ms', '')}")
for k, v in d.get("postprocess", {}).items():
    print(f"post {k}: {v['value']:.0f} img/s {v['ms_per_step'] * 1e3:.1f} us  kernels {v['roofline']['kernel_ms']}  "
          f"filter_frac {v['roofline']['filter_frac']:.2f}  e2e {v['e2e']['value'] if v.get('e2e') else None}  cpu {v['cpu_baseline']['value'] if v.get('cpu_baseline') else None}")
