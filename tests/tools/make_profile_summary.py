"""Turn gpurun_out/launches_<tag>.csv (ncu launch list) and gpurun_out/prof_<tag>.ncu-rep (ncu --set full) into the
tracked summaries under profiles/ : launch shares per kernel, key metrics per kernel, and profiles/dram_traffic.json
(the `roofline.traffic` figure bench.py reports).  Usage: python tests/tools/make_profile_summary.py r01"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = []

lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5 and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC,
    #          Section Name, Metric Name, Metric Unit, Metric Value
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("<unnamed>::", "").replace("rows::", "").replace("rawlv::", "raw_")
        unit, val = r[-2], float(r[-1].replace(",", ""))
        us = val / 1e3 if unit in ("nsecond", "ns") else val
        per.setdefault(name, []).append(us)
    ours = {k: v for k, v in per.items() if k.startswith("k_")}
    tot = sum(sum(v) for v in ours.values())
    lines.append(f"## Launch list ({os.path.basename(lp)}: `ncu --metrics gpu__time_duration.sum --clock-control none` "
                 f"on `bench.py --steps 3 --warmup 3`; cold-cache, serialised — compare SHARES)\n")
    lines.append("| kernel | launches | mean us | share of our kernels' time |\n|---|---|---|---|")
    for k, v in ours.items():
        lines.append(f"| {k} | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f} % |")
    others = {k: v for k, v in per.items() if not k.startswith("k_")}
    lines.append(f"\nOther launches in the process (torch fill / copy / RNG kernels outside the timed step): "
                 f"{sum(len(v) for v in others.values())}\n")

rp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
traffic = {}
if os.path.exists(rp):
    txt = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "duration"), ("smsp__inst_executed.sum", "warp inst"),
            ("sm__inst_executed.avg.per_cycle_elapsed", "IPC / SM"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
            ("lts__t_bytes.sum", "L2 bytes")]
    lines.append(f"## Per-kernel metrics ({os.path.basename(rp)}: `ncu --set full --clock-control none`, one step of "
                 f"batch 20 @640x640, 20 GT/img)\n")
    lines.append("| kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 1))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("rows::", "").replace("rawlv::", "raw_")
        cells = []
        for key, _ in want:
            v = r[idx[key]] if key in idx else ""
            u = units[idx[key]] if key in idx else ""
            try:
                cells.append(f"{float(v):,.2f} {u}".strip())
            except ValueError:
                cells.append(v)
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
        try:
            mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            rd = float(r[idx["dram__bytes_read.sum"]]) * mult.get(units[idx["dram__bytes_read.sum"]], 1.0)
            wr = float(r[idx["dram__bytes_write.sum"]]) * mult.get(units[idx["dram__bytes_write.sum"]], 1.0)
            traffic[name] = rd + wr
        except Exception:
            pass
    lines.append("")

with open(os.path.join(out_dir, f"{tag}_summary.md"), "w") as fh:
    fh.write(f"# ncu summary {tag}\n\n" + "\n".join(lines) + "\n")
if traffic:
    tp = os.path.join(out_dir, "dram_traffic.json")
    data = json.load(open(tp)) if os.path.exists(tp) else {}
    data["train"] = traffic
    data["_source"] = f"profiles/{tag}_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(data, open(tp, "w"), indent=1)
print("\n".join(lines))
