"""Join an ncu SASS source page with nvdisasm line info -> per-source-line instruction / stall-sample shares.
Usage: python tests/tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n] [mangled-name substring]   (build container)"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIB = os.path.join(ROOT, "exploration-of-potential_b200", "p24", "_lib", "libp24_b200.so")


def disasm_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    out = []
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, rec = None, None, False
        for ln in txt.splitlines():
            m = re.match(r"^\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                rec = kernel in cur_fn
                continue
            if not rec:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                out.append((cur_line, ln.strip()))
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    dis = sys.argv[4] if len(sys.argv) > 4 else kernel   # substring of the MANGLED name (rows6k_pass / rawlv6k_pass)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    inst = []
    for r in rows[hi + 1:]:
        if len(r) <= max(ci, cs):
            break
        try:
            inst.append((int(r[ci]), int(r[cs]), r[1]))
        except ValueError:
            break
    lines = disasm_lines(dis)
    print(f"ncu instructions: {len(inst)}  nvdisasm instructions: {len(lines)}")
    n = min(len(inst), len(lines))
    agg = {}
    for (cnt, smp, _), (loc, _) in zip(inst[:n], lines[:n]):
        a = agg.setdefault(loc, [0, 0])
        a[0] += cnt
        a[1] += smp
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    srcs = {}
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if loc:
            for base in (os.path.join(ROOT, "exploration-of-potential_b200", "csrc"),):
                p = os.path.join(base, loc[0])
                if os.path.exists(p):
                    if p not in srcs:
                        srcs[p] = open(p).read().splitlines()
                    if loc[1] - 1 < len(srcs[p]):
                        text = srcs[p][loc[1] - 1].strip()[:90]
        print(f"{a[0] / ti * 100:5.1f}% inst {a[1] / ts * 100:5.1f}% stall-samples  {loc}  {text}")


if __name__ == "__main__":
    main()
