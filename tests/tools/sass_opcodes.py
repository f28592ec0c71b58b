"""profiles/sass_opcodes.txt: per-kernel SASS opcode counts of the built library (cuobjdump -sass), with the mnemonics
that prove the sm_100a features the design relies on.  Usage: python tests/tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIB = os.path.join(ROOT, "exploration-of-potential_b200", "p24", "_lib", "libp24_b200.so")
KEY = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "UCGABAR_ARV", "UCGABAR_WAIT", "ACQBULK", "PREEXIT", "BAR",
       "MATCH", "SHFL", "VOTE", "ATOM", "ATOMG", "ATOMS", "RED", "MUFU", "FFMA", "FMUL", "FADD", "DADD", "DMUL", "LDG",
       "STG", "LDS", "STS", "LDL", "STL", "CCTL", "MEMBAR", "ERRBAR", "NANOSLEEP"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = name.replace("<unnamed>::", "") or m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
    if m and cur:
        per[cur][m.group(1)] += 1
        if m.group(1) in ("SYNCS", "UBLKCP", "ATOM", "ATOMG", "RED", "BAR", "MEMBAR", "LDG", "LDGSTS"):
            per[cur][m.group(1) + m.group(2)] += 1
print(f"# SASS opcode counts per kernel: cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a)")
print("# UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, LDGSTS = cp.async, UCGABAR_* = cluster barrier,")
print("# PREEXIT / ACQBULK = programmatic dependent launch (trigger / grid dependency sync), MATCH = __match_any_sync,")
print("# LDL / STL = local memory (register spills), DADD / DMUL = fp64 (fixed-point sums, label packing)")
for k, c in per.items():
    tot = sum(v for o, v in c.items() if "." not in o)
    print(f"\n{k}: {tot} instructions")
    shown = [(o, c[o]) for o in KEY if c.get(o)]
    print("   " + "  ".join(f"{o}={v}" for o, v in shown))
    detail = sorted((o, v) for o, v in c.items() if "." in o)
    if detail:
        print("   detail: " + "  ".join(f"{o}={v}" for o, v in detail))
