"""Opcode histogram of the built kernels (cuobjdump -sass): shows the packed-FMA (FFMA2), bulk-copy (UBLKCP),
mbarrier (SYNCS), named-barrier (BAR) and reduction (RED) instructions the design relies on.
Usage: python tools/sass_opcodes.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tssplat_b200", "libtssplat_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::|tsb::", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and fn:
        hist[fn][m.group(1).split(".")[0]] += 1
KEY = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DADD", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "UBLKCP",
       "SYNCS", "BAR", "SHFL", "MEMBAR", "ERRBAR", "CCTL", "ACQBULK", "LDL", "STL", "BRA"]
print("# SASS opcode counts per kernel (static instructions, sm_100a), from cuobjdump -sass of", os.path.basename(lib))
for fn, h in hist.items():
    tot = sum(h.values())
    print(f"\n{fn}\n  total {tot}: " + ", ".join(f"{k} {h[k]}" for k in KEY if h.get(k)))
    rest = [(k, v) for k, v in h.most_common() if k not in KEY][:8]
    print("  other: " + ", ".join(f"{k} {v}" for k, v in rest))
print("\n# cuobjdump -res-usage")
for line in res.splitlines():
    if "Function" in line or "REG" in line:
        line = re.sub(r"_ZN\S*?(\d+)([a-z_]+kernel)", r"\2", line)
        print(" ", line.strip()[:200])
