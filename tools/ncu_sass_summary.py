"""Summarise an `ncu --page source --csv` dump: per-segment (split at BAR.SYNC) instruction and
stall-sample totals, plus the hottest SASS instructions.  Usage: ncu_sass_summary.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
col = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot_s = sum(int(r[col["# Samples"]]) for r in data)
tot_i = sum(int(r[col["Instructions Executed"]]) for r in data)
print(f"total samples {tot_s}, warp instructions {tot_i}")
seg, segs = dict(n=0, s=0, i=0, st={k: 0 for k in stalls}, first=0), []
for idx, r in enumerate(data):
    seg["n"] += 1
    seg["s"] += int(r[col["# Samples"]]); seg["i"] += int(r[col["Instructions Executed"]])
    for k in stalls:
        seg["st"][k] += int(r[col[k]] or 0)
    if "BAR.SYNC" in r[col["Source"]] or idx == len(data) - 1:
        seg["last"] = idx
        segs.append(seg)
        seg = dict(n=0, s=0, i=0, st={k: 0 for k in stalls}, first=idx + 1)
for s in segs:
    st = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
    print(f"seg sass[{s['first']:4d}..{s['last']:4d}] n={s['n']:4d} samples={s['s']:6d} ({100*s['s']/tot_s:5.1f}%) "
          f"warp-instr={s['i']:8d} ({100*s['i']/tot_i:5.1f}%)  " + ", ".join(f"{k[6:]}={v}" for k, v in st if v))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    i = data.index(r)
    st = sorted(((k, int(r[col[k]] or 0)) for k in stalls), key=lambda kv: -kv[1])[:3]
    print(f"  [{i:4d}] {r[col['# Samples']]:>6} {r[col['Instructions Executed']]:>8}  {r[col['Source']].strip():60s} "
          + ", ".join(f"{k[6:]}={v}" for k, v in st if v))
