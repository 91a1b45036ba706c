"""Writes profiles/traffic.json from an `ncu --set full` report of the fused kernel (64-sphere pack), stamped
with the hash of the kernel / plan sources so that bench.py refuses it once they change.

    python tools/make_traffic.py gpurun_out/prof_r02_final_S64.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_hash  # noqa: E402

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"dram_bytes_per_step": rd + wr, "dram_read": rd, "dram_write": wr, "kernel_hash": kernel_hash(),
       "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "energy_grad_kernel",
       "duration_us_under_ncu": get("gpu__time_duration.sum") / 1e3 if units[hdr.index("gpu__time_duration.sum")] == "ns" else float(vals[hdr.index("gpu__time_duration.sum")]),
       "source": f"ncu --set full --clock-control none, one launch of the 64 x 4096 pack ({os.path.basename(rep)}); cold caches, so it is an upper bound for the L2-warm steps"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(out)
