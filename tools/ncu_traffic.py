"""DRAM traffic per launch from an `ncu --set full` capture -> profiles/<round>/traffic.json (read by bench.py's `roofline.traffic`).
    python tools/ncu_traffic.py <rep> <kernel-name regex> <json key> [<out json>]
Takes the LONGEST launch matching the regex (the 360x640 upsampling conv, not the 180x320 one) and stores
dram__bytes_read.sum + dram__bytes_write.sum in bytes."""
import csv
import json
import os
import re
import subprocess
import sys

rep, pat, key = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02", "traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U, data = rows[0], rows[1], rows[2:]
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
ki, ti = H.index("Kernel Name"), H.index("gpu__time_duration.sum")
ri, wi = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
tunit = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
best = None
for d in data:
    if re.search(pat, d[ki]):
        t = float(d[ti].replace(",", "")) * tunit.get(U[ti], 1.0)
        b = float(d[ri].replace(",", "")) * unit[U[ri]] + float(d[wi].replace(",", "")) * unit[U[wi]]
        if best is None or t > best[0]:
            best = (t, b, d[ki][:80])
assert best, "no kernel matched"
j = json.load(open(out)) if os.path.exists(out) else {}
j[key] = best[1]
j[key + "_source"] = f"{os.path.basename(rep)}: {best[2]} ({best[0] * 1e3:.3f} ms under ncu)"
json.dump(j, open(out, "w"), indent=1)
print(key, best)
