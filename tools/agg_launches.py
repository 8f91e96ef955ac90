"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = {}
for r in data:
    if len(r) > vi:
        agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:92s} n={len(v):3d} total={sum(v)/1e3:8.3f} ms ({100*sum(v)/tot:5.1f}%)  avg={sum(v)/len(v):9.1f} us")
print(f"sum {tot/1e3:.3f} ms")
