"""Per-kernel SASS evidence of Blackwell-native code (B200_PROFILING.md "What proves a Blackwell-native kernel"):
    python tools/sass_summary.py > profiles/r02/sass_summary.txt
counts, per kernel of fast-srgan_b200/libfsr_b200.so, the tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA loads / stores
(UTMALDG / UTMASTG), tcgen05.commit (UTCBAR) and legacy warp-level MMAs (HMMA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fast-srgan_b200", "libfsr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "2CTA MMA": r"UTC[A-Z]*MMA\.2CTA", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
        "UTCBAR": r"\bUTCBAR", "HMMA": r"\bHMMA", "RED/ATOM": r"\b(RED|ATOMG|ATOM)\b"}
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:110]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels; SASS instruction counts")
print(f"{'kernel':112s} " + " ".join(f"{k:>9s}" for k in pats))
tot = collections.Counter()
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTC*MMA"]):
    if sum(c.values()):
        print(f"{k:112s} " + " ".join(f"{c[p]:9d}" for p in pats))
    tot.update(c)
print(f"{'TOTAL':112s} " + " ".join(f"{tot[p]:9d}" for p in pats))
