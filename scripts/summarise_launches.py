"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, share, count and mean per kernel.
    python scripts/summarise_launches.py gpurun_out/r02_launches_128.csv "<header text>" > profiles/r02_launches_128_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    started = False
    for r in csv.reader(f):
        if not started:
            started = bool(r) and r[0] == "ID"
            if started:
                hdr = r
            continue
        rows.append(r)
k_name, k_val, k_unit, k_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
tot = defaultdict(float)
cnt = defaultdict(int)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows:
    if len(r) <= k_val or r[k_metric] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[k_name]).strip()
    tot[name] += float(r[k_val].replace(",", "")) * scale.get(r[k_unit], 1e-6)
    cnt[name] += 1
total = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"total {total:.1f} ms over {sum(cnt.values())} launches")
for name in sorted(tot, key=tot.get, reverse=True)[:40]:
    print(f"{tot[name]:10.1f} ms {100 * tot[name] / total:5.1f}%  n={cnt[name]:6d}  avg {tot[name] / cnt[name]:10.3f} ms  {name}")
