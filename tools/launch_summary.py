"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total and mean time.
   usage: python tools/launch_summary.py launches.csv [title]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("gpe::<unnamed>::", "").replace("gpe::", "")
    name = re.sub(r"^.*unnamed>::", "", name)
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for k, (n, t) in agg.items():
    print(f"{k[:60]:<60} launches={n:5d} total_us={t:12.1f} share={100*t/tot:5.1f}%  per_launch_us={t/n:10.1f}")
