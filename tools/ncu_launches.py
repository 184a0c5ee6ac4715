"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_launches.py launches.csv > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"void arl::tc::tc_kernel<arl::(?:<unnamed>::)?(.*)>\(.*", r"tc_kernel<\1>", name)
    name = re.sub(r"\(.*", "", name)[:78]
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("%-78s %5s %12s %9s %6s" % ("kernel", "n", "total_us", "avg_us", "share"))
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-78s %5d %12.1f %9.1f %5.1f%%" % (name, n, us, us / n, 100 * us / tot))
print("# total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
