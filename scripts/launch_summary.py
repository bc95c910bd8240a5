"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
    python scripts/launch_summary.py gpurun_out/launches.csv "<command line that produced it>" > profiles/rN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("dafk::", "")
    rows.append((name, us))
agg = defaultdict(lambda: [0, 0.0])
for n, us in rows:
    agg[n][0] += 1
    agg[n][1] += us
tot = sum(v[1] for v in agg.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print("# %d launches, %.1f ms summed kernel time" % (len(rows), tot / 1000.0))
print("%-70s %5s %12s %7s" % ("kernel", "n", "total_us", "share"))
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %5d %12.1f %6.1f%%" % (n[:70], c, us, 100.0 * us / tot))
