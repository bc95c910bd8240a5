"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
    python scripts/launch_summary.py gpurun_out/launches.csv "<command line that produced it>" > profiles/rN_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
dram = defaultdict(float)
if lines and "gpu__time_duration.sum" in lines[0]:
    # `--page raw` list: one row per launch, one column per metric, units in the second row
    table = list(csv.reader(l for l in lines if l.startswith('"')))
    header, units = table[0], dict(zip(table[0], table[1]))
    tscale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}[units["gpu__time_duration.sum"]]
    bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in table[2:]:
        d = dict(zip(header, r))
        name = re.sub(r"\(.*$", "", d["Kernel Name"]).replace("void ", "").replace("dafk::", "")
        rows.append((name, float(d["gpu__time_duration.sum"].replace(",", "")) * tscale))
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if d.get(k):
                dram[name] += float(d[k].replace(",", "")) * bscale[units[k]]
    lines = []
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
print("%-70s %5s %12s %7s %14s" % ("kernel", "n", "total_us", "share", "dram MB/launch"))
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %5d %12.1f %6.1f%% %14s" % (n[:70], c, us, 100.0 * us / tot, ("%.2f" % (dram[n] / c / 1e6)) if n in dram else "-"))
