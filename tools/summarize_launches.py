"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last step only, grouped by kernel."""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns, r.get("Grid Size", ""), r.get("Block Size", "")))
if not rows:
    print("no rows")
    sys.exit(0)
# last step = the launches after the last multi_cast_kernel (one per forward)
starts = [i for i, r in enumerate(rows) if "multi_cast" in r[1]]
step = rows[starts[-1]:] if starts else rows
tot = sum(r[2] for r in step)
agg = OrderedDict()
for _, name, ns, grid, blk in step:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"vlb::\(anonymous namespace\)::", "", short)
    k = (short, grid, blk)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
print("last step: %d launches, %.3f ms total (serialised, cold-cache ncu timing)" % (len(step), tot / 1e6))
print("%-90s %6s %10s %7s %9s" % ("kernel (grid, block)", "count", "total us", "share", "avg us"))
for (short, grid, blk), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-90s %6d %10.1f %6.1f%% %9.2f" % ((short + " " + grid + " " + blk)[:90], n, ns / 1e3, 100 * ns / tot, ns / 1e3 / n))
