"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares.
    python scripts/launch_shares.py gpurun_out/launches.csv [skip_regex] > profiles/x.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's kernels[].share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) < len(hdr) or row[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row[ci["Kernel Name"]])
    if skip and skip.search(name):
        continue
    v = float(row[ci["Metric Value"]].replace(",", ""))
    unit = row[ci["Metric Unit"]]
    us = v / 1e3 if unit.startswith("n") else v * (1e3 if unit.startswith("m") else 1.0)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print("# %s   (total %.1f us over %d launches)" % (path, tot, sum(v[0] for v in agg.values())))
print("%-70s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %12.1f %10.2f %6.1f%%" % (k[:70], n, us, us / n, 100 * us / tot))
