"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total ms and launches per kernel."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = None
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    grid = {}
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(d["Metric Unit"], 1e-6)
            k = d["Kernel Name"][:60]
            tot[k] += v
            cnt[k] += 1
            grid[k] = (d.get("Grid Size"), d.get("Block Size"))
    print(f)
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:14]:
        print(f"{v:10.2f} ms {cnt[k]:4d}  {grid[k]}  {k}")
