"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    agg = defaultdict(lambda: [0, 0.0, set()])
    for name, ns, grid, blk in rows:
        short = re.sub(r"\(.*$", "", name)
        agg[short][0] += 1
        agg[short][1] += ns
        agg[short][2].add(grid)
    total = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}  grids")
    for k, (n, ns, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:5d} {ns / 1e3:10.1f} {ns / 1e3 / n:8.2f} {100 * ns / total:5.1f}%  {sorted(grids)[:4]}")


if __name__ == "__main__":
    main(sys.argv[1])
