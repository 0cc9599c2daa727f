"""Per-kernel DRAM bytes of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list,
grouped by (kernel, grid, duration bucket).   python profiles/summarize_dram.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
idx = {k: i for i, k in enumerate(h)}
per = collections.OrderedDict()
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    d = per.setdefault((r[idx["ID"]], r[idx["Kernel Name"]][:48], r[idx["Grid Size"]]), {})
    d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * MULT.get(r[idx["Metric Unit"]], 1)
agg = collections.OrderedDict()
for (_, k, g), d in per.items():
    t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
    a = agg.setdefault((k, g, round(t, -1) if t < 200 else round(t, -2)), [0, 0, 0, 0])
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
tot = sum(a[1] for a in agg.values())
print(f"{len(per)} launches, {tot:.1f} us, {sum(a[2] for a in agg.values()) / 1e9:.2f} GB read, {sum(a[3] for a in agg.values()) / 1e9:.2f} GB written")
for (k, g, _), a in sorted(agg.items(), key=lambda x: -x[1][1])[:18]:
    print(f"{k:50s} n={a[0]:3d} avg {a[1] / a[0]:8.1f} us  share {100 * a[1] / tot:5.1f}%  read {a[2] / a[0] / 1e6:8.1f} MB  write {a[3] / a[0] / 1e6:8.1f} MB")
