"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = d.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    name = d["Kernel Name"].split("(")[0].replace("void ", "")[:56]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'n':>5s} {'ms':>10s} {'share':>7s} {'ms/launch':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {v[0]:5d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:10.4f}")
print(f"{'total':58s} {sum(v[0] for v in agg.values()):5d} {tot:10.3f}")
