"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ns, share.
usage: python scripts/launch_summary.py gpurun_out/rXX_launches.csv > profiles/rXX_launches_summary.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    a = agg.setdefault(r[ix["Kernel Name"]][:90], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ns", "share", "mean_ns"])
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, a[0], int(a[1]), round(a[1] / tot, 4), int(a[1] / a[0])])
