"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...) per kernel name.
   python tools/launch_summary.py X.csv "<command line that was profiled>" """
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    tot += ms
print(f"# ncu launch list summary (gpu__time_duration.sum, --clock-control none), command: {' '.join(sys.argv[2:])}")
print("# cold-cache serialised times: compare SHARES, not absolutes")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% n={n:4d} {k[-72:]}")
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
