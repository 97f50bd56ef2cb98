"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the summed device time).
usage: python profiles/summarize_launches.py <launches.csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:48]
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-50s %6s %12s %8s" % ("kernel", "n", "avg us", "share"))
for k, v in agg.items():
    print("%-50s %6d %12.1f %7.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
