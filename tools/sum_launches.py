"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys

t = collections.Counter()
n = collections.Counter()
for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")):
    try:
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        t[name] += ms
        n[name] += 1
    except Exception:
        pass
tot = sum(t.values())
for k, v in t.most_common(12):
    print("%-44s %6d launches %10.3f ms %5.1f%%" % (k[:44], n[k], v, 100 * v / tot))
print("total %.3f ms" % tot)
