"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name and share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))[skip:]
agg, tot = collections.OrderedDict(), 0.0
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"fhe_b200::|<unnamed>::|void ", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%10.1f us %4dx %5.1f%%  %s" % (v, c, 100 * v / tot, k[:80]))
print("total %.1f us over %d launches" % (tot, len(rows)))
