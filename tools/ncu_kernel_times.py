"""Sum gpu__time_duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    d[r[ik][:70]][0] += 1; d[r[ik][:70]][1] += float(r[iv].replace(",", ""))
for k, v in d.items():
    print("%-72s %4d launches %9.3f ms" % (k, v[0], v[1] / 1e6))
