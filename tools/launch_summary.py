"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name and grid."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
i = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i]
kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[i + 1:]:
    if len(r) <= mv:
        continue
    a = agg.setdefault((r[kn].split("(")[0][:48], r[gs]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:50s} {k[1]:>16s} x{v[0]:<5d} avg {v[1] / v[0] / 1e3:9.1f} us  share {100 * v[1] / tot:5.1f}%")
