"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    name = r[4].split("(")[0][-70:]
    t = float(r[-1].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print("%-72s %5s %10s %6s" % ("kernel", "n", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-72s %5d %10.3f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("%-72s %5d %10.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
