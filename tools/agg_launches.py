"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel / grid."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = (r[ki][:64], r[gi], r[bi])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v[1] / 1e3:10.1f} us  n={v[0]:5d} avg={v[1] / v[0] / 1e3:8.2f} us {100 * v[1] / tot:5.1f}%  {k[0]} {k[1]} {k[2]}")
