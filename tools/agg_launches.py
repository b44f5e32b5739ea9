#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, total and mean time."""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
acc = OrderedDict()
for r in rows[hdr_i + 1 + skip:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r[kn])).replace("(anonymous namespace)::", "")
    t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1.0)
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in acc.values())
print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:5d} {t:10.1f} {t / n:9.2f} {t / tot:6.1%}")
print(f"{'TOTAL':70s} {sum(a[0] for a in acc.values()):5d} {tot:10.1f}")
