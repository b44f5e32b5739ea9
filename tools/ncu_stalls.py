"""Summarise warp-stall samples of one kernel from an .ncu-rep (source page): python tools/ncu_stalls.py rep kernel_regex [n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", "0",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, src = h.index("# Samples"), h.index("Source")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
body = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
agg = {h[i]: 0 for i in stall_cols}
for r in body:
    for i in stall_cols:
        if r[i].isdigit():
            agg[h[i]] += int(r[i])
print(rows[0][:2], "total samples", tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"  {k}: {v} ({100 * v / tot:.1f}%)")
for i in sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]:
    r = body[i]
    st = sorted([(int(r[j]), h[j]) for j in stall_cols if r[j].isdigit() and int(r[j]) > 0], reverse=True)[:2]
    print(f"  {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% idx={i} {r[src].strip()[:60]:60s} {st}")
    for rr in body[max(0, i - 3):i]:
        print(f"             prev: {rr[src].strip()[:80]}")
