#!/usr/bin/env python
"""ncu report -> profiles/<tag>_ncu_traffic.csv (what bench.py's `roofline.traffic` is read from).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv profiles/r2_ncu_traffic.csv

One output row per kernel name: launches captured, dram__bytes_read.sum and dram__bytes_write.sum summed over them
(bench.py divides by the launch count), plus the duration, DRAM throughput, tensor-pipe activity and registers that
the per-round summaries under profiles/ quote.
"""
import csv
import re
import sys
from collections import OrderedDict

WANT = {
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__time_duration.sum": "duration_ns",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6,
        "nsecond": 1.0, "second": 1e9}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    return name.strip()


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {h: i for i, h in enumerate(hdr)}
    out = OrderedDict()
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        k = short(r[col["Kernel Name"]])
        acc = out.setdefault(k, {"launches": 0, **{v: 0.0 for v in WANT.values()}})
        acc["launches"] += 1
        for metric, key in WANT.items():
            if metric not in col:
                continue
            try:
                val = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            val *= UNIT.get(units[col[metric]], 1.0)
            if key in ("dram_pct", "tensor_pct", "tensor_inst_pct", "registers", "grid"):
                acc[key] = max(acc[key], val)
            else:
                acc[key] += val
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches"] + list(WANT.values()))
        for k, a in out.items():
            w.writerow([k, a["launches"]] + [f"{a[v]:.6g}" for v in WANT.values()])
    print(f"{dst}: {len(out)} kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
