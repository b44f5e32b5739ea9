"""configs[3] record alone (bench.py config4_record): STGCN.forward + backward at 14,641 nodes, batch 32.
    python tools/config4_bench.py            # timing
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4.csv python tools/config4_bench.py 8 8 1
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
peak, _ = bench.peaks()
print(json.dumps(bench.config4_record(peak, bench.tensor_peak(), batch=batch, chunk=chunk, reps=reps)))
