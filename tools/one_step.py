"""One un-graphed meta-step of the benchmark shape (15 tasks, 4 window passes) for kernel-by-kernel profiling:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weatherforecast_stgcn_maml_b200 import synth  # noqa: E402
from weatherforecast_stgcn_maml_b200.engine import V5Dims  # noqa: E402
from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = V5Dims(num_nodes=bench.NLAT * bench.NLON)
sd = synth.init_v5_state_dict(42)
tasks = bench.build_tasks(0, 1)
drop = (0.2, 0.2, 0.2) if os.environ.get("WF_ONE_STEP_DROPOUT") else (0, 0, 0)
tr = MetaTrainer(sd, tasks, dims, "cuda", use_cuda_graph=False, support_rows=bench.SUPPORT_ROWS, accum=15, dropout=drop)
for _ in range(steps):
    tr.meta_step()
torch.cuda.synchronize()
print("loss", tr.read_loss())
