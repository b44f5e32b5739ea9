"""A few un-graphed fine-tune steps (configs[2]: batch-1 Adam steps on one 441-node region) for a kernel-by-kernel profile:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ft_launches.csv python tools/one_finetune.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weatherforecast_stgcn_maml_b200 import synth  # noqa: E402
from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import FineTuner  # noqa: E402
from weatherforecast_stgcn_maml_b200.engine import V5Dims  # noqa: E402
from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dims = V5Dims(num_nodes=bench.NLAT * bench.NLON)
sd = synth.init_v5_state_dict(42)
lats, lons, feats, _ = synth.synth_task(7, num_windows=steps + 8, nlat=bench.NLAT, nlon=bench.NLON)
ei = knn_edge_index_device(lats, lons, bench.KNN, "cuda")
ft = FineTuner(sd, feats, ei, dims, "cuda", region_name="profile", max_samples=steps, train_frac=1.0, use_cuda_graph=False,
               dropout=(0, 0, 0))
for i in range(steps):
    ft.step(i)
torch.cuda.synchronize()
ft.engine.check()
print("steps", steps)
