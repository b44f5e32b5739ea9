"""Soak: 300 graph-replayed meta-steps of the benchmark shape without and with dropout; the loss must fall, the weights stay
finite and no kernel may flag an error."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from weatherforecast_stgcn_maml_b200 import synth
from weatherforecast_stgcn_maml_b200.engine import V5Dims
from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer
dims = V5Dims(num_nodes=bench.NLAT * bench.NLON)
sd = synth.init_v5_state_dict(42)
tasks = bench.build_tasks(0, 1)
for drop in ((0, 0, 0), (0.2, 0.2, 0.2)):
    tr = MetaTrainer(sd, tasks, dims, "cuda", use_cuda_graph=True, support_rows=bench.SUPPORT_ROWS, accum=15, dropout=drop)
    losses = []
    t0 = time.time()
    for i in range(300):
        tr.meta_step()
        if i % 50 == 49:
            losses.append(round(tr.read_loss(), 5))
    torch.cuda.synchronize()
    tr.check()
    print("dropout", drop, "300 meta-steps in %.1f s" % (time.time() - t0), "losses", losses, "theta finite", bool(torch.isfinite(tr.theta).all()))
