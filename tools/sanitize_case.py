"""One small forward + backward on the persistent tensor-core path, for `compute-sanitizer --tool memcheck`."""
import sys, torch
sys.path.insert(0, ".")
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable, gcn_weights_from_state_dict
from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

nlat, nlon, T, H, G = 5, 7, 3, 2, 2
n = nlat * nlon
dims = V5Dims(num_nodes=n, window=T, horizon=H)
lats, lons = synth.region_grid(nlat, nlon)
ei = P.knn_edges_canonical(lats, lons, 4)
sd = synth.init_v5_state_dict(9, gcn_bias_scale=0.05, horizon=H)
feats = torch.stack([synth.synth_features(T + H + 3, n, 200 + g) for g in range(G)]).cuda()
per, per_task = n * 24, (T + H + 3) * n * 24
xo = torch.tensor([g * per_task for g in range(G)], device="cuda")
to = xo + (T + 1) * per
theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
graphs = StackedGraphs([RegionGraph(ei, dims.R, "cuda") for _ in range(G)])
eng = HybridEngine(dims, G, 1, "cuda")
assert eng.seq
loss, grads = eng.forward_backward(feats, 24, 0, xo, gcn_weights_from_state_dict(sd, "cuda"), graphs, theta, eng.P,
                                   feat=feats, tgt_off=to, feat_ld=24)
eng.sgd_step(theta, 0.01)
torch.cuda.synchronize()
eng.check()
print("loss", loss.tolist(), "grad norm", float(grads.norm()))
