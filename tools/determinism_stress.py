"""Race detector by repetition: the hybrid pass is deterministic (fixed-order reductions, no atomics on data), so repeated
runs on the same inputs must be BIT-identical; a protocol race in the persistent kernels shows up as a differing bit."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable, gcn_weights_from_state_dict
from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs

def run_case(nlat, nlon, G, Bw, reps):
    n = nlat * nlon
    dims = V5Dims(num_nodes=n)
    lats, lons = synth.region_grid(nlat, nlon)
    ei = P.knn_edges_canonical(lats, lons, 8 if n > 9 else 2)
    sd = synth.init_v5_state_dict(5, gcn_bias_scale=0.05)
    time_rows = dims.window + dims.horizon + 1 + Bw + 2
    feats = torch.stack([synth.synth_features(time_rows, n, 100 + g) for g in range(G)]).cuda()
    per, per_task = n * 24, time_rows * n * 24
    xo = torch.tensor([g * per_task + b * per for g in range(G) for b in range(Bw)], device="cuda")
    to = xo + (dims.window + 1) * per
    theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
    graphs = StackedGraphs([RegionGraph(ei, dims.R, "cuda") for _ in range(G)])
    gcn_w = gcn_weights_from_state_dict(sd, "cuda")
    eng = HybridEngine(dims, G, Bw, "cuda")
    first, bad = None, 0
    for i in range(reps):
        loss, grads = eng.forward_backward(feats, 24, 0, xo, gcn_w, graphs, theta, eng.P, feat=feats, tgt_off=to, feat_ld=24)
        eng.check()
        cur = (eng.pred.clone(), loss.clone(), grads.clone())
        if first is None: first = cur
        elif not all(torch.equal(a, b) for a, b in zip(first, cur)): bad += 1
    print(f"{nlat}x{nlon} nodes, G={G}, Bw={Bw}: {reps} runs, {bad} differ from the first, grads finite {bool(torch.isfinite(first[2]).all())}")
    return bad

tot = 0
for cfg in ((21, 21, 15, 1, 40), (21, 21, 3, 4, 30), (3, 43, 2, 2, 30), (10, 13, 1, 3, 30), (2, 64, 4, 1, 30), (33, 33, 2, 1, 20)):
    tot += run_case(*cfg)
print("TOTAL differing runs:", tot)
sys.exit(1 if tot else 0)
