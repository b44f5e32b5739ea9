# quick GPU diagnostic: persistent vs fp32 path, prints errors per stage
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable, gcn_weights_from_state_dict, unflatten_trainable
from weatherforecast_stgcn_maml_b200.graph import RegionGraph, StackedGraphs
def run(nlat, nlon, T, G, Bw):
    n, H = nlat * nlon, 3
    dims = V5Dims(num_nodes=n, window=T, horizon=H)
    lats, lons = synth.region_grid(nlat, nlon)
    eis = [P.knn_edges_canonical(lats, lons, 4) for g in range(G)]
    base = synth.init_v5_state_dict(9, gcn_bias_scale=0.05, horizon=H)
    sds = [{k: (v + 0.02 * torch.randn_like(v) * (g > 0) if k.startswith(("lstm.", "output_layer.")) else v) for k, v in base.items()} for g in range(G)]
    time_rows = T + H + 1 + Bw + 2
    feats = torch.stack([synth.synth_features(time_rows, n, 200 + g) for g in range(G)])
    per, per_task = n * 24, time_rows * n * 24
    starts = [[(g + 2 * b) % (Bw + 2) for b in range(Bw)] for g in range(G)]
    xo = torch.tensor([g * per_task + s * per for g in range(G) for s in starts[g]], device="cuda")
    to = xo + (T + 1) * per
    theta = torch.stack([flatten_trainable(sd, dims) for sd in sds]).cuda()
    graphs = StackedGraphs([RegionGraph(ei, dims.R, "cuda") for ei in eis])
    fd = feats.cuda()
    out = {}
    for prec in ("fp32", "seq"):
        eng = HybridEngine(dims, G, Bw, "cuda", precision="fp32" if prec == "fp32" else "tf32x3")
        eng.gcn_forward(fd, 24, 0, xo, gcn_weights_from_state_dict(base, "cuda"), graphs)
        eng.lstm_head_forward(theta, eng.P)
        torch.cuda.synchronize(); print(prec, "fwd err flag", int(eng.err.item()))
        loss = eng.mse(feat=fd, tgt_off=to, feat_ld=24, grad_scale=1.0)
        hcl = eng.hidden_states()
        grads = eng.backward(theta, eng.P)
        torch.cuda.synchronize(); print(prec, "bwd err flag", int(eng.err.item()))
        out[prec] = (eng.pred.clone(), loss.clone(), grads.clone(), hcl)
    rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    p0, l0, g0, h0 = out["fp32"]; p1, l1, g1, h1 = out["seq"]
    for l in range(4):
        print(f"  h layer {l}: rel err {rel(h1[l], h0[l]):.3e}  nan={bool(torch.isnan(h1[l]).any())}")
    print(f"  pred {rel(p1,p0):.3e} loss {rel(l1,l0):.3e}")
    for g in range(G):
        a, b = unflatten_trainable(g1[g], dims), unflatten_trainable(g0[g], dims)
        print("  task", g, " ".join(f"{k.split('.')[-1]}:{rel(a[k], b[k]):.1e}" for k in a))
for cfg in [(5, 7, 6, 2, 2), (21, 21, 24, 1, 1), (12, 13, 5, 3, 1)]:
    print("config", cfg); run(*cfg)
