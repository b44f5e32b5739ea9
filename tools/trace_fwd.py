import os, sys, torch
sys.path.insert(0, ".")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 15
tr = torch.zeros(8 * 4 * 8, dtype=torch.int64, device="cuda")
os.environ["WF_TRACE_PTR"] = str(tr.data_ptr())
from weatherforecast_stgcn_maml_b200 import _lib, synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable
dims = V5Dims()
e = HybridEngine(dims, G, 1, "cuda")
sd = synth.init_v5_state_dict(42)
theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
e.feats = torch.randn(e.rows, dims.hidden, device="cuda").relu()
e.lstm_head_forward(theta, e.P)
d = dims
for _ in range(3):
    _lib.call("wf_lstm_seq_recur_fwd", _lib.ptr(e.gates[1]), _lib.ptr(e.c[1]), _lib.ptr(e.h[1]), _lib.ptr(e.hT[1]),
              _lib.ptr(e.hT_lo[1]), _lib.ptr(e.w16[0]), _lib.ptr(e.w16[1]), 1, 4, 128, d.window, d.num_nodes, e.G, e.Bw, _lib.ptr(e.err), _lib.stream_ptr())
torch.cuda.synchronize()
t = tr.cpu().view(8, 4, 8)
names = ["a_ready", "dfull", "chunk0", "chunk1", "chunk2", "chunk3", "arrive", "stored"]
base = t[0, 1, 0].item()
for step in (1, 2):
    print(f"--- step {8+step} (cycles relative to warp0 a_ready of step 9; 1.965 GHz => 1000 cyc = 0.51 us)")
    for w in range(8):
        print(f"  warp {w}: " + "  ".join(f"{names[p]}={t[w, step, p].item() - base:6d}" for p in range(8) if t[w, step, p].item() > 0))
print("step period (warp0 a_ready 10 - 9):", t[0, 2, 0].item() - t[0, 1, 0].item(), "cycles")
