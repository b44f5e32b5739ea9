"""Per-step timeline of the backward LSTM kernel.  Needs a library built with -DWF_SEQ_TRACE (see tools/trace_fwd.py)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from weatherforecast_stgcn_maml_b200 import _lib, synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable
G = int(sys.argv[1]) if len(sys.argv) > 1 else 15
dims = V5Dims()
eng = HybridEngine(dims, G, 1, "cuda")
sd = synth.init_v5_state_dict(42)
theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
eng.feats = torch.randn(eng.rows, dims.hidden, device="cuda").relu()
eng.dpred.normal_()
for _ in range(3):
    eng.lstm_head_forward(theta, eng.P); eng.backward(theta, eng.P)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros(32 * 16 * 24, dtype=np.int64)
lib.wf_seq_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
tr = buf.reshape(32, 16, 24)
names = ["dfull", "sent", "x_ready", "chunk0", "chunk1", "chunk2", "chunk3", "staged", "mma issued", "stores done"]
for s in (8, 9, 10):
    base = tr[s, 0, 0]
    print(f"step {s}: period {tr[s+1,0,0]-tr[s,0,0]} cycles")
    for w in (0, 1, 3, 4, 7):
        print(f"  warp {w}: " + "  ".join(f"{names[p]}={tr[s,w,p]-base:6d}" for p in range(10) if tr[s, w, p]))
