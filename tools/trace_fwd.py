"""Per-step timeline of the 16-warp forward LSTM kernel.  Needs a library built with -DWF_SEQ_TRACE:
   NVCC_EXTRA=-DWF_SEQ_TRACE python -c "import __graft_entry__ as g; g.build(force=True)"   (rebuild without afterwards)"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
os.environ["WF_SEQ_FWD16"] = "1"
from weatherforecast_stgcn_maml_b200 import _lib, synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable
G = int(sys.argv[1]) if len(sys.argv) > 1 else 15
dims = V5Dims()
eng = HybridEngine(dims, G, 1, "cuda")
sd = synth.init_v5_state_dict(42)
theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
eng.feats = torch.randn(eng.rows, dims.hidden, device="cuda").relu()
for _ in range(3): eng.lstm_head_forward(theta, eng.P)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros(32 * 16 * 24, dtype=np.int64)
rc = lib.wf_seq_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
tr = buf.reshape(32, 16, 24)
names = ["dfull seen", "phase A done", "arrived+prefetch", "a_ready seen (w0)", "mma issued (w0)", "phase B done"]
for t in (8, 9, 10):
    base = tr[t, 0, 0]
    print(f"step {t}: period {tr[t+1,0,0]-tr[t,0,0]} cycles")
    for w in (0, 1, 5, 10, 15):
        print(f"  warp {w:2d}: " + "  ".join(f"{names[p].split()[0]}{p}={tr[t,w,p]-base:6d}" for p in range(6) if tr[t, w, p]))
        print("           chunks (ld-done, math-done, A-written): " + " | ".join(" ".join(str(tr[t, w, 8 + 4 * c + i] - base) for i in range(3)) for c in range(4)))
