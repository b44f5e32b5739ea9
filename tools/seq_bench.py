"""Micro-benchmark of the persistent LSTM kernels at the benchmark shape (CUDA events around the C-ABI calls)."""
import sys, torch
sys.path.insert(0, ".")
from weatherforecast_stgcn_maml_b200 import _lib, synth
from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims, flatten_trainable

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

G = 15
dims = V5Dims()
eng = HybridEngine(dims, G, 1, "cuda")
sd = synth.init_v5_state_dict(42)
theta = torch.stack([flatten_trainable(sd, dims) for _ in range(G)]).cuda()
eng.feats = torch.randn(eng.rows, dims.hidden, device="cuda").relu()
t_f = timeit(lambda: eng.lstm_head_forward(theta, eng.P))
eng.dpred.normal_()
def fb():
    eng.lstm_head_forward(theta, eng.P); eng.backward(theta, eng.P)
t_fb = timeit(fb)
print(f"lstm fwd (prep + 4 x (proj + recurrence) + head): {t_f:8.1f} us   fwd+bwd: {t_fb:8.1f} us   bwd: {t_fb - t_f:8.1f} us   err {int(eng.err.item())}")
