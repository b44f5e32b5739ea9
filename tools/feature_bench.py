"""GB/s of the feature-assembly passes (csrc/wf_features.cu) at 632 steps x 14,641 nodes (configs[3] graph)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from weatherforecast_stgcn_maml_b200.featurePreprocessor import assemble_features, feature_stats

T, N = 632, 14641
w = torch.randn(T, N, 12, device="cuda") * 3 + 5
w[torch.rand(T, N, 12, device="cuda") < 0.01] = float("nan")
tf = np.random.rand(T, 4).astype(np.float32)
row = np.arange(8, dtype=np.float32)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
st = feature_stats(w)
mean, std = st["mean"].astype(np.float32), st["std"].astype(np.float32) + np.float32(1e-8)
rows = T * N
t_s = timeit(lambda: feature_stats(w))
t_a = timeit(lambda: assemble_features(w, tf, row, fill=st["fill"], mean=mean, std=std))
print(f"rows {rows}: stats (2 passes, 96 B/row) {t_s*1e3:.3f} ms = {rows*96/t_s/1e9:.0f} GB/s incl. readback sync; "
      f"assemble (48 B in + 96 B out per row) {t_a*1e3:.3f} ms = {rows*144/t_a/1e9:.0f} GB/s incl. output allocation")
