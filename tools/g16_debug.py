import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from weatherforecast_stgcn_maml_b200 import _lib
def run(rows_g, G, K, N):
    torch.manual_seed(0)
    A = torch.randn(G * rows_g, K, device="cuda"); W = torch.randn(G, N, K, device="cuda") / K ** 0.5
    Whi = torch.empty(G, N, K, dtype=torch.int16, device="cuda"); Wlo = torch.empty_like(Whi)
    C = torch.full((G * rows_g, N), float("nan"), device="cuda"); err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = _lib.stream_ptr()
    _lib.call("wf_split16", _lib.ptr(W), _lib.ptr(Whi), _lib.ptr(Wlo), W.numel(), 0, st)
    _lib.call("wf_g16_gemm_nt", _lib.ptr(A), rows_g, G, K, _lib.ptr(Whi), _lib.ptr(Wlo), N * K, N, None, None, N, 0, 0, _lib.ptr(C), _lib.ptr(err), st)
    torch.cuda.synchronize()
    ref = torch.bmm(A.double().view(G, rows_g, K), W.double().transpose(1, 2)).view(G * rows_g, N)
    d = (C.double() - ref).abs()
    mt = (rows_g + 127) // 128
    bad = []
    for g in range(G):
        for m in range(mt):
            for n in range(N // 128):
                blk = d[g * rows_g + m * 128: g * rows_g + min((m + 1) * 128, rows_g), n * 128:(n + 1) * 128]
                if float(blk.max()) > 1e-4 or torch.isnan(blk).any(): bad.append((g, m, n))
    print(f"rows {rows_g} G {G} K {K} N {N}: err {int(err.item())} bad tiles {len(bad)} of {G*mt*(N//128)}: first {bad[:12]}")
    if bad:
        rbs = sorted({g * mt + m for g, m, n in bad}); print("   bad row blocks:", rbs[:20], "...", "cols of first:", [n for g, m, n in bad if g * mt + m == rbs[0]])
for cfg in ((10584, 2, 256, 512), (10584, 2, 256, 256), (25000, 1, 128, 256), (25000, 1, 64, 256), (20000, 1, 256, 256)):
    run(*cfg)
