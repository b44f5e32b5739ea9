"""tcgen05 3xTF32 GEMM (wf_tc_gemm_nt) against float64 matmul: FP32-class accuracy on the tensor cores."""
import pytest
import torch

from weatherforecast_stgcn_maml_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(rows_g, G, K, N, bias=False, relu=False, seed=0):
    torch.manual_seed(seed)
    A = torch.randn(G * rows_g, K, device="cuda")
    W = torch.randn(G, N, K, device="cuda") / K ** 0.5
    Wlo = torch.empty_like(W)
    b1 = torch.randn(G, N, device="cuda") if bias else None
    b2 = torch.randn(G, N, device="cuda") if bias else None
    C = torch.full((G * rows_g, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = _lib.stream_ptr()
    _lib.call("wf_split_lo", _lib.ptr(W), _lib.ptr(Wlo), W.numel(), st)
    _lib.call("wf_tc_gemm_nt", _lib.ptr(A), rows_g, G, K, _lib.ptr(W), _lib.ptr(Wlo), N * K, N, _lib.ptr(b1), _lib.ptr(b2),
              N, int(relu), _lib.ptr(C), _lib.ptr(err), st)
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"pipeline timeout code {int(err.item())}"
    ref = torch.bmm(A.double().view(G, rows_g, K), W.double().transpose(1, 2))
    if bias:
        ref = ref + (b1 + b2).double()[:, None, :]
    if relu:
        ref = ref.clamp_min(0)
    ref = ref.view(G * rows_g, N)
    return float((C.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("rows_g,G,K,N", [(128, 1, 32, 128), (128, 1, 64, 256), (441, 1, 128, 512), (300, 3, 256, 256),
                                          (1000, 2, 128, 128), (10584, 1, 256, 512)])
def test_tc_gemm_matches_fp64(rows_g, G, K, N):
    assert _run(rows_g, G, K, N) <= 5e-6


def test_tc_gemm_bias_relu_epilogue():
    assert _run(441, 2, 256, 256, bias=True, relu=True) <= 5e-6
    assert _run(200, 1, 128, 512, bias=True) <= 5e-6


def test_split_lo_is_exact():
    x = torch.randn(4096, device="cuda") * 37.0
    lo = torch.empty_like(x)
    _lib.call("wf_split_lo", _lib.ptr(x), _lib.ptr(lo), x.numel(), _lib.stream_ptr())
    hi = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert torch.equal(hi + lo, x) and torch.all(lo.abs() <= x.abs() * 2.0 ** -10)
