"""The C-ABI library loads on a CPU-only box and exports exactly what include/wf_stgcn.h declares."""
import ctypes as C

from abi_parse import parse_header
from weatherforecast_stgcn_maml_b200 import _lib

TAG = {C.c_void_p: "p", C.c_longlong: "ll", C.c_int: "i", C.c_float: "f", C.c_size_t: "sz", C.c_char_p: "char_p"}


def test_header_and_binding_agree():
    hdr = parse_header()
    assert len(hdr) >= 20
    assert set(hdr) == set(_lib.SIGNATURES), set(hdr) ^ set(_lib.SIGNATURES)
    for name, (ret, args) in hdr.items():
        pres, pargs = _lib.SIGNATURES[name]
        assert TAG[pres] == ret, name
        assert [TAG[a] for a in pargs] == args, name


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in parse_header():
        assert hasattr(lib, name), name
    assert lib.wf_abi_version() == 2


def test_param_count_matches_reference_model():
    # 606,304 grad-bearing elements of the v5 model
    assert _lib.query("wf_param_count", 4, 256, 128, 96) == 606304
    assert _lib.query("wf_param_count", 9, 256, 128, 96) == -1


def test_workspace_queries_are_pure():
    assert _lib.query("wf_gcn_norm_workspace_bytes", 1764, 10584) > 0
    assert _lib.query("wf_lstm_bwd_workspace_bytes", 4, 256, 128, 24, 441, 15, 1) > 0
    assert _lib.query("wf_optim_workspace_bytes", 15) >= 15 * 64 * 4


def test_no_cpu_fallback():
    import pytest
    import torch
    from weatherforecast_stgcn_maml_b200.engine import HybridEngine, V5Dims

    with pytest.raises(RuntimeError):
        HybridEngine(V5Dims(num_nodes=4), 1, 1, device="cpu")
    with pytest.raises(RuntimeError):
        _lib.require_cuda(torch.zeros(3))
