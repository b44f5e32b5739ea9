"""ctypes binding of libwf_stgcn.so (the C ABI declared in include/wf_stgcn.h).

This is the reference-side stub INTEGRATION.md describes: plain pointers and
sizes, no torch types cross the boundary.  There is no fallback: if the shared
library is missing or a launcher reports an error, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwf_stgcn.so")

c_p = C.c_void_p
c_ll = C.c_longlong
c_i = C.c_int
c_f = C.c_float
c_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/wf_stgcn.h declares
SIGNATURES = {
    "wf_abi_version": (c_i, []),
    "wf_last_error": (C.c_char_p, []),
    "wf_param_count": (c_ll, [c_i, c_i, c_i, c_i]),
    "wf_knn_grid_build": (c_i, [c_p, c_i, c_p, c_i, c_i, c_i, c_p, c_p]),
    "wf_gcn_norm_workspace_bytes": (c_sz, [c_ll, c_i]),
    "wf_gcn_norm_csr": (c_i, [c_p, c_ll, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "wf_gcn_layer_fwd": (c_i, [c_p, c_i, c_ll, c_p, c_p, c_p, c_ll, c_ll, c_p, c_p, c_p, c_ll, c_ll,
                               c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "wf_gcn_layer_bwd_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i]),
    "wf_gcn_layer_bwd_ss_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "wf_gcn_layer_bwd_ss": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_i,
                                  c_p, c_p, c_p, c_p, c_sz, c_p, c_p]),
    "wf_gcn_layer_bwd": (c_i, [c_p, c_i, c_ll, c_p, c_p, c_p, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p,
                               c_ll, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_ll, c_ll,
                               c_p, c_sz, c_p]),
    "wf_lstm_fwd": (c_i, [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_f, c_p, c_p, c_p]),
    "wf_lstm_bwd_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "wf_lstm_bwd": (c_i, [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p,
                          c_ll, c_f, c_p, c_p, c_p, c_sz, c_p]),
    "wf_head_fwd": (c_i, [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "wf_mse_fwd_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p]),
    "wf_head_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i]),
    "wf_head_bwd": (c_i, [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_ll,
                          c_p, c_sz, c_p]),
    "wf_optim_workspace_bytes": (c_sz, [c_i]),
    "wf_clip_sgd_step": (c_i, [c_p, c_ll, c_p, c_ll, c_ll, c_i, c_f, c_f, c_p, c_p, c_sz, c_p]),
    "wf_clip_adam_step": (c_i, [c_p, c_p, c_p, c_p, c_ll, c_p, c_f, c_i, c_p, c_p, c_sz, c_p]),
    "wf_sum_groups": (c_i, [c_p, c_ll, c_i, c_ll, c_p, c_i, c_p]),
    "wf_feature_stats_workspace_bytes": (c_sz, [c_ll]),
    "wf_feature_stats": (c_i, [c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "wf_assemble_features": (c_i, [c_p, c_ll, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "wf_dropout_apply": (c_i, [c_p, c_ll, c_i, c_i, c_ll, c_i, c_f, c_p, c_i, c_p, c_p]),
    "wf_rng_advance": (c_i, [c_p, c_p]),
    "wf_param_count_transposed": (c_ll, [c_i, c_i, c_i, c_i]),
    "wf_prep_weights_tc": (c_i, [c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "wf_transposed_pitch": (c_ll, [c_i, c_i]),
    "wf_gcn_layer_fwd_tc": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                  c_p, c_p, c_p, c_p, c_p]),
    "wf_lstm_fwd_tc": (c_i, [c_p, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p,
                             c_p, c_p]),
    "wf_lstm_bwd_tc_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "wf_lstm_bwd_tc": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p,
                             c_p, c_ll, c_p, c_sz, c_p, c_p]),
    "wf_tc_wgrad": (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_ll, c_p, c_p]),
    "wf_split_lo": (c_i, [c_p, c_p, c_ll, c_p]),
    "wf_tb4_elems": (c_ll, [c_i, c_i, c_i, c_ll]),
    "wf_tile_rows": (c_i, [c_i]),
    "wf_seq_weight_elems": (c_ll, [c_i, c_i, c_i]),
    "wf_param_stride16": (c_ll, [c_i, c_i, c_i, c_i, c_i]),
    "wf_split16": (c_i, [c_p, c_p, c_p, c_ll, c_i, c_p]),
    "wf_transposed_pitch16": (c_ll, [c_i, c_i]),
    "wf_g16_gemm_nt": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_ll, c_i, c_p, c_p, c_ll, c_i, c_i, c_p, c_p, c_p]),
    "wf_gcn_layer_fwd_g16": (c_i, [c_p, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_p, c_i, c_ll, c_p, c_i, c_i, c_i, c_i,
                                   c_i, c_i, c_i, c_p, c_p, c_p, c_f, c_p, c_i, c_p, c_p]),
    "wf_prep_weights_seq": (c_i, [c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "wf_tb8_elems": (c_ll, [c_i, c_i, c_i, c_ll]),
    "wf_lstm_fwd_seq": (c_i, [c_p, c_p, c_p, c_p, c_ll, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p,
                              c_p, c_p, c_f, c_p, c_p, c_p, c_p]),
    "wf_lstm_bwd_seq_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "wf_lstm_bwd_seq": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p,
                              c_p, c_ll, c_f, c_p, c_p, c_p, c_sz, c_p, c_p]),
    "wf_tc_gemm_nt": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_ll, c_i, c_p, c_p, c_ll, c_i, c_p, c_p, c_p]),
    "wf_lstm_seq_recur_fwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "wf_lstm_seq_recur_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "wf_gcn_layer_fwd_ss": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_ll, c_i, c_p, c_i, c_i, c_i,
                                  c_i, c_i, c_i, c_p, c_p, c_f, c_p, c_i, c_p, c_p]),
    "wf_join16": (c_i, [c_p, c_p, c_ll, c_i, c_p, c_p]),
    "wf_ss_nodes_gemm": (c_i, [c_i, c_i, c_p, c_ll, c_i, c_i, c_p, c_p, c_ll, c_i, c_i, c_p, c_p, c_ll, c_p, c_i, c_i, c_i, c_i,
                               c_i, c_p, c_p]),
    "wf_ss_wgrad": (c_i, [c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_i, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                          c_p, c_ll, c_p, c_i, c_i, c_p, c_i, c_i, c_p, c_ll, c_p, c_p]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device (or host) address of a tensor, or NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Invoke a launcher; raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.wf_last_error().decode()}")


def query(name, *args):
    return getattr(load(), name)(*args)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("wf_stgcn kernels need CUDA tensors; there is no CPU fallback")
