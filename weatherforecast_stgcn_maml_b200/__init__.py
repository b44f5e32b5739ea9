"""B200-native implementation of the Hybrid MAML-STGCN-LSTM v5 hot path.

Drop-in modules (same names, signatures and ``state_dict`` layout as the reference files):
``graphBuilder``, ``model``, ``hybrid_model``, ``embed_utils``, ``dataset``,
``train_hybrid_maml_v5``, ``adapt_hybrid_v5``, ``adaptive_scheduler``, plus ``schedule`` (outer LR schedule, task
sampler) and ``featurePreprocessor`` (de-normalisation, forecast metrics).  Every compute call goes
through the C ABI of ``libwf_stgcn.so`` (include/wf_stgcn.h, sm_100a CUDA); there is no CPU or
eager-PyTorch fallback -- a missing library or a CPU tensor raises.
"""
from . import _lib  # noqa: F401
from .engine import AdamState, HybridEngine, V5Dims  # noqa: F401

__version__ = "0.1.0"


def install_dropin_modules():
    """Register this package's modules under the reference's top-level module names
    (``model``, ``hybrid_model``, ``graphBuilder``, ``embed_utils``, ``dataset``,
    ``adaptive_scheduler``) so unmodified reference drivers import the CUDA path."""
    import importlib
    import sys

    for name in ("model", "hybrid_model", "graphBuilder", "embed_utils", "dataset", "adaptive_scheduler"):
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")
