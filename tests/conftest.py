import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    import __graft_entry__ as g

    g.build()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_case(name):
    """(golden npz, cfg dict, state_dict, features, edge_index) of a hybrid fixture; the full-size
    cases regenerate weights/features from their seeds exactly as oracle/make_golden.py did."""
    from weatherforecast_stgcn_maml_b200 import synth

    z = load_golden(name)
    cin, hidden, L, layers, out, T, H = (int(v) for v in z["cfg"])
    cfg = dict(cin=cin, hidden=hidden, L=L, layers=layers, out=out, T=T, H=H,
               nlat=int(z["nlat"]), nlon=int(z["nlon"]), k=int(z["k"]))
    seed = int(z["seed"])
    n = cfg["nlat"] * cfg["nlon"]
    if "features" in z.files:
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        order = [k for k, _ in synth.v5_shapes(in_channels=cin, hidden=hidden, lstm_hidden=L, lstm_layers=layers,
                                               out_channels=out, horizon=H)]
        sd = {k: sd[k] for k in order}
        feats = torch.from_numpy(z["features"])
    else:
        sd = synth.init_v5_state_dict(seed, gcn_bias_scale=0.05, in_channels=cin, hidden=hidden, lstm_hidden=L,
                                      lstm_layers=layers, out_channels=out, horizon=H)
        feats = synth.synth_features(T + H + 8, n, seed + 1, synth.koppen_table(seed)[3])
    ei = torch.from_numpy(z["edge_index"].astype(np.int64))
    return z, cfg, sd, feats, ei


def sample_indices(numel, count=64):
    g = np.random.RandomState(1234 + numel % 9973)  # same rule as oracle/make_golden.py
    return np.sort(g.choice(numel, size=min(count, numel), replace=False))


def check_summary(t, summary, samples, rtol, what=""):
    """Compare a tensor against the (norm, sum) + sampled-values fixture of a big tensor."""
    a = t.detach().double().reshape(-1).cpu().numpy()
    nrm = float(np.sqrt((a * a).sum()))
    scale = max(float(np.abs(samples).max()), 1e-30)
    assert abs(nrm - summary[0]) <= rtol * max(summary[0], 1e-30), f"{what}: norm {nrm} vs {summary[0]}"
    got = a[sample_indices(a.size)]
    err = np.abs(got - samples.astype(np.float64)).max() / scale
    assert err <= rtol, f"{what}: sampled values rel err {err}"


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
