"""The oracle port (oracle/ref_port.py) against the fixtures frozen from the unmodified reference
(oracle/make_golden.py).  CPU only; this is what pins the oracle outside the build container."""
import numpy as np
import pytest
import torch

from conftest import check_summary, golden_case, load_golden, rel_err
from oracle import ref_port as P
from weatherforecast_stgcn_maml_b200 import synth


@pytest.mark.parametrize("key", ["21x21_k4", "21x21_k8", "5x7_k4", "3x3_k2", "40x30_k4"])
def test_knn_ckdtree_matches_fixture(key):
    z = load_golden("knn_ckdtree")
    grid, k = key.split("_k")
    nlat, nlon = (int(v) for v in grid.split("x"))
    lats, lons = synth.region_grid(nlat, nlon)
    ei = P.knn_edges_ckdtree(lats, lons, int(k))
    assert ei.dtype == torch.int64 and tuple(ei.shape) == (2, nlat * nlon * int(k))
    assert np.array_equal(ei.numpy(), z[key].astype(np.int64))


@pytest.mark.parametrize("key", ["21x21_k4", "21x21_k8", "5x7_k4", "3x3_k2"])
def test_canonical_knn_vs_ckdtree_contract(key):
    """Same distance multiset everywhere; same neighbour set wherever no tie straddles the k-th place."""
    z = load_golden("knn_ckdtree")
    grid, k = key.split("_k")
    k = int(k)
    nlat, nlon = (int(v) for v in grid.split("x"))
    n = nlat * nlon
    lats, lons = synth.region_grid(nlat, nlon)
    ref = torch.from_numpy(z[key].astype(np.int64))
    can = P.knn_edges_canonical(lats, lons, k)
    d_ref = P.knn_sq_distances(lats, lons, ref).reshape(n, k)
    d_can = P.knn_sq_distances(lats, lons, can).reshape(n, k)
    assert np.array_equal(np.sort(d_ref, 1), d_can)  # canonical rows are already sorted
    pos = P.node_positions(lats, lons)
    differing = 0
    for i in range(n):
        d2 = ((pos - pos[i]) ** 2).sum(1)
        d2[i] = np.inf
        kth = np.sort(d2)[k - 1]
        straddle = (d2 == kth).sum() > (d_can[i] == kth).sum()
        same = set(ref[1, i * k:(i + 1) * k].tolist()) == set(can[1, i * k:(i + 1) * k].tolist())
        assert same or straddle, f"node {i}: sets differ without a straddling tie"
        differing += not same
    if key == "21x21_k8":
        assert differing <= 8  # only the edge nodes next to the corners


def test_dataset_window_layout():
    feats = synth.synth_features(60, 6, 0)
    assert P.num_windows(feats, 24, 8) == 28
    x, y = P.window_xy(feats, 3, 24, 8)
    assert x.shape == (24 * 6, 24) and y.shape == (8 * 6, 12)
    assert torch.equal(x[0], feats[3, 0]) and torch.equal(x[6 * 23 + 2], feats[26, 2])
    assert torch.equal(y[0], feats[28, 0, :12]) and torch.equal(y[6 * 7 + 5], feats[35, 5, :12])


@pytest.mark.parametrize("name", ["hybrid_small", "hybrid_v5_k4", "hybrid_v5_k8"])
def test_port_forward_backward_matches_reference(name):
    z, cfg, sd, feats, ei = golden_case(name)
    T, H = cfg["T"], cfg["H"]
    assert list(sd.keys()) == [str(k) for k in z["state_dict_keys"]]
    x, y = P.window_xy(feats, 0, T, H)
    loss, grads, pred = P.loss_and_grads(sd, x, y, ei, T, H, 1.0, cfg["layers"])
    assert rel_err(pred, torch.from_numpy(z["pred"])) <= 5e-6
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * float(z["loss"])
    assert sorted(grads) == sorted(P.trainable(sd)) and len(grads) == 2 + 4 * cfg["layers"]
    for k, g in grads.items():
        check_summary(g, z[f"grad_summary/{k}"], z[f"grad_samples/{k}"], 5e-5, k)
    bf = P.gcn_stack(sd, x, ei)
    check_summary(bf, z["base_features_summary"], z["base_features_samples"], 1e-5, "base features")


@pytest.mark.parametrize("name", ["hybrid_small", "hybrid_v5_k4"])
def test_port_stgcn_forward_backward_matches_reference(name):
    z, cfg, sd, feats, ei = golden_case(name)
    T, H = cfg["T"], cfg["H"]
    x, y = P.window_xy(feats, 0, T, H)
    leaf = {k[len("base_stgcn."):]: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("base_stgcn.")}
    xs = x.clone().requires_grad_(True)
    pred = P.stgcn_forward(leaf, xs, ei, T, H, cfg["out"])
    loss = torch.nn.functional.mse_loss(pred, y)
    grads = torch.autograd.grad(loss, list(leaf.values()) + [xs])
    assert rel_err(pred, torch.from_numpy(z["stgcn_pred"])) <= 5e-6
    for (k, _), g in zip(leaf.items(), grads):
        check_summary(g, z[f"stgcn_grad_summary/{k}"], z[f"stgcn_grad_samples/{k}"], 5e-5, k)
    check_summary(grads[-1], z["stgcn_dx_summary"], z["stgcn_dx_samples"], 5e-5, "dx")


@pytest.mark.parametrize("name", ["hybrid_small", "hybrid_v5_k4"])
def test_port_inner_loop_and_fomaml_match_reference(name):
    z, cfg, sd, feats, ei = golden_case(name)
    steps, accum = int(z["inner_steps"]), int(z["accum"])
    kw = dict(window=cfg["T"], horizon=cfg["H"], lr=float(z["inner_lr"]), lstm_layers=cfg["layers"])
    qloss, qgrads, fast = P.fomaml_task(sd, feats, ei, list(range(steps)), steps, accum, **kw)
    assert abs(float(qloss) - float(z["query_loss_scaled"])) <= 2e-5 * float(z["query_loss_scaled"])
    for k in P.trainable(sd):
        check_summary(fast[k], z[f"adapted_summary/{k}"], z[f"adapted_samples/{k}"], 5e-5, "adapted " + k)
        check_summary(qgrads[k], z[f"fomaml_summary/{k}"], z[f"fomaml_samples/{k}"], 2e-4, "fomaml " + k)
    for k in sd:
        if k.startswith("base_stgcn."):
            assert torch.equal(fast[k], sd[k])  # SGD never touches grad=None parameters (D4)


def test_literal_reference_grad_is_stale_plus_query():
    """SURVEY.md D5 detail: the reference's copy keeps the last inner step's clipped grad."""
    z, cfg, sd, feats, ei = golden_case("hybrid_small")
    steps, accum = int(z["inner_steps"]), int(z["accum"])
    kw = dict(window=cfg["T"], horizon=cfg["H"], lr=float(z["inner_lr"]), lstm_layers=cfg["layers"])
    fast, _ = P.inner_loop(sd, feats, ei, list(range(steps - 1)), **kw)
    x, y = P.window_xy(feats, steps - 1, cfg["T"], cfg["H"])
    _, g_last, _ = P.loss_and_grads(fast, x, y, ei, cfg["T"], cfg["H"], 1.0, cfg["layers"])
    gl = [g_last[k].clone() for k in g_last]
    P.clip_grad_norm(gl, 1.0)
    _, qgrads, _ = P.fomaml_task(sd, feats, ei, list(range(steps)), steps, accum, **kw)
    for k, stale in zip(g_last, gl):
        check_summary(stale, z[f"stale_summary/{k}"], z[f"stale_samples/{k}"], 1e-4, "stale " + k)
        check_summary(stale + qgrads[k], z[f"literal_summary/{k}"], z[f"literal_samples/{k}"], 2e-4, "literal " + k)


FEATURE_CASES = ("nan", "allnan", "clean", "given", "raw")
GIVEN_STATS = {"mean": [float(x) for x in np.linspace(-2.0, 3.0, 12)], "std": [float(x) for x in np.linspace(0.5, 4.0, 12)]}


def feature_case_inputs(z, name):
    """(weather copy, time features, Koppen row, normalize, stats) of one features_prepare.npz case."""
    w = z[f"{name}_weather"].copy()
    tf = P.time_features(z[f"{name}_doy"], z[f"{name}_tod"])
    row = torch.from_numpy(z["koppen_weight"])[int(z[f"{name}_code"])][None, :]
    return w, tf, row, name != "raw", (GIVEN_STATS if name == "given" else None)


@pytest.mark.parametrize("name", FEATURE_CASES)
def test_port_prepare_features_matches_reference(name):
    """oracle restatement of prepare_model_input (featurePreprocessor.py:84-182) == the reference's own output, bit for
    bit (numpy and torch are the same libraries on both sides), incl. NaN fill, an all-NaN variable, given statistics
    (f64 arithmetic) and normalize=False."""
    import warnings

    z = load_golden("features_prepare")
    w, tf, row, normalize, stats = feature_case_inputs(z, name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        feats, st = P.prepare_features(w, tf, row, normalize=normalize, stats=stats)
    assert feats.dtype == torch.float32 and tuple(feats.shape) == (40, 35, 24)
    assert np.array_equal(feats.numpy(), z[f"{name}_features"])
    if normalize:
        assert np.array_equal(np.asarray(st["mean"]), z[f"{name}_mean"]) and np.array_equal(np.asarray(st["std"]), z[f"{name}_std"])
    assert not np.isnan(feats.numpy()).any()


def test_time_features_formula():
    """embed_utils.py:12-26: year phase over 365.25 days, day phase over 24 h, order sin/cos year, sin/cos day."""
    from weatherforecast_stgcn_maml_b200.embed_utils import time_features

    tf = time_features([1, 100, 366], [0.0, 6.0, 23.5])
    assert tf.shape == (3, 4) and tf.dtype == np.float64
    assert np.allclose(tf[1], [np.sin(2 * np.pi * 100 / 365.25), np.cos(2 * np.pi * 100 / 365.25), 1.0, 0.0], atol=1e-15)
    assert np.array_equal(tf, P.time_features([1, 100, 366], [0.0, 6.0, 23.5]))


def test_masked_port_reproduces_the_reference_in_train_mode():
    """hybrid_small_dropout.npz: the UNMODIFIED reference in ``.train()`` mode with dropout_rate = lstm_dropout = 0.2, its
    masks replayed from torch's generator (oracle/make_golden.py:dropout_case).  The restatement with those masks as
    explicit inputs must give the reference's predictions, loss and gradients: this pins where each of the three dropout
    sites sits (after GCN layers 1-3, between LSTM layers, on the head input) and its 1/(1-p) scaling."""
    z = load_golden("hybrid_small_dropout")
    cin, hidden, L, layers, out, T, H = (int(v) for v in z["cfg"])
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    feats, ei = torch.from_numpy(z["features"]), torch.from_numpy(z["edge_index"].astype(np.int64))
    p = float(z["p"])
    masks = {"gcn": [torch.from_numpy(z[f"mask_gcn{i}"]) for i in range(3)],
             "lstm": [torch.from_numpy(z[f"mask_lstm{l}"]) for l in range(layers - 1)],
             "head": torch.from_numpy(z["mask_head"])}
    for m in masks["gcn"] + masks["lstm"] + [masks["head"]]:
        vals = torch.unique(m)
        assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / (1 - p)) < 1e-6
    x, y = P.window_xy(feats, 0, T, H)
    loss, grads, pred = P.loss_and_grads(sd, x, y, ei, T, H, 1.0, layers, masks=masks)
    assert rel_err(pred, torch.from_numpy(z["pred"])) <= 2e-6
    assert abs(float(loss) - float(z["loss"])) <= 1e-6 * float(z["loss"])
    for k, g in grads.items():
        assert rel_err(g, torch.from_numpy(z[f"grad/{k}"])) <= 2e-5, k
    # without the masks (eval mode) the result is a different one: the fixture really exercises dropout
    assert rel_err(P.hybrid_forward(sd, x, ei, T, H, out, layers), torch.from_numpy(z["pred"])) > 1e-2
    # STGCN.forward: dropout after all four convolutions, differentiable (model.py:31-42)
    base_sd = {k[len("base_stgcn."):]: v for k, v in sd.items() if k.startswith("base_stgcn.")}
    leaf = {k: v.clone().requires_grad_(True) for k, v in base_sd.items()}
    xr = x.clone().requires_grad_(True)
    m4 = [torch.from_numpy(z[f"stgcn_mask{i}"]) for i in range(4)]
    pr = P.stgcn_forward(leaf, xr, ei, T, H, out, masks=m4)
    gr = torch.autograd.grad(torch.nn.functional.mse_loss(pr, y), list(leaf.values()) + [xr])
    assert rel_err(pr, torch.from_numpy(z["stgcn_pred"])) <= 2e-6
    for (k, _), g in zip(leaf.items(), gr):
        assert rel_err(g, torch.from_numpy(z[f"stgcn_grad/{k}"])) <= 2e-5, k
    assert rel_err(gr[-1], torch.from_numpy(z["stgcn_dx"])) <= 2e-5


def test_reference_archive_recipe_round_trips(tmp_path):
    """oracle/build_ref.py: what travels to the GPU box as the reference arm is the reference's own files, unchanged."""
    import os

    from oracle import build_ref

    src = tmp_path / "src"
    src.mkdir()
    (src / "a.py").write_text("x = 1\n")
    (src / "b.py").write_text("import a\ny = a.x + 1\n")
    arc = tmp_path / "out" / "ref.tar.gz"
    assert build_ref.build(str(src), str(arc)) == ["a.py", "b.py"]
    first = arc.read_bytes()
    assert build_ref.build(str(src), str(arc)) == ["a.py", "b.py"] and arc.read_bytes() == first  # idempotent
    assert build_ref.build(str(tmp_path / "missing"), str(arc)) == []
    if os.path.isdir(build_ref.REF_SRC) and os.path.exists(build_ref.ARCHIVE):
        import tarfile

        with tarfile.open(build_ref.ARCHIVE) as tar:
            for m in tar.getmembers():
                assert tar.extractfile(m).read() == open(os.path.join(build_ref.REF_SRC, m.name), "rb").read(), m.name
