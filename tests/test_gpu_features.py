"""Feature assembly on the device (csrc/wf_features.cu through the C ABI) against the reference's
prepare_model_input: the fixtures frozen from the unmodified reference (tests/golden/features_prepare.npz) and the numpy
oracle at other sizes; size-independent properties at the benchmark size (632 steps x 441 nodes).

Tolerances.  Channels 12..23 (time features, Koppen row), everything with normalize=False and everything with GIVEN
statistics (the reference then works in f64 and rounds once) are bit-exact on entries that were not NaN.  NaN entries are
filled with the per-variable nanmean, which numpy accumulates in f32 (pairwise) and the kernel in f64: 1e-6 relative.
Statistics the reference derives itself are numpy f32 running sums over f32 data (``mean(axis=(0, 1))`` adds the
rows one after the other), whose rounding error grows like sqrt(samples) * 2^-24 relative to |mean| + std (2.5e-6 observed
for 28,224 samples of a 98,000 Pa mean; the kernel accumulates in f64 and is the accurate side).  Bounds used, with
e = 1e-6 + 2 sqrt(samples) 6e-8:  |mean - ref| <= e (|mean| + std),  |std - ref| <= (1e-5 + 2 e) std,  and for the z-scores
computed from them the same expressed in standard deviations, e (|mean| + std) / std + (1e-5 + 2 e) |z| + 1e-6.
"""
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_port as P
from test_oracle_golden import FEATURE_CASES, feature_case_inputs

pytestmark = pytest.mark.gpu


class _Var:
    def __init__(self, values):
        self.values = values


def _dataset(w, tf):
    ds = {v: _Var(w[..., i]) for i, v in enumerate(P.WEATHER_VARS)}
    ds.update({v: _Var(tf[:, i]) for i, v in enumerate(P.TIME_VARS)})
    return ds


def _koppen(weight):
    from weatherforecast_stgcn_maml_b200.embed_utils import KoppenEmbedding

    k = KoppenEmbedding(8)
    with torch.no_grad():
        k.embedding.weight.copy_(torch.as_tensor(weight))
    return k


def _check(name, got, ref, nan_mask, st, ref_mean, ref_std, derived):
    got, ref = got.cpu().numpy(), np.asarray(ref)
    assert got.shape == ref.shape and got.dtype == np.float32 and not np.isnan(got).any()
    assert np.array_equal(got[..., 12:], ref[..., 12:]), name  # time features and Koppen row: bit-exact
    gw, rw = got[..., :12], ref[..., :12]
    if derived:  # statistics derived by both sides in different precisions
        e = 1e-6 + 2 * np.sqrt(gw.size / 12) * 6e-8
        assert np.all(np.abs(st["mean"] - ref_mean) <= e * (np.abs(ref_mean) + ref_std)), name
        assert np.all(np.abs(st["std"] - ref_std) <= (1e-5 + 2 * e) * ref_std), name
        tol = e * (np.abs(ref_mean) + ref_std) / ref_std + (1e-5 + 2 * e) * np.abs(rw) + 1e-6
        assert np.all(np.abs(gw - rw) <= tol), (name, float(np.abs(gw - rw).max()))
    else:
        assert np.array_equal(gw[~nan_mask], rw[~nan_mask]), name
        assert np.all(np.abs(gw[nan_mask] - rw[nan_mask]) <= 1e-6 * (1.0 + np.abs(rw[nan_mask]))), name


@pytest.mark.parametrize("name", FEATURE_CASES)
def test_prepare_model_input_matches_reference_fixture(name):
    from weatherforecast_stgcn_maml_b200.featurePreprocessor import prepare_model_input

    z = load_golden("features_prepare")
    w, tf, row, normalize, stats = feature_case_inputs(z, name)
    nan_mask = np.isnan(w).reshape(40, 35, 12)
    feats, st = prepare_model_input(_dataset(w, tf), int(z[f"{name}_code"]), _koppen(z["koppen_weight"]), normalize=normalize,
                                    stats=stats)
    assert feats.is_cuda and tuple(feats.shape) == (40, 35, 24)
    if normalize:
        _check(name, feats, z[f"{name}_features"], nan_mask, st, z[f"{name}_mean"], z[f"{name}_std"], derived=stats is None)
        if stats is None:
            assert np.asarray(st["mean"]).dtype == np.float32 and np.asarray(st["std"]).dtype == np.float32
    else:
        assert st == {}
        _check(name, feats, z[f"{name}_features"], nan_mask, None, None, None, derived=False)


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 2, 5), (17, 9, 4), (64, 21, 21)])
def test_prepare_model_input_matches_oracle_other_sizes(shape):
    """Ragged sizes (rows not a multiple of the block), a single row, the benchmark grid."""
    from oracle.make_golden import synth_raw_weather
    from weatherforecast_stgcn_maml_b200.featurePreprocessor import prepare_model_input

    time, nlat, nlon = shape
    w, doy, tod = synth_raw_weather(5 + time, time, nlat, nlon, nan_frac=0.02 if time > 1 else 0.0)
    tf = P.time_features(doy, tod)
    kop = _koppen(torch.randn(31, 8, generator=torch.Generator().manual_seed(3)))
    nan_mask = np.isnan(w).reshape(time, nlat * nlon, 12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, rst = P.prepare_features(w.copy(), tf, kop(torch.tensor([9])), normalize=True, stats=None)
    feats, st = prepare_model_input(_dataset(w, tf), 9, kop, normalize=True, stats=None)
    if time * nlat * nlon == 1:  # one sample: std = 0 + 1e-8, z = 0 exactly on both sides
        assert np.array_equal(feats.cpu().numpy(), ref.detach().numpy())
        return
    _check(str(shape), feats, ref.detach().numpy(), nan_mask, st, rst["mean"], rst["std"], derived=True)


def test_feature_stats_and_roundtrip_at_benchmark_size():
    """632 steps x 441 nodes (configs[1]: ~600 windows): the z-scored channels have mean 0 and unit variance per variable,
    time / Koppen channels do not depend on the node, de-normalising gives the (filled) input back, and the statistics are
    independent of how the rows are split across blocks (two halves combine to the whole)."""
    from oracle.make_golden import synth_raw_weather
    from weatherforecast_stgcn_maml_b200.featurePreprocessor import (assemble_features, denormalize_all_predictions,
                                                                      feature_stats)

    time, n = 632, 441
    w, doy, tod = synth_raw_weather(77, time, 21, 21, nan_frac=0.005)
    tf = P.time_features(doy, tod).astype(np.float32)
    wd = torch.from_numpy(w).reshape(time, n, 12).cuda()
    st = feature_stats(wd)
    assert int(st["nan_count"].sum()) == int(np.isnan(w).sum())
    w64 = np.where(np.isnan(w), st["fill"], w).astype(np.float64).reshape(-1, 12)
    assert np.allclose(st["mean"], w64.mean(0), rtol=1e-12, atol=0) and np.allclose(st["std"], w64.std(0), rtol=1e-10)
    mean, std = st["mean"].astype(np.float32), st["std"].astype(np.float32) + np.float32(1e-8)
    row = np.arange(8, dtype=np.float32) / 7
    f = assemble_features(wd, tf, row, fill=st["fill"], mean=mean, std=std)
    z = f[..., :12].double().reshape(-1, 12)
    unit = torch.from_numpy(st["std"] / (st["std"] + 1e-8)).cuda()  # the reference's epsilon shrinks tiny-variance channels
    assert float(z.mean(0).abs().max()) <= 2e-4 and float((z.std(0, unbiased=False) - unit).abs().max()) <= 1e-5
    assert torch.equal(f[:, :1, 12:].expand(-1, n, -1), f[:, :, 12:])
    assert torch.equal(f[0, 0, 16:].cpu(), torch.from_numpy(row)) and torch.equal(f[:, 0, 12:16].cpu(), torch.from_numpy(tf))
    back = denormalize_all_predictions(f[..., :12].reshape(-1, 12), {"mean": mean, "std": std}).cpu().numpy()
    filled = np.where(np.isnan(w), st["fill"], w).reshape(-1, 12)
    assert np.all(np.abs(back - filled) <= 4e-7 * (np.abs(filled) + np.abs(mean)) + 1e-30)
    # split invariance: statistics of the halves combine to the statistics of the whole
    a, b = feature_stats(wd[: time // 2].contiguous()), feature_stats(wd[time // 2:].contiguous())
    na, nb = (time // 2) * n, (time - time // 2) * n
    va = (a["nan_count"], b["nan_count"])
    sum_valid = lambda s, cnt, nn: s["mean"] * cnt - nn * s["fill"].astype(np.float64)
    tot = (sum_valid(a, na, va[0]) + sum_valid(b, nb, va[1])) / (na + nb - va[0] - va[1])
    assert np.allclose(tot.astype(np.float32), st["fill"], rtol=2e-7)
