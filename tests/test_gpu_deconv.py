"""GPU parity of bsb200_deconv (iterate_deconv, src/cluster.cpp:742-1141) against the CPU oracle.

The oracle visits parent spots colour class by colour class (the order of the CUDA path) and, inside a
parent, subspots r = 0..S-1 as the reference does; with the same Philox key the two chains must agree:
labels exactly, Y / w / mu / lambda / plogLik / Ychange to rounding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(bsb, name, platform_k=None):
    from bayesspace_b200 import synth
    w = synth.make_workload(name)
    strings, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    if w.platform == "Visium":
        pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
        dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium")
    else:
        pos = np.c_[w.array_col.astype(float), w.array_row.astype(float)]
        dc = synth.deconv_inputs(w.Y, pos, 1.0, 1.0, w.truth, platform=w.platform, nsubspots_per_edge=platform_k or 3)
    col, _ = bsb.colour_graph(csr)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    return w, dc, strings, order


def _run_both(bsb, oracle, w, dc, strings, order, tdist, jitter_scale, nrep, thin, seed=77, adapt_before=0):
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], tdist, nrep, thin, dc["n"], dc["n0"], w.d, 2.0, w.q,
            dc["init1"], dc["subspots"], False, jitter_scale, adapt_before, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    got = bsb.iterate_deconv(*args, seed=seed, mean_from=1)
    ref = oracle.iterate_deconv(*args, seed=seed, order=order)
    return got, ref


def _compare(got, ref, tdist):
    assert np.array_equal(got["df_j"], ref["df_j"])
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["Y"], ref["Y"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["lambda"], ref["lambda"], rtol=1e-8, atol=1e-10)
    assert np.allclose(got["weights"], ref["weights"], rtol=1e-9, atol=1e-12)
    assert np.isnan(got["Ychange"][0]) and np.array_equal(got["Ychange"][1:], ref["Ychange"][1:])
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10)
    if not tdist:
        assert np.all(got["weights"] == 1.0)


@pytest.mark.parametrize("tdist", [False, True])
@pytest.mark.parametrize("name,k", [("tiny_hex", None), ("small_hex", None), ("small_sq", 3), ("small_sq", 2)])
def test_deconv_fixed_jitter_matches_oracle(bsb, oracle, name, k, tdist):
    w, dc, strings, order = _case(bsb, name, k)
    got, ref = _run_both(bsb, oracle, w, dc, strings, order, tdist, 5.0, 45, 5)
    _compare(got, ref, tdist)
    # running mean of the Y snapshots (what spatialEnhance averages, R/spatialEnhance.R:441)
    assert np.allclose(got["Y_mean"], ref["Y"][1:].mean(axis=0), rtol=1e-9, atol=1e-9)
    # the subspot mean of every parent never moves (zero-sum jitter)
    S, n0 = dc["subspots"], dc["n0"]
    assert np.allclose(got["Y"][-1].reshape(S, n0, w.d).mean(axis=0), w.Y, atol=1e-9)


@pytest.mark.parametrize("adapt_before", [0, 20])
def test_deconv_adaptive_matches_oracle(bsb, oracle, adapt_before):
    """jitter_scale == 0 switches on Robust Adaptive Metropolis (src/cluster.cpp:49-65, :987-995); the CUDA
    path applies the update as a rank-one Cholesky up/down-date, the oracle as the reference's chol(...)."""
    w, dc, strings, order = _case(bsb, "tiny_hex")
    got, ref = _run_both(bsb, oracle, w, dc, strings, order, True, 0.0, 40, 4, adapt_before=adapt_before)
    _compare(got, ref, True)


def test_deconv_c3_geometry_short(bsb, oracle):
    """BASELINE config C3 geometry: 4,992 parents x 6 subspots = 29,952 subspots, d = 15, q = 7, t model."""
    w, dc, strings, order = _case(bsb, "c2")
    got, ref = _run_both(bsb, oracle, w, dc, strings, order, True, 5.0, 6, 1)
    _compare(got, ref, True)


def test_deconv_errors(bsb):
    w, dc, strings, order = _case(bsb, "tiny_hex")
    with pytest.raises(bsb.BayesSpaceError, match="finding neighbors of subspots"):
        bsb.iterate_deconv(dc["positions2"] * 50, dc["dist"], strings, dc["Y2"], True, 10, 1, dc["n"], dc["n0"], w.d,
                           2.0, w.q, dc["init1"], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0, 1.0, 0.01)
    with pytest.raises(bsb.BayesSpaceError, match="Invalid arguments"):
        bsb.iterate_deconv(dc["positions2"][:-1], dc["dist"], strings, dc["Y2"][:-1], True, 10, 1, dc["n"] - 1,
                           dc["n0"], w.d, 2.0, w.q, dc["init1"][:-1], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0,
                           w.lambda0, 1.0, 0.01)
