"""GPU parity over the dimension d: k_param and the sweep / deconvolution kernels are instantiated per d (k_param for
the exact d = 1..16 with register-resident factorisations, one run-time-d instance for 17..32), so every code path
class is run against the CPU oracle here: labels identical at every iteration, parameters to 1e-7.

Runs last (file name) so that a failure in one exotic dimension cannot hide the core parity tests under `-x`.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _workload(d, q=3, rows=24, cols=25, seed=5, tdist=True):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(seed + d)
    ar, ac = synth.square_lattice(rows, cols)
    truth = synth.voronoi_labels(ar, ac, q, rng)
    Y, _, _ = synth.make_pcs(truth, d, q, rng, tdist=tdist)
    init = synth.noisy_init(truth, q, rng)
    w = synth.Workload(array_row=ar, array_col=ac, n=int(ar.shape[0]), d=d, q=q, gamma=2.0, Y=Y, init=init,
                       platform="ST")
    w.update(synth.priors(Y))
    return w


DIMS = [1, 2, 3, 7, 8, 9, 12, 16, 17, 24, 32]


@pytest.mark.parametrize("fname", ["iterate_t", "iterate_vvv"])
@pytest.mark.parametrize("d", DIMS)
def test_chain_parity_over_dimensions(bsb, oracle, d, fname):
    big = d > 17   # the per-cluster Wishart needs n_k comfortably above d (quirk Q9: an empty cluster is an error)
    w = _workload(d, q=2 if big else 3, rows=30 if big else 24, cols=30 if big else 25, tdist=fname == "iterate_t")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    args = (w.Y, csr, 8, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    got = getattr(bsb, fname)(*args, seed=11, colour=(col, nc))
    ref = getattr(oracle, fname)(*args, seed=11, order=order)
    assert np.array_equal(got["z"], ref["z"]), "labels differ at d = %d" % d
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-7, atol=1e-8)
    assert np.allclose(got["lambda"], ref["lambda"], rtol=1e-6, atol=1e-9)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-9, atol=0)


@pytest.mark.parametrize("d", [1, 5, 16, 17, 32])
def test_deconv_parity_over_dimensions(bsb, oracle, d):
    from bayesspace_b200 import synth
    w = _workload(d, rows=10, cols=12)
    strings, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    pos = np.c_[w.array_col.astype(float), w.array_row.astype(float)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, 1.0, w.init, platform="ST", nsubspots_per_edge=3)
    col, _ = bsb.colour_graph(csr)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], True, 12, 3, dc["n"], dc["n0"], w.d, 2.0, w.q,
            dc["init1"], dc["subspots"], False, 5.0 if d <= 5 else 0.5, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    got = bsb.iterate_deconv(*args, seed=3)
    ref = oracle.iterate_deconv(*args, seed=3, order=order)
    assert np.array_equal(got["z"], ref["z"]), "labels differ at d = %d" % d
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-7, atol=1e-8)
    assert np.allclose(got["Y"], ref["Y"], rtol=1e-8, atol=1e-8)
    assert np.array_equal(got["Ychange"][1:], ref["Ychange"][1:])
