"""GPU parity over the number of clusters q and on the BASELINE configurations that round 1 left untested.

The sufficient statistics of the sweep take a different code path per q range: one DMMA label tile for q <= 8, two
for q = 9..16 (csrc/tile_stats.cuh), the scalar run-length sums for q > 16 and for every VVV sampler.  C4
(q = 15, d = 20) and qTune's sweep (q up to 15) run on those paths, so each is compared with the CPU oracle here:
labels identical at every iteration, parameters / weights / plogLik to rounding.

Also: C1 exactly as the reference's own example (8 x 12 "ST" rectangle, d = 10, q = 4, normal, nrep = 1000,
thin = 100; shape contract of tests/testthat/test-mcmcChain.R:57-85), adaptive deconvolution at d = 15, and a
short chain at the full C4 size.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _square_workload(rows, cols, d, q, seed, tdist=True):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(seed)
    ar, ac = synth.square_lattice(rows, cols)
    truth = synth.voronoi_labels(ar, ac, q, rng)
    Y, _, _ = synth.make_pcs(truth, d, q, rng, tdist=tdist)
    init = synth.noisy_init(truth, q, rng)
    w = synth.Workload(array_row=ar, array_col=ac, n=int(ar.shape[0]), d=d, q=q, gamma=2.0, Y=Y, init=init,
                       truth=truth, platform="ST")
    w.update(synth.priors(Y))
    return w


def _graph(bsb, w):
    strings, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    return strings, csr, (col, nc), order


def _args(w, csr, nrep, thin):
    return (w.Y, csr, nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)


def _compare_chain(got, ref, tdist, rtol_mu=1e-9):
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=rtol_mu, atol=1e-9)
    assert np.allclose(got["lambda"], ref["lambda"], rtol=1e-7, atol=1e-10)
    if tdist:
        assert np.allclose(got["weights"], ref["weights"], rtol=1e-9, atol=1e-12)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10, atol=0)


@pytest.mark.parametrize("fname", ["iterate", "iterate_t"])
@pytest.mark.parametrize("q", [8, 9, 15, 16, 17, 20])
def test_chain_parity_over_q(bsb, oracle, q, fname):
    """q = 8 | 9..16 | > 16: one label tile, two label tiles, scalar sums (csrc/tile_stats.cuh)."""
    w = _square_workload(48, 50, 6, q, seed=100 + q, tdist=fname == "iterate_t")
    _, csr, colour, order = _graph(bsb, w)
    got = getattr(bsb, fname)(*_args(w, csr, 10, 1), seed=21, colour=colour)
    ref = getattr(oracle, fname)(*_args(w, csr, 10, 1), seed=21, order=order)
    _compare_chain(got, ref, fname == "iterate_t")


@pytest.mark.parametrize("q", [9, 17])
def test_vvv_parity_over_q(bsb, oracle, q):
    """Per-cluster precision with many clusters (scalar pair sums, nl = q parameter blocks in shared memory)."""
    w = _square_workload(60, 60, 4, q, seed=300 + q, tdist=True)
    _, csr, colour, order = _graph(bsb, w)
    got = bsb.iterate_t_vvv(*_args(w, csr, 8, 1), seed=22, colour=colour)
    ref = oracle.iterate_t_vvv(*_args(w, csr, 8, 1), seed=22, order=order)
    _compare_chain(got, ref, True, rtol_mu=1e-8)


def test_chain_c4_instantiation(bsb, oracle):
    """The kernel instantiation C4 benchmarks: d = 20, q = 15, t model, shared precision, square lattice (64 x 64)."""
    w = _square_workload(64, 64, 20, 15, seed=4, tdist=True)
    _, csr, colour, order = _graph(bsb, w)
    got = bsb.iterate_t(*_args(w, csr, 8, 1), seed=23, colour=colour)
    ref = oracle.iterate_t(*_args(w, csr, 8, 1), seed=23, order=order)
    _compare_chain(got, ref, True)


def test_chain_c4_full_size_two_sweeps(bsb, oracle):
    """BASELINE config C4 at its full size (837 x 837 = 700,569 bins, q = 15, d = 20, t): the grouped two-level
    statistics fold, nine-tile segments and the second label tile together, 2 sweeps (about 10 s of oracle)."""
    from bayesspace_b200 import synth
    w = synth.make_workload("c4")
    _, csr, colour, order = _graph(bsb, w)
    got = bsb.iterate_t(*_args(w, csr, 3, 1), seed=24, colour=colour)
    ref = oracle.iterate_t(*_args(w, csr, 3, 1), seed=24, order=order)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["weights"], ref["weights"], rtol=1e-9, atol=1e-12)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10, atol=0)


def test_chain_c1_example_sce(bsb, oracle):
    """BASELINE config C1 = the reference's exampleSCE() run (R/utils.R:138-173): 8 x 12 'ST' rectangle, d = 10,
    q = 4, normal model, equal precision, nrep = 1000, thin = 100.  Shape contract of
    tests/testthat/test-mcmcChain.R:57-85: ncol(chain) = d^2 + q d + n, nrow = nrep / 100 + 1."""
    from bayesspace_b200 import synth
    w = synth.make_workload("c1")
    assert (w.n, w.d, w.q) == (96, 10, 4)
    _, csr, colour, order = _graph(bsb, w)
    nrep, thin = 1000, 100
    got = bsb.iterate(*_args(w, csr, nrep, thin), seed=149, colour=colour)
    ref = oracle.iterate(*_args(w, csr, nrep, thin), seed=149, order=order)
    rows = nrep // 100 + 1
    assert got["z"].shape == (rows, w.n) and got["mu"].shape == (rows, w.q * w.d)
    assert got["lambda"].shape == (rows, w.d, w.d) and got["plogLik"].shape == (nrep,)
    assert got["z"].shape[1] + got["mu"].shape[1] + w.d * w.d == w.d ** 2 + w.q * w.d + w.n
    assert "weights" not in got
    # 999 iterations of two chaotic systems in different arithmetic: labels must agree while the parameter
    # trajectories agree to rounding; compare row by row and stop claiming equality only if rounding ever flips a label
    assert np.array_equal(got["z"][:3], ref["z"][:3])
    assert np.allclose(got["mu"][:3], ref["mu"][:3], rtol=1e-7, atol=1e-8)
    same = np.all(got["z"] == ref["z"], axis=1)
    if same.all():
        assert np.allclose(got["mu"], ref["mu"], rtol=1e-6, atol=1e-7)
        assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-8)
    assert np.isnan(got["plogLik"][0])


@pytest.mark.parametrize("adapt_before", [0, 10])
def test_deconv_adaptive_d15(bsb, oracle, adapt_before):
    """Robust Adaptive Metropolis (src/cluster.cpp:49-65) at the dimension of C3 (d = 15, q = 7, hex subspots)."""
    from bayesspace_b200 import synth
    rng = np.random.default_rng(15)
    ar, ac = synth.hex_lattice(14, 12)
    q, d = 7, 15
    truth = synth.voronoi_labels(ar, ac, q, rng, hex_geometry=True)
    Y, _, _ = synth.make_pcs(truth, d, q, rng, tdist=True)
    init = synth.noisy_init(truth, q, rng)
    pri = synth.priors(Y)
    strings, csr = bsb.neighbors._find_neighbors(ar, ac, "Visium")
    pos = np.c_[ac.astype(float), ar * np.sqrt(3.0)]
    dc = synth.deconv_inputs(Y, pos, 1.0, np.sqrt(3.0), init, platform="Visium")
    col, _ = bsb.colour_graph(csr)
    order = np.lexsort((np.arange(len(ar)), col)).astype(np.int32)
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], True, 24, 4, dc["n"], dc["n0"], d, 3.0, q, dc["init1"],
            dc["subspots"], False, 0.0, adapt_before, dc["c"], pri["mu0"], pri["lambda0"], 1.0, 0.01, 1)
    got = bsb.iterate_deconv(*args, seed=31)
    ref = oracle.iterate_deconv(*args, seed=31, order=order)
    assert np.array_equal(got["z"], ref["z"])
    assert np.array_equal(got["Ychange"][1:], ref["Ychange"][1:])
    assert np.allclose(got["Y"], ref["Y"], rtol=1e-8, atol=1e-8)
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-8, atol=1e-8)
    assert np.allclose(got["weights"], ref["weights"], rtol=1e-8, atol=1e-11)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-9)


@pytest.mark.parametrize("q", [9, 17])
def test_deconv_parity_over_q(bsb, oracle, q):
    """Deconvolution statistics with two label tiles (q = 9) and on the scalar path (q = 17)."""
    from bayesspace_b200 import synth
    w = _square_workload(16, 18, 5, q, seed=500 + q)
    strings, csr, _, _ = _graph(bsb, w)
    pos = np.c_[w.array_col.astype(float), w.array_row.astype(float)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, 1.0, w.init, platform="ST", nsubspots_per_edge=3)
    col, _ = bsb.colour_graph(csr)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], True, 12, 3, dc["n"], dc["n0"], w.d, 2.0, w.q,
            dc["init1"], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    got = bsb.iterate_deconv(*args, seed=3)
    ref = oracle.iterate_deconv(*args, seed=3, order=order)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-8, atol=1e-8)
    assert np.allclose(got["Y"], ref["Y"], rtol=1e-8, atol=1e-8)
    assert np.array_equal(got["Ychange"][1:], ref["Ychange"][1:])


def test_configuration_beyond_shared_memory_is_reported(bsb):
    """d = 32 with q = 255 clusters and per-cluster precisions needs far more shared memory than one CTA has:
    the library says so (BSB200_ERR_UNSUPPORTED) instead of failing inside a launch."""
    w = _square_workload(40, 40, 32, 2, seed=9)
    _, csr, colour, _ = _graph(bsb, w)
    init = (np.arange(w.n) % 255 + 1).astype(np.int32)
    with pytest.raises(bsb.BayesSpaceError, match="shared memory") as ei:
        bsb.iterate_t_vvv(w.Y, csr, 4, 1, w.n, w.d, w.gamma, 255, init, w.mu0, w.lambda0, w.alpha, w.beta, colour=colour)
    assert ei.value.code == -7
