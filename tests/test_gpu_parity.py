"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Levels (BASELINE.json north_star):
  * keyed random variates, neighbour lists, colourings: exact / 1e-14;
  * log-likelihood, sufficient statistics, MH ratios on identical states: <= 1e-10 relative;
  * whole chains: the oracle run with the CUDA path's visiting order (colour classes in sequence) and the
    same Philox key must give the SAME labels at every iteration, and mu / lambda / w / plogLik to 1e-9.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20261020


def _setup(bsb, name, **kw):
    from bayesspace_b200 import synth
    w = synth.make_workload(name, **kw)
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    return w, csr, (col, nc), order


def _args(w, csr, nrep, thin):
    return (w.Y, csr, nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)


def _relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ---- randomness ---------------------------------------------------------------------------------------
def test_philox_known_answers(bsb):
    """Random123 kat vectors for philox4x32-10, evaluated on the device."""
    from bayesspace_b200.samplers import debug_rng
    # counter = (index, slot, iter, purpose), key = seed
    o = debug_rng(3, 0, 0, 0, 0, 0)
    assert [int(x) for x in o] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    o = debug_rng(3, 0xffffffffffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    o = debug_rng(3, (0x299f31d0 << 32) | 0xa4093822, 0x03707344, 0x243f6a88, 0x13198a2e, 0x85a308d3)
    assert [int(x) for x in o] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_keyed_variates_match_oracle(bsb, oracle):
    from bayesspace_b200.samplers import debug_rng
    for (purpose, index, it, slot) in [(4, 17, 3, 0), (7, 123456, 99, 0), (4, 0, 1, 5)]:
        assert np.array_equal(debug_rng(0, SEED, purpose, index, it, slot), oracle.uniform2(SEED, purpose, index, it, slot))
    g = debug_rng(1, SEED, 1, 3, 7, 0, count=64)
    assert _relerr(g, oracle.normal(SEED, 1, 3, 7, 0, 64)) < 1e-14
    for shape, scale in [(9.0, 0.3), (2.0, 2.0), (0.5, 2.0), (2500.5, 2.0)]:
        g = debug_rng(2, SEED, 5, 1000, 4, shape=shape, scale=scale, count=512)
        assert _relerr(g, oracle.gamma(SEED, 5, 1000, 512, 4, shape, scale)) < 1e-13


# ---- state-level probes -------------------------------------------------------------------------------
@pytest.mark.parametrize("d,q,nl", [(4, 3, 1), (10, 4, 4), (15, 7, 1), (20, 5, 5)])
def test_loglik_probe(bsb, oracle, d, q, nl):
    """K3: spot x cluster log-density, covariance form with quirk Q1, without it, and precision form."""
    from bayesspace_b200.samplers import eval_loglik
    rng = np.random.default_rng(d * 100 + q)
    n = 777
    Y = rng.normal(size=(n, d)) * 3
    mu = rng.normal(size=(q, d)) * 2
    mats = []
    for _ in range(nl):
        A = rng.normal(size=(d, d))
        mats.append(A @ A.T / d + 0.5 * np.eye(d))
    mats = np.array(mats)
    w = rng.gamma(2.0, 0.5, size=n)
    for form, compat in [(0, bsb.COMPAT_REFERENCE), (0, 0), (1, bsb.COMPAT_REFERENCE)]:
        for ww in (None, w):
            got = eval_loglik(Y, mu, mats, w=ww, form=form, compat=compat)
            for k in range(q):
                M = mats[k if nl > 1 else 0]
                ref = (oracle.dmvnrm_arma_fast(Y, mu[k], M, w=ww, q1_compat=bool(compat & 1)) if form == 0
                       else oracle.dmvnrm_prec_arma_fast(Y, mu[k], M, w=ww))
                assert np.max(np.abs(got[:, k] - ref) / np.abs(ref)) < 1e-10


@pytest.mark.parametrize("name,nl,weighted", [("small_hex", 1, True), ("small_hex", 4, True), ("small_sq", 1, False),
                                               ("c2", 1, True)])
def test_suffstats_probe(bsb, oracle, name, nl, weighted):
    """K1: n_k, s_k and the scatter about the cluster means from the fused moment sums."""
    from bayesspace_b200 import synth
    from bayesspace_b200.samplers import eval_suffstats
    w = synth.make_workload(name)
    rng = np.random.default_rng(5)
    wts = rng.gamma(2.0, 0.5, size=w.n) if weighted else None
    mu = w.mu_true + rng.normal(size=w.mu_true.shape) * 0.1
    q = w.q if nl == 1 else nl
    z = w.init if nl == 1 else ((w.init - 1) % nl + 1)
    mu = mu[:q]
    nk, sk, S = eval_suffstats(w.Y, z, wts, mu, nl=nl)
    rnk, rsk, rS, rSk = oracle.suffstats(w.Y, z, wts, mu)
    assert _relerr(nk, rnk) < 1e-12
    assert _relerr(sk, rsk) < 1e-11
    if nl == 1:
        assert _relerr(S, rS) < 1e-10
    else:
        for k in range(q):
            assert _relerr(S[k], rSk[k]) < 1e-10


VARIANTS = [("iterate", "normal", "equal"), ("iterate_vvv", "normal", "variable"), ("iterate_t", "t", "equal"),
            ("iterate_t_vvv", "t", "variable")]


@pytest.mark.parametrize("fname,model,precision", VARIANTS)
@pytest.mark.parametrize("name", ["small_hex", "small_sq"])
def test_mh_ratio_probe(bsb, oracle, fname, model, precision, name):
    """The proposal the chain makes at its next iteration and the two values the MH step compares,
    evaluated by the sweep kernel itself (dry run) vs the oracle on the same state."""
    w, csr, colour, order = _setup(bsb, name)
    ch = bsb.Chain(*_args(w, csr, 50, 1), model=model, precision=precision, seed=SEED, colour=colour)
    ch.run(3)
    st = ch.state()
    dry = ch.dry_sweep()
    st2 = ch.state()
    assert np.array_equal(st["z"], st2["z"]) and st2["iters_done"] == 3      # dry run leaves the chain untouched
    # parameters the dry sweep used = those of iteration 4: rerun the oracle to the same point.
    # With thin = 1 iteration i is stored in row i + 1 (row (i+1)/thin, src/cluster.cpp:310-313).
    ref = getattr(oracle, fname)(*_args(w, csr, 6, 1), seed=SEED, order=order)
    mu4 = ref["mu"][5].reshape(w.q, w.d)
    lam4 = ref["lambda"][5]
    sig4 = np.array([np.linalg.inv(l) for l in lam4]) if precision == "variable" else np.linalg.inv(lam4)
    assert np.array_equal(st["z"], ref["z"][4])
    wts = dry["w"] if model == "t" else None
    hp, hn = oracle.h_values(fname, w.Y, csr, st["z"], wts, mu4, sig4, w.gamma, dry["z_new"])
    # relative to the size of the terms (Potts + log-density, both O(1..100)): h itself crosses zero
    assert np.max(np.abs(dry["h_prev"] - hp) / np.maximum(np.abs(hp), 1.0)) < 1e-9
    assert np.max(np.abs(dry["h_new"] - hn) / np.maximum(np.abs(hn), 1.0)) < 1e-9
    pr = np.minimum(1.0, np.exp(hn - hp))
    assert np.max(np.abs(dry["prob"] - pr)) < 1e-9
    ch.close()


# ---- whole chains -------------------------------------------------------------------------------------
@pytest.mark.parametrize("fname,model,precision", VARIANTS)
@pytest.mark.parametrize("name", ["tiny_hex", "small_hex", "small_sq"])
def test_chain_matches_oracle_chromatic_scan(bsb, oracle, fname, model, precision, name):
    w, csr, colour, order = _setup(bsb, name)
    nrep = 40
    got = getattr(bsb, fname)(*_args(w, csr, nrep, 1), seed=SEED, colour=colour)
    ref = getattr(oracle, fname)(*_args(w, csr, nrep, 1), seed=SEED, order=order)
    assert got["z"].shape == ref["z"].shape == (nrep + 1, w.n)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["lambda"], ref["lambda"], rtol=1e-8, atol=1e-10)
    if model == "t":
        assert np.allclose(got["weights"], ref["weights"], rtol=1e-9, atol=1e-12)
    assert np.isnan(got["plogLik"][0]) and np.isnan(ref["plogLik"][0])
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10, atol=0)


def test_chain_thinning_and_shapes(bsb, oracle):
    """Output containers as App. D: rows nrep/thin + 1, row 0 = initial state, thin boundaries."""
    w, csr, colour, order = _setup(bsb, "small_hex")
    got = bsb.iterate_t(*_args(w, csr, 57, 10), seed=3, colour=colour)
    ref = oracle.iterate_t(*_args(w, csr, 57, 10), seed=3, order=order)
    assert got["z"].shape == (6, w.n) and got["mu"].shape == (6, w.q * w.d) and got["lambda"].shape == (6, w.d, w.d)
    assert np.array_equal(got["z"][0], w.init) and np.all(got["weights"][0] == 1.0)
    assert np.allclose(got["mu"][0], np.tile(w.mu0, w.q)) and np.allclose(got["lambda"][0], w.lambda0)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert got["plogLik"].shape == (57,)


def test_chain_c2_visium(bsb, oracle):
    """BASELINE config C2 geometry (4,992-spot hex lattice, q=7, d=15, t model), short chain."""
    w, csr, colour, order = _setup(bsb, "c2")
    nrep = 12
    got = bsb.iterate_t(*_args(w, csr, nrep, 1), seed=SEED, colour=colour)
    ref = oracle.iterate_t(*_args(w, csr, nrep, 1), seed=SEED, order=order)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10)


def test_greedy_colouring_chain(bsb, oracle):
    """No colouring supplied: the library colours the graph itself (greedy); the oracle follows that order."""
    w, csr, _, _ = _setup(bsb, "small_hex")
    col, nc = bsb.colour_graph(csr)
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    got = bsb.iterate_t(*_args(w, csr, 20, 1), seed=5)
    ref = oracle.iterate_t(*_args(w, csr, 20, 1), seed=5, order=order)
    assert got["_info"]["ncolours"] == nc
    assert np.array_equal(got["z"], ref["z"])


def test_large_lattice_grouped_reduction(bsb, oracle):
    """n >= 262,144 switches on the 64-group two-level reduction; chain must still match the oracle."""
    w, csr, colour, order = _setup(bsb, "small_sq", n_override=(520, 512))
    nrep = 4
    got = bsb.iterate(*_args(w, csr, nrep, 1), seed=SEED, colour=colour)
    ref = oracle.iterate(*_args(w, csr, nrep, 1), seed=SEED, order=order)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-9)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-10)


def test_chain_statistical_equivalence_with_reference_scan(bsb, oracle):
    """Chromatic scan (CUDA) vs the reference's sequential scan (oracle, identity order): same target,
    different chains.  MAP labels agree with each other and with the truth; posterior means of mu agree
    within Monte-Carlo error (north_star, third level)."""
    from sklearn.metrics import adjusted_rand_score as ari
    w, csr, colour, _ = _setup(bsb, "small_hex")
    nrep, burn = 1500, 300
    got = bsb.iterate_t(*_args(w, csr, nrep, 1), seed=1, colour=colour)
    ref = oracle.iterate_t(*_args(w, csr, nrep, 1), seed=2)

    def mode(zs):
        return np.array([np.bincount(zs[:, j], minlength=w.q + 1).argmax() for j in range(zs.shape[1])])

    zg, zr = mode(got["z"][burn:]), mode(ref["z"][burn:])
    assert ari(zg, zr) >= 0.95 and ari(zg, w.truth) >= 0.95
    mg, mr = got["mu"][burn:], ref["mu"][burn:]

    def mcse(x):   # batch means
        b = x[: (x.shape[0] // 20) * 20].reshape(20, -1, x.shape[1]).mean(axis=1)
        return b.std(axis=0, ddof=1) / np.sqrt(20)

    zscore = np.abs(mg.mean(axis=0) - mr.mean(axis=0)) / np.sqrt(mcse(mg) ** 2 + mcse(mr) ** 2)
    assert np.mean(zscore < 3) > 0.9 and zscore.max() < 6
    assert abs(got["plogLik"][burn:].mean() - ref["plogLik"][burn:].mean()) < 0.01 * abs(ref["plogLik"][burn:].mean())


def test_errors(bsb):
    w, csr, colour, _ = _setup(bsb, "tiny_hex")
    bad = w.init.copy()
    bad[0] = w.q + 1
    with pytest.raises(bsb.BayesSpaceError):
        bsb.iterate(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, bad, w.mu0, w.lambda0, w.alpha, w.beta)
    with pytest.raises(bsb.BayesSpaceError, match="mu0"):
        bsb.iterate(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0[:-1], w.lambda0, w.alpha, w.beta)
    # empty cluster in a VVV sampler: Wishart df <= 0 (quirk Q9) -> explicit numeric error
    init = np.where(w.init == w.q, 1, w.init).astype(np.int32)
    with pytest.raises(bsb.BayesSpaceError) as ei:
        bsb.iterate_vvv(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, init, w.mu0, w.lambda0, w.alpha, w.beta)
    assert ei.value.code == -6
