"""CPU tests of the oracle (oracle/ref_cpu.cpp): known answers, independent numpy/scipy restatements of
its building blocks, the reference's own neighbour test, output containers, and the committed golden
vectors.  No GPU needed."""
import os

import numpy as np
import pytest
from scipy import stats

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_philox_random123_known_answers(oracle):
    assert [int(x) for x in oracle.philox([0] * 4, [0] * 2)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_variate_distributions(oracle):
    u = np.concatenate([oracle.uniform2(3, 4, i, 1, 0) for i in range(4000)])
    assert 0 < u.min() and u.max() < 1 and stats.kstest(u, "uniform").pvalue > 1e-3
    x = oracle.normal(5, 1, 0, 2, 0, 40000)
    assert stats.kstest(x, "norm").pvalue > 1e-3
    for shape, scale in [(9.0, 0.25), (2.0, 2.0), (0.5, 2.0), (60.5, 2.0)]:
        g = oracle.gamma(7, 5, 0, 40000, 3, shape, scale)
        assert stats.kstest(g, "gamma", args=(shape, 0, scale)).pvalue > 1e-3
    assert np.isnan(oracle.gamma(7, 5, 0, 1, 3, 0.0, 2.0)[0])   # rchisq(df <= 0) is NaN in R (quirk Q9)


def test_dense_helpers_against_numpy(oracle):
    rng = np.random.default_rng(1)
    for d in (1, 3, 15, 30):
        A = rng.normal(size=(d, d))
        X = A @ A.T + d * np.eye(d)
        R = oracle.chol_upper(X)
        assert np.allclose(R, np.linalg.cholesky(X).T, rtol=1e-12, atol=1e-12)
        assert np.allclose(oracle.inv(X), np.linalg.inv(X), rtol=1e-10, atol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        oracle.chol_upper(np.array([[1.0, 2.0], [2.0, 1.0]]))


def test_density_forms(oracle):
    """dmvnrm_arma_fast carries quirk Q1: quadratic form (x-m)^T (R R^T)^-1 (x-m), R = chol_upper(Sigma);
    the log-determinant part is the true one.  The precision form is the true Gaussian."""
    rng = np.random.default_rng(0)
    for d in (2, 5, 15):
        A = rng.normal(size=(d, d))
        Sig = A @ A.T / d + 0.5 * np.eye(d)
        x, m = rng.normal(size=(7, d)) * 2, rng.normal(size=d)
        R = np.linalg.cholesky(Sig).T
        Mq1 = np.linalg.inv(R @ R.T)
        const = -np.log(np.diag(R)).sum() - d / 2 * np.log(2 * np.pi)
        ref_q1 = np.array([const - 0.5 * (xi - m) @ Mq1 @ (xi - m) for xi in x])
        true = stats.multivariate_normal.logpdf(x, m, Sig)
        assert np.allclose(oracle.dmvnrm_arma_fast(x, m, Sig), ref_q1, rtol=1e-12)
        assert np.allclose(oracle.dmvnrm_arma_fast(x, m, Sig, q1_compat=False), true, rtol=1e-12)
        assert np.allclose(oracle.dmvnrm_prec_arma_fast(x, m, np.linalg.inv(Sig)), true, rtol=1e-10)
        w = rng.gamma(2, 0.5, size=7)
        tw = np.array([stats.multivariate_normal.logpdf(x[i], m, Sig / w[i]) for i in range(7)])
        assert np.allclose(oracle.dmvnrm_arma_fast(x, m, Sig, w=w, q1_compat=False), tw, rtol=1e-12)
        assert np.allclose(oracle.dmvnrm_prec_arma_fast(x, m, np.linalg.inv(Sig), w=w), tw, rtol=1e-10)
        if d > 1:   # Q1 only vanishes for diagonal covariances
            assert not np.allclose(ref_q1, true)
        Dg = np.diag(rng.uniform(0.5, 2, d))
        assert np.allclose(oracle.dmvnrm_arma_fast(x, m, Dg), stats.multivariate_normal.logpdf(x, m, Dg), rtol=1e-12)


def test_reference_find_neighbors_known_answer(oracle):
    """tests/testthat/test-find_neighbors.R:1-12 -- the only known answer the reference holds on this path."""
    positions = np.c_[np.arange(1, 6), np.arange(5, 0, -1)]
    df_j = oracle.find_neighbors(positions, 1, "manhattan")
    assert all(len(v) == 0 for v in df_j)
    df_j = oracle.find_neighbors(positions, 4, "manhattan")
    assert list(df_j[0]) == [1, 2] and list(df_j[2]) == [0, 1, 3, 4]
    df_e = oracle.find_neighbors(positions, 1.5, "euclidean")   # 0 < dist <= radius, sqrt(2) apart
    assert list(df_e[2]) == [1, 3]


def test_lattice_graph_facts(oracle):
    """Visium 78 x 64 hex lattice and its S = 6 subspot graph."""
    from bayesspace_b200 import synth
    r, c = synth.hex_lattice(78, 64)
    ptr, idx = oracle.lattice_neighbors(r, c, "Visium")
    deg = np.diff(ptr)
    assert ptr[-1] == 2 * 14693
    assert {int(k): int(v) for k, v in zip(*np.unique(deg, return_counts=True))} == {2: 2, 3: 78, 4: 124, 5: 76, 6: 4712}
    pos = np.c_[c.astype(float), r * np.sqrt(3.0)]
    dc = synth.deconv_inputs(np.zeros((len(r), 2)) + np.arange(len(r))[:, None], pos, 1.0, np.sqrt(3.0),
                             np.ones(len(r), dtype=np.int32))
    sp, si = oracle.find_subspot_neighbors((ptr, idx), dc["positions2"], dc["dist"])
    sdeg = np.diff(sp)
    assert {int(k): int(v) for k, v in zip(*np.unique(sdeg, return_counts=True))} == {2: 566, 3: 29386}
    nb = [set(si[sp[s]:sp[s + 1]]) for s in range(len(sp) - 1)]
    assert all(s in nb[u] for s in range(len(nb)) for u in nb[s])   # symmetric


def test_parse_neighbor_strings(oracle):
    ptr, idx = oracle.parse_neighbors(["2,3", None, "1", "1,,2", ""])
    assert list(ptr) == [0, 2, 2, 3, 5, 5] and list(idx[:5]) == [1, 2, 0, 0, 1]
    with pytest.raises(ValueError):
        oracle.parse_neighbors(["1,x"])


def test_subspot_errors(oracle):
    ptr = np.array([0, 1, 2], dtype=np.int64)
    idx = np.array([1, 0], dtype=np.int32)
    with pytest.raises(RuntimeError, match="Invalid arguments"):
        oracle.find_subspot_neighbors((ptr, idx), np.zeros((5, 2)), 1.0)
    with pytest.raises(RuntimeError, match="finding neighbors of subspots"):
        oracle.find_subspot_neighbors((ptr, idx), np.arange(8.0).reshape(4, 2) * 10, 1.0)


@pytest.mark.parametrize("fname,tdist,vvv", [("iterate", False, False), ("iterate_vvv", False, True),
                                             ("iterate_t", True, False), ("iterate_t_vvv", True, True)])
def test_output_containers(oracle, fname, tdist, vvv):
    """tests/testthat/test-mcmcChain.R:57-85 (shape contract) + row 0 = initial state + NA first plogLik."""
    from bayesspace_b200 import synth
    # exampleSCE geometry (8 x 12 rectangle, platform "ST", d = 10, q = 4); the per-cluster Wishart of the
    # VVV samplers needs n_k + alpha > d - 1 in every cluster (quirk Q9), hence a larger lattice there
    w = synth.make_workload("small_sq" if vvv else "c1")
    ptr, idx = oracle.lattice_neighbors(w.array_row, w.array_col, "ST")
    nrep, thin = 200, 100
    out = getattr(oracle, fname)(w.Y, (ptr, idx), nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0,
                                 w.alpha, w.beta, seed=149)
    rows = nrep // thin + 1
    assert out["z"].shape == (rows, w.n) and out["mu"].shape == (rows, w.q * w.d)
    assert out["lambda"].shape == ((rows, w.q, w.d, w.d) if vvv else (rows, w.d, w.d))
    assert ("weights" in out) == tdist
    assert np.array_equal(out["z"][0], w.init) and np.allclose(out["mu"][0], np.tile(w.mu0, w.q))
    assert np.isnan(out["plogLik"][0]) and np.all(np.isfinite(out["plogLik"][1:]))
    assert out["z"].min() >= 1 and out["z"].max() <= w.q
    lam = out["lambda"][-1]
    assert np.allclose(lam, np.swapaxes(lam, -1, -2), rtol=1e-10)


def test_visiting_order_only_permutes_the_scan(oracle):
    """order = identity must be the same chain as order = None (the reference's scan)."""
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    g = oracle.lattice_neighbors(w.array_row, w.array_col, "Visium")
    a = (w.Y, g, 30, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    o1 = oracle.iterate_t(*a, seed=4)
    o2 = oracle.iterate_t(*a, seed=4, order=np.arange(w.n, dtype=np.int32))
    assert np.array_equal(o1["z"], o2["z"]) and np.array_equal(o1["mu"], o2["mu"])
    o3 = oracle.iterate_t(*a, seed=4, order=np.arange(w.n, dtype=np.int32)[::-1].copy())
    assert not np.array_equal(o1["z"], o3["z"])


def test_sampler_recovers_domains(oracle):
    from sklearn.metrics import adjusted_rand_score as ari
    from bayesspace_b200 import synth
    w = synth.make_workload("small_hex")
    g = oracle.lattice_neighbors(w.array_row, w.array_col, "Visium")
    out = oracle.iterate_t(w.Y, g, 300, 10, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta, seed=1)
    zs = out["z"][10:]
    mode = np.array([np.bincount(zs[:, j]).argmax() for j in range(w.n)])
    assert ari(mode, w.truth) > 0.95
    mu = out["mu"][10:].mean(axis=0).reshape(w.q, w.d)
    emp = np.array([w.Y[w.truth == k + 1].mean(axis=0) for k in range(w.q)])   # labels were initialised near truth
    assert np.abs(mu - emp).max() < 0.5


def test_adaptive_mcmc_is_a_rank_one_update(oracle):
    rng = np.random.default_rng(3)
    d = 6
    A = np.tril(rng.normal(size=(d, d))) + 3 * np.eye(d)
    u = rng.normal(size=d)
    for acc in (0.05, 0.234, 0.9):
        out = oracle.adaptive_mcmc(acc, 0.234, 20, A, u)
        eta = min(1.0, d * 20 ** (-2.0 / 3))
        M = A @ (np.eye(d) + eta * (acc - 0.234) * np.outer(u, u) / (u @ u)) @ A.T
        assert np.allclose(out @ out.T, M, rtol=1e-12) and np.allclose(out, np.tril(out))


def _deconv_case(oracle, tdist, jitter_scale, nrep=60, seed=5):
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    ptr, idx = oracle.lattice_neighbors(w.array_row, w.array_col, "Visium")
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth)
    strings = synth.neighbor_strings(ptr, idx)
    out = oracle.iterate_deconv(dc["positions2"], dc["dist"], strings, dc["Y2"], tdist, nrep, 10, dc["n"], dc["n0"], w.d,
                                2.0, w.q, dc["init1"], dc["subspots"], False, jitter_scale, 30, dc["c"], w.mu0,
                                w.lambda0, w.alpha, w.beta, 2, seed=seed)
    return w, dc, out


@pytest.mark.parametrize("tdist,jitter_scale", [(False, 5.0), (True, 5.0), (True, 0.0)])
def test_deconv_invariants(oracle, tdist, jitter_scale):
    w, dc, out = _deconv_case(oracle, tdist, jitter_scale)
    n, n0, S = dc["n"], dc["n0"], dc["subspots"]
    rows = 60 // 10 + 1
    assert out["z"].shape == (rows, n) and out["Y"].shape == (rows, n, w.d) and out["df_j"].shape == (n, 4)
    assert np.array_equal(out["Y"][0], dc["Y2"])
    # the jitter is centred per parent, so the mean over a parent's subspots never moves (:932-935)
    for r in (1, rows - 1):
        m = out["Y"][r].reshape(S, n0, w.d).mean(axis=0)
        assert np.allclose(m, w.Y, atol=1e-10)
    assert np.isnan(out["Ychange"][0]) and np.all((out["Ychange"][1:] >= 0) & (out["Ychange"][1:] <= 1))
    assert out["Ychange"][1:].mean() > 0.005
    assert np.all(out["weights"][-1] == 1.0) if not tdist else np.all(out["weights"][-1] > 0)
    assert ((out["df_j"] > 0).sum(axis=1) >= 2).all() and out["df_j"].max() <= n


def test_golden_vectors(oracle):
    """Committed outputs of the oracle (tests/golden/make_golden.py): pins the oracle against drift.
    They are NOT outputs of the reference (which cannot run here): parity stays 'unpinned' w.r.t. it."""
    import json
    from bayesspace_b200 import synth
    meta = json.load(open(os.path.join(GOLD, "golden_meta.json")))
    gold = np.load(os.path.join(GOLD, "golden_chains.npz"))
    for case in meta["cases"]:
        w = synth.make_workload(case["workload"])
        g = oracle.lattice_neighbors(w.array_row, w.array_col, w.platform)
        order = gold[case["key"] + "_order"]
        out = getattr(oracle, case["sampler"])(w.Y, g, case["nrep"], case["thin"], w.n, w.d, w.gamma, w.q, w.init,
                                               w.mu0, w.lambda0, w.alpha, w.beta, seed=case["seed"], order=order)
        assert np.array_equal(out["z"], gold[case["key"] + "_z"])
        assert np.allclose(out["mu"], gold[case["key"] + "_mu"], rtol=1e-12, atol=1e-12)
        assert np.allclose(out["plogLik"][1:], gold[case["key"] + "_plogLik"][1:], rtol=1e-12)


def test_mode_rows_is_R_Mode():
    """R/utils.R:63-66: most frequent value, ties to the earliest first appearance."""
    from oracle import oracle as O
    zs = np.array([[2, 1, 3, 1],
                   [1, 1, 3, 2],
                   [2, 3, 1, 2],
                   [1, 3, 1, 3]])
    # col 0: 2,1,2,1 -> tie, 2 appears first; col 1: 1,1,3,3 -> 1; col 2: 3,3,1,1 -> 3; col 3: 1,2,2,3 -> 2
    assert O.mode_rows(zs).tolist() == [2, 1, 3, 2]
    assert O.mode_rows(zs[:1]).tolist() == [2, 1, 3, 1]
    rng = np.random.default_rng(0)
    big = rng.integers(1, 5, size=(9, 200))
    got = O.mode_rows(big)
    for j in range(200):
        c = np.bincount(big[:, j], minlength=5)
        assert c[got[j]] == c.max()
        first = {v: int(np.argmax(big[:, j] == v)) for v in range(1, 5) if c[v] == c.max()}
        assert got[j] == min(first, key=first.get)


@pytest.mark.parametrize("variant", ["iterate", "iterate_vvv", "iterate_t", "iterate_t_vvv"])
@pytest.mark.parametrize("name", ["tiny_hex", "small_sq"])
def test_oracle_agrees_with_an_independent_numpy_restatement(oracle, variant, name):
    """Two restatements of the reference written independently from its lines -- oracle/ref_cpu.cpp (C++) and
    tests/spec_numpy.py (numpy.linalg, the reference's own operation order) -- fed the same keyed variates must
    produce the same chain in the reference's sequential scan: labels identical, parameters to rounding."""
    import spec_numpy
    from bayesspace_b200 import synth
    w = synth.make_workload(name)
    ptr, idx = oracle.lattice_neighbors(w.array_row, w.array_col, w.platform)
    nbrs = [idx[ptr[j]:ptr[j + 1]] for j in range(w.n)]
    nrep, thin, seed = 7, 2, 31
    Y = np.ascontiguousarray(w.Y)
    ref = getattr(oracle, variant)(w.Y, (ptr, idx), nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0,
                                   w.alpha, w.beta, seed=seed)
    got = spec_numpy.run(oracle, variant, Y, nbrs, nrep, thin, w.n, w.d, w.gamma, w.q, w.init, np.asarray(w.mu0),
                         np.asarray(w.lambda0), w.alpha, w.beta, seed)
    assert np.array_equal(got["z"], ref["z"])
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-10)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-11, atol=0)
    if variant.startswith("iterate_t"):
        assert np.allclose(got["weights"], ref["weights"], rtol=1e-10, atol=0)
    for r in range(nrep // thin + 1):
        lam_ref = np.asarray(ref["lambda"][r])
        lam_got = np.asarray(got["lambda"][r])
        assert lam_got.shape == lam_ref.shape
        assert np.allclose(lam_got, lam_ref, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("tdist,jitter_scale,adapt_before", [(True, 5.0, 0), (False, 5.0, 0), (True, 0.0, 0),
                                                              (True, 0.0, 14)])
def test_oracle_deconv_agrees_with_an_independent_numpy_restatement(oracle, tdist, jitter_scale, adapt_before):
    """iterate_deconv (src/cluster.cpp:742-1141) restated twice, independently: same chain from the same variates,
    for the fixed-jitter and the adaptive (RAM) proposal, normal and t errors."""
    import spec_numpy
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    ptr, idx = oracle.lattice_neighbors(w.array_row, w.array_col, w.platform)
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium")
    strings = synth.neighbor_strings(ptr, idx)
    nrep, thin, seed = 19, 3, 77
    ref = oracle.iterate_deconv(dc["positions2"], dc["dist"], strings, dc["Y2"], tdist, nrep, thin, dc["n"], dc["n0"],
                                w.d, 2.0, w.q, dc["init1"], dc["subspots"], False, jitter_scale, adapt_before, dc["c"],
                                w.mu0, w.lambda0, w.alpha, w.beta, 1, seed=seed)
    sp, si = oracle.find_subspot_neighbors((ptr, idx), dc["positions2"], dc["dist"])
    sub_nbrs = [si[sp[s]:sp[s + 1]] for s in range(dc["n"])]
    got = spec_numpy.run_deconv(oracle, dc["Y2"], sub_nbrs, tdist, nrep, thin, dc["n"], dc["n0"], w.d, 2.0, w.q,
                                dc["init1"], dc["subspots"], jitter_scale, adapt_before, dc["c"], np.asarray(w.mu0),
                                np.asarray(w.lambda0), w.alpha, w.beta, seed)
    assert np.array_equal(got["z"], ref["z"])
    assert np.array_equal(got["Ychange"][1:], ref["Ychange"][1:]) and np.isnan(ref["Ychange"][0])
    assert np.nanmax(ref["Ychange"]) > 0, "no proposal accepted: the Y path would be untested"
    assert np.allclose(got["mu"], ref["mu"], rtol=1e-9, atol=1e-10)
    assert np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-11, atol=0)
    assert np.allclose(got["weights"], ref["weights"], rtol=1e-10, atol=0)
    for r in range(nrep // thin + 1):
        assert np.allclose(np.asarray(got["lambda"][r]), np.asarray(ref["lambda"][r]), rtol=1e-8, atol=1e-12)
        assert np.allclose(got["Y"][r], np.asarray(ref["Y"][r]), rtol=1e-10, atol=1e-11)


def test_chromatic_visiting_order_in_both_restatements(oracle):
    """The GPU parity tests compare the CUDA chain with the oracle run in the CUDA path's visiting order (colour class
    by colour class).  That order is the oracle's only extension of the reference; the independent restatement given
    the same order must follow it step for step -- cluster sampler and deconvolution."""
    import spec_numpy
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    ptr, idx = oracle.lattice_neighbors(w.array_row, w.array_col, w.platform)
    col = (((w.array_col - 3 * w.array_row) // 2) % 3).astype(np.int32)      # closed-form 3-colouring of the hex lattice
    assert all(col[j] != col[u] for j in range(w.n) for u in idx[ptr[j]:ptr[j + 1]])
    order = np.lexsort((np.arange(w.n), col)).astype(np.int32)
    nbrs = [idx[ptr[j]:ptr[j + 1]] for j in range(w.n)]
    a = (7, 2, w.n, w.d, w.gamma, w.q, w.init)
    ref = oracle.iterate_t(w.Y, (ptr, idx), *a, w.mu0, w.lambda0, w.alpha, w.beta, seed=5, order=order)
    got = spec_numpy.run(oracle, "iterate_t", np.ascontiguousarray(w.Y), nbrs, *a, np.asarray(w.mu0),
                         np.asarray(w.lambda0), w.alpha, w.beta, 5, order=order)
    assert np.array_equal(got["z"], ref["z"]) and np.allclose(got["plogLik"][1:], ref["plogLik"][1:], rtol=1e-11)
    seq = oracle.iterate_t(w.Y, (ptr, idx), *a, w.mu0, w.lambda0, w.alpha, w.beta, seed=5)
    assert not np.array_equal(seq["z"], ref["z"])          # the order does matter: it is a different (valid) scan
    # deconvolution: parents in colour order
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium")
    strings = synth.neighbor_strings(ptr, idx)
    refd = oracle.iterate_deconv(dc["positions2"], dc["dist"], strings, dc["Y2"], True, 13, 3, dc["n"], dc["n0"], w.d,
                                 2.0, w.q, dc["init1"], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0,
                                 w.alpha, w.beta, 1, seed=9, order=order)
    sp, si = oracle.find_subspot_neighbors((ptr, idx), dc["positions2"], dc["dist"])
    gotd = spec_numpy.run_deconv(oracle, dc["Y2"], [si[sp[s]:sp[s + 1]] for s in range(dc["n"])], True, 13, 3, dc["n"],
                                 dc["n0"], w.d, 2.0, w.q, dc["init1"], dc["subspots"], 5.0, 0, dc["c"],
                                 np.asarray(w.mu0), np.asarray(w.lambda0), w.alpha, w.beta, 9, order=order)
    assert np.array_equal(gotd["z"], refd["z"]) and np.array_equal(gotd["Ychange"][1:], refd["Ychange"][1:])
    assert np.allclose(gotd["plogLik"][1:], refd["plogLik"][1:], rtol=1e-11)


# ---- the label kernel against an enumerated posterior ----------------------------------------------------------
def _cycle(n):
    return [np.array(sorted({(j - 1) % n, (j + 1) % n}), dtype=np.int32) for j in range(n)]


def _enumerate_posterior(logf, n, q, gamma, nbrs):
    """pi(z) ~ exp(sum_j log f_{z_j}(y_j) + gamma * #{edges with equal labels}) over all q^n configurations, coded
    sum_j (z_j - 1) q^j.  On a graph whose degrees are all 2 the conditional the sampler uses,
    (gamma / |N_j|) * 2 * #{u in N_j: z_u = k} (src/cluster.cpp:289), is the conditional of exactly this joint."""
    edges = {(min(j, int(u)), max(j, int(u))) for j in range(n) for u in nbrs[j]}
    logp = np.zeros(q ** n)
    for code in range(q ** n):
        z = [(code // q ** j) % q for j in range(n)]
        logp[code] = sum(logf[j, z[j]] for j in range(n)) + gamma * sum(z[a] == z[b] for a, b in edges)
    p = np.exp(logp - logp.max())
    return p / p.sum()


@pytest.mark.parametrize("n,q,chromatic", [(6, 2, False), (6, 2, True), (5, 3, False), (5, 3, True)])
def test_label_kernel_targets_the_enumerated_posterior_normal(oracle, n, q, chromatic):
    """Pins WHAT the restated z scan samples, independently of how src/cluster.cpp:272-306 was read: with mu and
    lambda held fixed, the long-run frequencies of all q^n label configurations must match the exact Potts x Gaussian
    posterior (density in the reference's form, quirk Q1 included), for the reference's sequential scan and for
    the colour-class order the CUDA path uses."""
    rng = np.random.default_rng(10 * n + q)
    d, gamma = 2, 1.1
    mu = rng.normal(size=(q, d)) * 0.9
    A = rng.normal(size=(d, d))
    Sig = A @ A.T / d + 0.6 * np.eye(d)
    Y = mu[rng.integers(0, q, size=n)] + rng.normal(size=(n, d))
    nbrs = _cycle(n)
    R = np.linalg.cholesky(Sig).T                   # Q1: quadratic form with (R R^T)^-1, R = chol_upper(Sigma)
    M = np.linalg.inv(R @ R.T)
    logf = np.array([[-np.log(np.diag(R)).sum() - d / 2 * np.log(2 * np.pi) - 0.5 * (y - m) @ M @ (y - m) for m in mu]
                     for y in Y])
    exact = _enumerate_posterior(logf, n, q, gamma, nbrs)
    order = None
    if chromatic:   # a proper colouring of the cycle, classes visited in sequence
        col = np.array([j % 2 for j in range(n)]) if n % 2 == 0 else np.array([j % 2 for j in range(n - 1)] + [2])
        order = np.lexsort((np.arange(n), col)).astype(np.int32)
    niter = 120000
    hist, _ = oracle.scan_fixed("iterate", Y, nbrs, gamma, q, rng.integers(1, q + 1, size=n), mu, np.linalg.inv(Sig),
                                niter, seed=5, order=order)
    freq = hist / niter
    assert 0.5 * np.abs(freq - exact).sum() < 0.012          # total variation
    top = np.argsort(exact)[-8:]
    assert np.all(np.abs(freq[top] - exact[top]) < 6 * np.sqrt(exact[top] * (1 - exact[top]) * 4 / niter) + 1e-3)


def test_label_and_weight_kernel_targets_the_student_t_posterior(oracle):
    """t model with the quirks switched off (compat = 0; even d, so Q2's integer division is exact): alternating
    w_j | z_j ~ Gamma((d+4)/2, scale 2 / (quad + 4)) (:513-518) with the label step on N(y; mu_k, Sigma / w_j) leaves the
    label marginal Potts x multivariate-t_4 invariant; compared with its enumeration."""
    rng = np.random.default_rng(77)
    n, q, d, gamma = 6, 2, 2, 0.8
    mu = rng.normal(size=(q, d))
    A = rng.normal(size=(d, d))
    Sig = A @ A.T / d + 0.5 * np.eye(d)
    Y = mu[rng.integers(0, q, size=n)] + rng.standard_t(4, size=(n, d))
    nbrs = _cycle(n)
    logf = np.array([[stats.multivariate_t.logpdf(y, loc=m, shape=Sig, df=4) for m in mu] for y in Y])
    exact = _enumerate_posterior(logf, n, q, gamma, nbrs)
    niter = 150000
    hist, wm = oracle.scan_fixed("iterate_t", Y, nbrs, gamma, q, np.ones(n, dtype=np.int32), mu, np.linalg.inv(Sig), niter,
                                 seed=9, compat=0)
    freq = hist / niter
    assert 0.5 * np.abs(freq - exact).sum() < 0.015
    assert np.all(wm > 0.2) and np.all(wm < 3.0)
