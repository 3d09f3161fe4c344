"""Set-up of a chain on the device (csrc/setup_kernels.cu) against the host set-up (csrc/chain.cu).

The device path uploads the caller's CSR graph through the staged copies, checks the colouring, orders the spots by
(virtual shard, colour, original id) with a stable radix sort and writes the neighbour lists in internal order; the host
path (BSB200_HOST_SETUP=1) does the same with host threads.  Sharded chains take the device path too (own spots and
ghosts of every rank): tests/mgpu_check.py compares them bit for bit with the unsharded chain.  Both must give the same
internal order, hence bit-identical chains: labels, parameters, weights and plogLik are compared with array_equal.
The chains themselves are checked against the CPU oracle elsewhere (test_gpu_parity.py, test_gpu_zy_clusters.py).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _workload(platform, rows, cols, d, q, seed):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(seed)
    ar, ac = (synth.hex_lattice if platform == "Visium" else synth.square_lattice)(rows, cols)
    truth = synth.voronoi_labels(ar, ac, q, rng)
    Y, _, _ = synth.make_pcs(truth, d, q, rng, tdist=True)
    w = synth.Workload(array_row=ar, array_col=ac, n=int(ar.shape[0]), d=d, q=q, gamma=2.0, Y=Y,
                       init=synth.noisy_init(truth, q, rng), truth=truth, platform=platform)
    w.update(synth.priors(Y))
    return w


def _both(bsb, fn, **kw):
    out = []
    for host in (False, True):
        if host:
            os.environ["BSB200_HOST_SETUP"] = "1"
        try:
            out.append(fn(**kw))
        finally:
            os.environ.pop("BSB200_HOST_SETUP", None)
    return out


def _same(a, b, keys=("z", "mu", "lambda", "weights", "plogLik")):
    for k in keys:
        if k in a or k in b:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert a["_info"]["ncolours"] == b["_info"]["ncolours"]


@pytest.mark.parametrize("platform,rows,cols,model", [("Visium", 40, 50, "t"), ("ST", 33, 47, "normal"),
                                                      ("ST", 600, 520, "t")])   # the last one has 64 virtual shards
def test_device_setup_matches_host_setup(bsb, platform, rows, cols, model):
    w = _workload(platform, rows, cols, 7, 5, seed=rows)
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    colour = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    run = lambda: bsb.cluster(w.Y, w.q, csr, init=w.init, model=model, precision="equal", mu0=w.mu0, lambda0=w.lambda0,
                              gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=7, thin=2, seed=3, colour=colour)
    dev, host = _both(bsb, run)
    _same(dev, host)


def test_device_setup_without_and_with_improper_colouring(bsb):
    """No colouring given, or one that is not proper (two neighbours share a class): both paths colour the graph
    themselves (greedy, host) and agree; a colour id out of range is refused by both."""
    w = _workload("ST", 30, 31, 5, 4, seed=9)
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    kw = dict(init=w.init, model="t", precision="variable", mu0=w.mu0, lambda0=w.lambda0, gamma=w.gamma, alpha=w.alpha,
              beta=w.beta, nrep=5, thin=1, seed=4)
    none = _both(bsb, lambda: bsb.cluster(w.Y, w.q, csr, **kw))
    _same(*none)
    bad = (np.arange(w.n) // 2 % 2).astype(np.int32)   # consecutive spots are neighbours and share a class
    improper = _both(bsb, lambda: bsb.cluster(w.Y, w.q, csr, colour=(bad, 2), **kw))
    _same(*improper)
    _same(none[0], improper[0])
    worse = bad.copy()
    worse[5] = 7
    for host in (False, True):
        if host:
            os.environ["BSB200_HOST_SETUP"] = "1"
        try:
            with pytest.raises(bsb.BayesSpaceError):
                bsb.cluster(w.Y, w.q, csr, colour=(worse, 2), **kw)
        finally:
            os.environ.pop("BSB200_HOST_SETUP", None)


def test_device_setup_general_graph(bsb):
    """Radius graph: more than 6 neighbours for some spots, none for others; no colouring given."""
    rng = np.random.default_rng(5)
    n, d, q = 3000, 4, 3
    pos = rng.uniform(0, 60, size=(n, 2))
    pos[:5] += 1000.0   # isolated
    ptr, idx = bsb.neighbors.as_csr(bsb.find_neighbors(pos, radius=1.2, method="euclidean"))
    deg = np.diff(ptr)
    assert deg.min() == 0 and deg.max() > 6
    truth = rng.integers(0, q, n)
    Y = rng.normal(size=(n, d)) + truth[:, None]
    init = (truth + 1).astype(np.int32)
    run = lambda: bsb.cluster(Y, q, (ptr, idx), init=init, model="normal", precision="equal", mu0=Y.mean(axis=0),
                              lambda0=0.01 * np.eye(d), gamma=2.0, alpha=1.0, beta=0.01, nrep=5, thin=1, seed=2)
    _same(*_both(bsb, run))


def test_release_memory_and_reuse(bsb):
    """Device memory of a destroyed chain stays in the pool for the next one; bsb200_release_memory() hands it back and
    the library keeps working (the next chain allocates afresh)."""
    import ctypes as C
    w = _workload("ST", 64, 64, 6, 4, seed=2)
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    colour = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    run = lambda: bsb.cluster(w.Y, w.q, csr, init=w.init, model="t", precision="equal", mu0=w.mu0, lambda0=w.lambda0,
                              gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=5, thin=1, seed=3, colour=colour)
    first = run()
    second = run()          # served from the pool
    bsb.release_memory(0)
    third = run()
    _same(first, second)
    _same(first, third)
    with pytest.raises(bsb.BayesSpaceError):
        bsb.release_memory(99)
