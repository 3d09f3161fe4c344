"""bsb200_init_cluster (.init_cluster, R/spatialCluster.R:317-328) against its numpy restatement (tests/spec_init.py):
same partition and centres, labels 1..q all present, deterministic under the seed, and good enough to start a chain."""
import numpy as np
import pytest

from spec_init import init_cluster as spec_init_cluster

pytestmark = pytest.mark.gpu


def _data(n_side=60, d=6, q=5, seed=8, tdist=False):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(seed)
    ar, ac = synth.square_lattice(n_side, n_side)
    truth = synth.voronoi_labels(ar, ac, q, rng)
    Y, _, _ = synth.make_pcs(truth, d, q, rng, tdist=tdist)
    return np.ascontiguousarray(Y), truth


@pytest.mark.parametrize("method", ["kmeans", "mclust"])
@pytest.mark.parametrize("q,d", [(5, 6), (7, 15), (15, 20)])
def test_init_cluster_matches_numpy_restatement(bsb, oracle, method, q, d):
    from sklearn.metrics import adjusted_rand_score as ari
    Y, truth = _data(d=d, q=q, seed=q * 10 + d)
    got, cen, iters = bsb.init_cluster(Y, q, init_method=method, seed=123, return_centers=True)
    want, wcen = spec_init_cluster(oracle, Y, q, method, 123)
    assert got.dtype == np.int32 and got.min() == 1 and got.max() == q and len(np.unique(got)) == q
    assert np.mean(got == want) > 0.999          # identical up to spots on a rounding knife edge
    assert np.allclose(cen, wcen, rtol=1e-6, atol=1e-8)
    assert ari(got, truth) > 0.6                 # a sensible partition of well separated synthetic domains
    again = bsb.init_cluster(Y, q, init_method=method, seed=123)
    assert np.array_equal(got, again)            # deterministic under the seed
    other = bsb.init_cluster(Y, q, init_method=method, seed=124)
    assert other.min() == 1 and other.max() == q


def test_init_cluster_feeds_the_sampler(bsb):
    """spatialCluster()'s sequence: .init_cluster, then cluster() (R/spatialCluster.R:158, :181) -- all on the device."""
    from bayesspace_b200 import synth
    from sklearn.metrics import adjusted_rand_score as ari
    w = synth.make_workload("small_hex")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    init = bsb.init_cluster(w.Y, w.q, init_method="mclust", seed=5)
    out = bsb.cluster(w.Y, w.q, csr, init=init, model="t", nrep=400, thin=10, seed=5, mode_from=11)
    assert ari(out["z_mode"], w.truth) > 0.8
    assert np.array_equal(bsb.init_cluster(w.Y, w.q, init=w.init), w.init)     # a given init is passed through
    with pytest.raises(ValueError):
        bsb.init_cluster(w.Y, w.q, init_method="hclust")


def test_init_cluster_large(bsb):
    """700k spots (C4 size): runs in well under a second per step on the device; every label present."""
    from bayesspace_b200 import synth
    w = synth.make_workload("c4")
    lab, cen, iters = bsb.init_cluster(w.Y, w.q, init_method="kmeans", seed=1, max_iter=30, return_centers=True)
    assert len(np.unique(lab)) == w.q and np.isfinite(cen).all() and iters >= 1
