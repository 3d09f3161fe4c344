"""src/compare_enhanced.cpp (map_subspot2ref, compute_corr): oracle restatement against numpy (CPU), CUDA path against
the oracle (GPU)."""
import numpy as np
import pytest


def _coords(seed=3, ns=700, nr=90):
    rng = np.random.default_rng(seed)
    ref = np.c_[np.repeat(np.arange(9), 10), np.tile(np.arange(10), 9)].astype(float) * 2.0
    sub = rng.uniform(-1, 19, size=(ns, 2))
    sub[:40] = ref[rng.integers(0, nr, 40)] + np.array([1.0, 0.0])      # exactly between two reference spots: ties
    sub[40:60] = ref[rng.integers(0, nr, 20)] + np.array([1.0, 1.0])    # centre of a cell: four-way ties
    return sub, ref


def test_oracle_map_subspot2ref_matches_numpy(oracle):
    sub, ref = _coords()
    got = oracle.map_subspot2ref(sub, ref)
    d2 = ((sub[:, None, :] - ref[None, :, :]) ** 2).sum(axis=2)
    rows = [(i, j, d2[i].min()) for i in range(len(sub)) for j in np.flatnonzero(d2[i] == d2[i].min())]
    assert got.shape == (len(rows), 3) and len(rows) > len(sub)         # there are ties
    assert np.array_equal(got[:, :2], np.array(rows)[:, :2]) and np.allclose(got[:, 2], np.array(rows)[:, 2], rtol=1e-15)


def test_oracle_compute_corr_matches_numpy(oracle):
    rng = np.random.default_rng(4)
    a, b = rng.normal(size=(50, 37)), rng.normal(size=(50, 37))
    b[:10] = 0.7 * a[:10] + 0.1 * b[:10]
    got = oracle.compute_corr(a, b)
    ref = np.array([np.corrcoef(a[i], b[i])[0, 1] for i in range(50)])
    assert np.array_equal(got[:, 0], np.arange(50)) and np.allclose(got[:, 1], ref, rtol=1e-12)
    with pytest.raises(RuntimeError, match="same shape"):
        oracle.compute_corr(a, b[:, :5])


@pytest.mark.gpu
def test_map_subspot2ref_matches_oracle(bsb, oracle):
    sub, ref = _coords(ns=5000)
    got = bsb.map_subspot2ref(sub, ref)
    want = oracle.map_subspot2ref(sub, ref)
    assert got.shape == want.shape and np.array_equal(got, want)         # ids, ties and distances bit for bit
    # the subspot geometry spatialEnhance builds: every subspot maps back to its own parent spot
    from bayesspace_b200 import synth
    ar, ac = synth.hex_lattice(20, 16)
    pos = np.c_[ac.astype(float), ar * np.sqrt(3.0)]
    shift, _ = synth.make_subspots("Visium", 1.0, np.sqrt(3.0))
    sub2 = np.tile(pos, (6, 1)) + np.repeat(shift, len(pos), axis=0)
    m = bsb.map_subspot2ref(sub2, pos)
    assert len(m) == len(sub2) and np.array_equal(m[:, 1], np.tile(np.arange(len(pos)), 6))


@pytest.mark.gpu
def test_compute_corr_matches_oracle(bsb, oracle):
    rng = np.random.default_rng(5)
    a, b = rng.normal(size=(3000, 61)), rng.normal(size=(3000, 61))
    b[::3] += a[::3]
    got, want = bsb.compute_corr(a, b), oracle.compute_corr(a, b)
    assert np.array_equal(got[:, 0], want[:, 0]) and np.allclose(got[:, 1], want[:, 1], rtol=1e-12, atol=1e-14)
    with pytest.raises(bsb.BayesSpaceError, match="same shape"):
        bsb.compute_corr(a, b[:, :7])
