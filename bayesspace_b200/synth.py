"""Synthetic inputs for the sampler hot path: lattices, ground-truth domains, PCs.

Pure numpy, no compute of the path itself.  Used by tests/ and bench.py so that both see the same
seeded inputs.  Geometry follows the reference's conventions:
  * Visium hex: array_row in 0..R-1, array_col = 2*c + (row mod 2)   (R/spatialCluster.R:228-233)
  * ST / VisiumHD square: array_row in 1..R, array_col in 1..C       (R/utils.R:149-152 exampleSCE)
"""
import numpy as np


def hex_lattice(nrows=78, ncols=64):
    r = np.repeat(np.arange(nrows, dtype=np.int32), ncols)
    c = np.tile(np.arange(ncols, dtype=np.int32), nrows)
    return r, (2 * c + (r % 2)).astype(np.int32)


def square_lattice(nrows, ncols, one_based=True):
    r = np.repeat(np.arange(nrows, dtype=np.int32), ncols) + (1 if one_based else 0)
    c = np.tile(np.arange(ncols, dtype=np.int32), nrows) + (1 if one_based else 0)
    return r.astype(np.int32), c.astype(np.int32)


def voronoi_labels(array_row, array_col, q, rng, hex_geometry=False):
    """Ground-truth domains: nearest of q random seeds on lattice coordinates (1-based labels)."""
    x = array_col.astype(np.float64)
    y = array_row.astype(np.float64) * (np.sqrt(3.0) if hex_geometry else 1.0)
    sx = rng.uniform(x.min(), x.max(), q)
    sy = rng.uniform(y.min(), y.max(), q)
    best = np.full(x.shape, np.inf)
    lab = np.zeros(x.shape, dtype=np.int32)
    for k in range(q):   # loop over seeds keeps memory O(n) at 10M spots
        dist = (x - sx[k]) ** 2 + (y - sy[k]) ** 2
        m = dist < best
        best[m] = dist[m]
        lab[m] = k + 1
    return lab


def make_pcs(truth, d, q, rng, tdist=False):
    """Y = mu_z + Sigma^{1/2} eps / sqrt(w), column-centred.  Returns (Y, mu_true, Sigma_true)."""
    n = truth.shape[0]
    scales = 6.0 * 0.85 ** np.arange(d)
    mu = rng.normal(size=(q, d)) * scales
    A = rng.normal(size=(d, d))
    Sigma = A @ A.T / d + 0.5 * np.eye(d)
    L = np.linalg.cholesky(Sigma)
    Y = np.empty((n, d))
    chunk = 1 << 20
    for s in range(0, n, chunk):   # chunked to bound temporaries at 10M spots
        e = min(n, s + chunk)
        eps = rng.normal(size=(e - s, d)) @ L.T
        if tdist:
            w = rng.gamma(2.0, 1.0 / 2.0, size=e - s)
            eps /= np.sqrt(w)[:, None]
        Y[s:e] = mu[truth[s:e] - 1] + eps
    Y -= Y.mean(axis=0)
    return np.asfortranarray(Y), mu, Sigma   # column-major like an R matrix: the C ABI takes it without a copy


def noisy_init(truth, q, rng, frac=0.2):
    init = truth.copy()
    m = rng.random(truth.shape[0]) < frac
    init[m] = rng.integers(1, q + 1, size=int(m.sum()))
    # every cluster must be present (the VVV samplers cannot survive an empty cluster, quirk Q9)
    for k in range(1, q + 1):
        if not np.any(init == k):
            init[rng.integers(0, truth.shape[0])] = k
    return init.astype(np.int32)


def priors(Y):
    """R/spatialCluster.R:166-178 defaults."""
    d = Y.shape[1]
    return dict(mu0=Y.mean(axis=0), lambda0=0.01 * np.eye(d), alpha=1.0, beta=0.01)


class Workload(dict):
    __getattr__ = dict.__getitem__


def make_workload(name, seed=149, n_override=None):
    """The BASELINE.json configs (C1..C5) and small test cases, as dicts of arrays + scalars."""
    rng = np.random.default_rng(seed)
    cfg = {
        # name: (platform, rows, cols, q, d, model, precision, gamma)
        "c1": ("ST", 8, 12, 4, 10, "normal", "equal", 2.0),
        "c2": ("Visium", 78, 64, 7, 15, "t", "equal", 3.0),
        "c4": ("VisiumHD", 837, 837, 15, 20, "t", "equal", 2.0),
        "c5": ("VisiumHD", 3200, 3125, 7, 15, "t", "equal", 2.0),
        "tiny_hex": ("Visium", 12, 10, 3, 4, "t", "equal", 3.0),
        "small_hex": ("Visium", 30, 24, 4, 6, "t", "equal", 3.0),
        "small_sq": ("ST", 20, 25, 3, 5, "normal", "equal", 2.0),
    }[name]
    platform, rows, cols, q, d, model, precision, gamma = cfg
    if n_override is not None:
        rows, cols = n_override
    if platform == "Visium":
        ar, ac = hex_lattice(rows, cols)
    else:
        ar, ac = square_lattice(rows, cols)
    truth = voronoi_labels(ar, ac, q, rng, hex_geometry=(platform == "Visium"))
    Y, mu_true, Sigma_true = make_pcs(truth, d, q, rng, tdist=(model == "t"))
    init = noisy_init(truth, q, rng)
    w = Workload(name=name, platform=platform, array_row=ar, array_col=ac, n=int(ar.shape[0]), d=d, q=q,
                 model=model, precision=precision, gamma=gamma, truth=truth, Y=Y, init=init,
                 mu_true=mu_true, Sigma_true=Sigma_true, rows=rows, cols=cols)
    w.update(priors(Y))
    return w


# ---- subspot geometry (R/spatialEnhance.R:119-167, 181-244) -------------------------------------
def make_subspots(platform, xdist, ydist, nsubspots_per_edge=3, tolerance=1.05):
    """.make_subspots: shift table (S x 2) and neighbour distance threshold."""
    if platform == "Visium":
        if abs(xdist) >= abs(ydist):
            raise ValueError("Unable to find neighbors of subspots. Please raise an issue to maintainers.")
        # expand.grid varies x fastest
        shift = np.array([(1 / 3, 1 / 3), (-1 / 3, 1 / 3), (1 / 3, -1 / 3), (-1 / 3, -1 / 3),
                          (2 / 3, 0.0), (-2 / 3, 0.0)])
        shift = shift * np.array([xdist, ydist])
        dist = np.abs(shift).sum(axis=1).max() * tolerance
    elif platform in ("VisiumHD", "ST"):
        k = nsubspots_per_edge
        v = np.arange(1 / (2 * k), 1, 1 / k)[:k] - 0.5
        shift = np.array([(x, y) for y in v for x in v]) * np.array([xdist, ydist])
        dist = tolerance / k
    else:
        raise ValueError("Only data from Visium, VisiumHD and ST currently supported.")
    return shift, float(dist)


def deconv_inputs(Y, positions, xdist, ydist, init, platform="Visium", nsubspots_per_edge=3,
                  jitter_prior=0.3):
    """What deconvolve() hands to iterate_deconv (R/spatialEnhance.R:119-156)."""
    n0, d = Y.shape
    c = jitter_prior * 1.0 / (2.0 * np.mean(np.diag(np.atleast_2d(np.cov(Y, rowvar=False)))))
    S = 6 if platform == "Visium" else nsubspots_per_edge ** 2
    shift, dist = make_subspots(platform, xdist, ydist, nsubspots_per_edge)
    init1 = np.tile(init, S).astype(np.int32)
    Y2 = np.tile(Y, (S, 1))
    pos2 = np.tile(np.asarray(positions, dtype=np.float64), (S, 1)) + np.repeat(shift, n0, axis=0)
    return dict(positions2=pos2, dist=dist, Y2=Y2, init1=init1, subspots=S, c=float(c), n=n0 * S, n0=n0)


def neighbor_strings(ptr, idx):
    """sce$spot.neighbors: "1-based,comma" strings, None for isolated spots."""
    out = []
    for j in range(len(ptr) - 1):
        nb = idx[ptr[j]:ptr[j + 1]]
        out.append(None if len(nb) == 0 else ",".join(str(int(v) + 1) for v in nb))
    return out
