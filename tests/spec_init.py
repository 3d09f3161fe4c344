"""numpy restatement of the initialiser behind bsb200_init_cluster (bayesspace_b200/csrc/init_cluster.cu): k-means++
seeding on Philox-keyed uniforms + Lloyd iterations, optionally followed by EM of the shared-covariance Gaussian
mixture (mclust's "EEE" model).  It replaces .init_cluster (/root/reference/R/spatialCluster.R:317-328) semantically
(labels 1..q, deterministic under the seed); stats::kmeans / mclust::Mclust themselves are not restated (different
algorithms on R's RNG).  Test infrastructure."""
import numpy as np

P_INIT = 8
CHUNK = 1024


def seed_uniform(O, seed, step):
    o = O.philox([step, 0, 0, P_INIT], [seed & 0xffffffff, seed >> 32])
    x = (int(o[1]) << 32) | int(o[0])
    return ((x >> 11) + 0.5) / 9007199254740992.0


def kmeans_pp(O, Y, q, seed):
    n = Y.shape[0]
    picks = [min(int(np.floor(seed_uniform(O, seed, 0) * n)), n - 1)]
    d2 = None
    for k in range(1, q):
        v = ((Y - Y[picks[-1]]) ** 2).sum(axis=1)
        d2 = v if d2 is None else np.minimum(d2, v)
        bsum = np.array([d2[b:b + CHUNK].sum() for b in range(0, n, CHUNK)])
        total = bsum.sum()
        target = seed_uniform(O, seed, k) * total
        run, b = 0.0, 0
        while b < len(bsum) - 1 and run + bsum[b] < target:
            run += bsum[b]
            b += 1
        cs = np.cumsum(d2[b * CHUNK:(b + 1) * CHUNK])
        hit = np.flatnonzero(cs >= target - run)
        picks.append(b * CHUNK + (int(hit[0]) if len(hit) else len(cs) - 1))
    return Y[picks].copy()


def lloyd(Y, cen, max_iter=100):
    z = np.full(Y.shape[0], -1)
    it = 0
    for it in range(max_iter):
        d2 = ((Y[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2)
        znew = d2.argmin(axis=1)
        if np.array_equal(znew, z):
            break
        z = znew
        for k in range(cen.shape[0]):
            if np.any(z == k):
                cen[k] = Y[z == k].mean(axis=0)
    return z, cen, it


def em_eee(Y, z, q, max_iter=100, tol=1e-5):
    n, d = Y.shape
    R = np.zeros((n, q))
    R[np.arange(n), z] = 1.0
    T = Y.T @ Y
    prev = -np.inf
    for _ in range(max_iter):
        nk = np.maximum(R.sum(axis=0), 1e-300)
        mu = (R.T @ Y) / nk[:, None]
        Sig = (T - (mu * nk[:, None]).T @ mu) / n
        M = np.linalg.inv(Sig)
        _, logdet = np.linalg.slogdet(Sig)
        g = mu @ M
        lc = np.log(nk / n) - 0.5 * np.einsum("kd,kd->k", g, mu)
        a = Y @ g.T + lc
        mx = a.max(axis=1)
        s = np.exp(a - mx[:, None]).sum(axis=1)
        R = np.exp(a - mx[:, None]) / s[:, None]
        z = a.argmax(axis=1)
        loglik = (mx + np.log(s)).sum() - 0.5 * np.trace(M @ T) - 0.5 * n * (d * np.log(2 * np.pi) + logdet)
        if abs(loglik - prev) <= tol * abs(loglik):
            break
        prev = loglik
    return z, mu


def init_cluster(O, Y, q, method, seed):
    cen = kmeans_pp(O, Y, q, seed)
    z, cen, _ = lloyd(Y, cen)
    if method == "mclust" and q > 1:
        z, cen = em_eee(Y, z, q)
    return z + 1, cen
