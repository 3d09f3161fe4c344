"""Host-side mirror of the reference's sampler interface, bound to libbsb200.so through ctypes.

The five functions below carry the names, argument order and argument meaning of the Rcpp exports
(/root/reference/R/RcppExports.R:4-22, src/cluster.cpp:217,323,446,570,742) and return dicts with
the reference's list names (z, mu, lambda, weights, Y, Ychange, plogLik, df_j).
No arithmetic happens in Python; everything runs in the CUDA library or raises.

Extra keyword arguments (no reference counterpart): seed (Philox key; the R shim draws it from R's
RNG), compat (quirk bitmask, default = reproduce the reference), device, colour (precomputed colour
classes for the chromatic scan).
"""
import ctypes as C

import numpy as np

from . import _lib, neighbors
from ._lib import (COMPAT_REFERENCE, ClusterArgs, ClusterOut, DeconvArgs, DeconvOut, check, colmajor, f64, i32,
                   lib, ptr)

MODEL = {"normal": 0, "t": 1}
PRECISION = {"equal": 0, "variable": 1}


class _Keep(list):
    """Keeps numpy buffers alive for the duration of an ABI call."""

    def add(self, a):
        self.append(a)
        return ptr(a)


def _cluster_args(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, model, precision,
                  seed, compat, device, colour, keep, groups=0, world=1, rank=0, comm_id=None):
    Y = np.asarray(Y, dtype=np.float64)
    if Y.shape != (n, d):
        raise _lib.BayesSpaceError(_lib.ERR_INVALID, "Dimensions of Y do not match n, d.")
    mu0 = f64(mu0)
    lambda0 = np.asarray(lambda0, dtype=np.float64)
    if mu0.shape[0] != d:
        raise _lib.BayesSpaceError(_lib.ERR_INVALID, "Dimensions of mu0 do not match input data Y.")
    if lambda0.shape != (d, d):
        raise _lib.BayesSpaceError(_lib.ERR_INVALID, "Dimensions of lambda0 do not match input data Y.")
    p, idx = neighbors.as_csr(df_j)
    if len(p) != n + 1 or len(np.asarray(init)) != n:
        raise _lib.BayesSpaceError(_lib.ERR_INVALID, "Invalid arguments!")
    a = ClusterArgs()
    a.model, a.precision = MODEL[model], PRECISION[precision]
    a.n, a.d, a.q, a.nrep, a.thin = n, d, q, nrep, thin
    a.gamma, a.alpha, a.beta = gamma, alpha, beta
    a.Y = keep.add(colmajor(Y))
    a.nbr_ptr, a.nbr_idx = keep.add(p), keep.add(idx if len(idx) else np.zeros(1, dtype=np.int32))
    a.init = keep.add(i32(init))
    a.mu0, a.lambda0 = keep.add(mu0), keep.add(colmajor(lambda0))
    a.seed, a.compat, a.device = seed, compat, device
    if colour is not None:
        col = i32(colour[0] if isinstance(colour, tuple) else colour)
        a.colour = keep.add(col)
        a.ncolours = int(colour[1]) if isinstance(colour, tuple) else int(col.max()) + 1
    a.groups, a.world, a.rank = int(groups), int(world), int(rank)
    if world > 1:
        if comm_id is None or len(comm_id) != 128:
            raise _lib.BayesSpaceError(_lib.ERR_INVALID, "world > 1 needs the 128-byte id of comm_unique_id()")
        a.comm_id = keep.add(np.frombuffer(bytes(comm_id), dtype=np.uint8).copy())
    return a


def _run_cluster(model, precision, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta,
                 seed=0, compat=COMPAT_REFERENCE, device=0, colour=None, groups=0, world=1, rank=0, comm_id=None,
                 mode_from=None, keep_z=True, keep_weights=True):
    """mode_from: also return 'z_mode', the per-spot mode of the label snapshots r >= mode_from with R's Mode() tie
    rule, computed on the device (what spatialCluster does with apply(zs, 2, Mode), R/spatialCluster.R:204-213);
    keep_z / keep_weights = False: the snapshot matrices are not materialised (they never leave the GPU)."""
    keep = _Keep()
    a = _cluster_args(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, model, precision,
                      seed, compat, device, colour, keep, groups, world, rank, comm_id)
    nsnap = nrep // thin + 1
    nl = q if precision == "variable" else 1
    z = np.zeros((nsnap, n), dtype=np.int32) if keep_z else None
    mu = np.zeros((nsnap, q * d))
    lam = np.zeros((nsnap, nl, d * d))
    w = np.zeros((nsnap, n)) if model == "t" and keep_weights else None
    pl = np.zeros(nrep)
    zm = np.zeros(n, dtype=np.int32) if mode_from is not None else None
    o = ClusterOut()
    o.z, o.mu, o.lambda_, o.weights, o.plogLik = ptr(z), ptr(mu), ptr(lam), ptr(w), ptr(pl)
    o.z_mode, o.mode_from = ptr(zm), 0 if mode_from is None else int(mode_from)
    check(lib().bsb200_cluster(C.byref(a), C.byref(o)))
    lam_m = lam.reshape(nsnap, nl, d, d).transpose(0, 1, 3, 2)   # col-major -> [i, j]
    out = {"mu": mu, "lambda": lam_m if nl > 1 else lam_m[:, 0], "plogLik": pl}
    if z is not None:
        out["z"] = z
    if w is not None:
        out["weights"] = w
    if zm is not None:
        out["z_mode"] = zm
    out["_info"] = {"ncolours": o.ncolours, "setup_ms": o.setup_ms, "device_ms": o.device_ms, "rows_done": o.rows_done,
                    "own": (int(o.own_lo), int(o.own_hi))}
    return out


def comm_unique_id():
    """128-byte communicator id (rank 0 calls this, every rank of a sharded chain receives the bytes)."""
    buf = np.zeros(128, dtype=np.uint8)
    check(lib().bsb200_comm_unique_id(ptr(buf)))
    return bytes(buf)


def debug_rendezvous(comm_id, world, rank, mine, rounds=1):
    """Host rendezvous of the peer transport alone (CPU): returns the (world, len(mine)) bytes of the last round."""
    mine = np.ascontiguousarray(mine, dtype=np.uint8)
    out = np.zeros((world, mine.size), dtype=np.uint8)
    idb = np.frombuffer(comm_id, dtype=np.uint8).copy()
    check(lib().bsb200_debug_rendezvous(ptr(idb), int(world), int(rank), ptr(mine), C.c_int64(mine.size), int(rounds),
                                        ptr(out)))
    return out


def shard_plan(df_j, colour, ncolours, groups, world, rank):
    """Host-only view of one rank's share of a sharded chain: own range, ghost ids, halo lists (original ids)."""
    p, idx = neighbors.as_csr(df_j)
    n = len(p) - 1
    col = i32(colour)
    lo, hi, ng = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    L = lib()
    check(L.bsb200_shard_plan(ptr(p), ptr(idx), C.c_int64(n), ptr(col), ncolours, groups, world, rank, C.byref(lo),
                              C.byref(hi), None, C.c_int64(0), C.byref(ng), None, None, None))
    ghosts = np.zeros(max(int(ng.value), 1), dtype=np.int32)
    counts = np.zeros(2 * ncolours * world, dtype=np.int32)
    check(L.bsb200_shard_plan(ptr(p), ptr(idx), C.c_int64(n), ptr(col), ncolours, groups, world, rank, C.byref(lo),
                              C.byref(hi), ptr(ghosts), C.c_int64(len(ghosts)), C.byref(ng), ptr(counts), None, None))
    K = ncolours * world
    send_ids = np.zeros(max(int(counts[:K].sum()), 1), dtype=np.int32)
    recv_ids = np.zeros(max(int(counts[K:].sum()), 1), dtype=np.int32)
    check(L.bsb200_shard_plan(ptr(p), ptr(idx), C.c_int64(n), ptr(col), ncolours, groups, world, rank, C.byref(lo),
                              C.byref(hi), ptr(ghosts), C.c_int64(len(ghosts)), C.byref(ng), ptr(counts), ptr(send_ids),
                              ptr(recv_ids)))
    return {"lo": int(lo.value), "hi": int(hi.value), "ghosts": ghosts[:int(ng.value)],
            "send_counts": counts[:K].reshape(ncolours, world), "recv_counts": counts[K:].reshape(ncolours, world),
            "send_ids": send_ids[:int(counts[:K].sum())], "recv_ids": recv_ids[:int(counts[K:].sum())]}


def iterate(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """Normal model, shared precision (src/cluster.cpp:217-321)."""
    return _run_cluster("normal", "equal", Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_vvv(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """Normal model, per-cluster precision (src/cluster.cpp:323-444)."""
    return _run_cluster("normal", "variable", Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_t(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """t model, shared precision (src/cluster.cpp:446-568)."""
    return _run_cluster("t", "equal", Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_t_vvv(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """t model, per-cluster precision (src/cluster.cpp:570-710)."""
    return _run_cluster("t", "variable", Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def init_cluster(Y, q, init=None, init_method="mclust", seed=0, max_iter=0, device=0, return_centers=False):
    """.init_cluster (R/spatialCluster.R:317-328): the given init, or labels 1..q from "kmeans" (k-means++ + Lloyd) /
    "mclust" (shared-covariance Gaussian mixture, EM) computed on the device."""
    if init is not None:
        return np.asarray(init, dtype=np.int32)
    if init_method not in ("mclust", "kmeans"):
        raise ValueError("'arg' should be one of 'mclust', 'kmeans'")
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    out = np.zeros(n, dtype=np.int32)
    cen = np.zeros((q, d))
    it = C.c_int32(0)
    yb = colmajor(Y)
    check(lib().bsb200_init_cluster(device, ptr(yb), C.c_int64(n), d, q, 1 if init_method == "mclust" else 0,
                                    C.c_uint64(seed), int(max_iter), ptr(out), ptr(cen), C.byref(it)))
    return (out, cen, int(it.value)) if return_centers else out


def cluster(Y, q, df_j, init=None, model="t", precision="equal", mu0=None, lambda0=None, gamma=3, alpha=1,
            beta=0.01, nrep=1000, thin=100, **kw):
    """cluster() (R/spatialCluster.R:70-112): picks the sampler by (model, precision)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    init = np.ones(n, dtype=np.int32) if init is None else init
    mu0 = Y.mean(axis=0) if mu0 is None else mu0
    lambda0 = 0.01 * np.eye(d) if lambda0 is None else lambda0
    if model not in MODEL or precision not in PRECISION:
        raise ValueError("'arg' should be one of the documented choices")
    if q == 1:
        return {"z": np.ones((1, n), dtype=np.int32)}
    fun = {("normal", "equal"): iterate, ("normal", "variable"): iterate_vvv, ("t", "equal"): iterate_t,
           ("t", "variable"): iterate_t_vvv}[(model, precision)]
    return fun(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_deconv(subspot_positions, dist, spot_neighbors, Y, tdist, nrep, thin, n, n0, d, gamma, q, init,
                   subspots, verbose, jitter_scale, adapt_before, c, mu0, lambda0, alpha, beta, thread_num=1,
                   seed=0, device=0, keep_Y=True, mean_from=None, mode_from=None, keep_z=True, keep_weights=True):
    """iterate_deconv (src/cluster.cpp:742-1141).  spot_neighbors: "a,b,c" strings / None (NA), or a
    CSR tuple.  Y (n x d) is not mutated (quirk Q10).  mean_from: also return the running mean of the
    Y snapshots r >= mean_from as 'Y_mean' (what spatialEnhance computes, R/spatialEnhance.R:441); mode_from: also
    return 'z_mode', the per-subspot mode of the label snapshots r >= mode_from (R/spatialEnhance.R:450-458)."""
    keep = _Keep()
    sp, si = spot_neighbors if isinstance(spot_neighbors, tuple) else neighbors.convert_neighbors(spot_neighbors)
    Y = np.asarray(Y, dtype=np.float64)
    pos = np.asarray(subspot_positions, dtype=np.float64)
    a = DeconvArgs()
    a.subspot_positions = keep.add(colmajor(pos))
    a.dist = dist
    a.spot_nbr_ptr, a.spot_nbr_idx = keep.add(np.ascontiguousarray(sp, dtype=np.int64)), keep.add(
        i32(si) if len(si) else np.zeros(1, dtype=np.int32))
    a.Y = keep.add(colmajor(Y))
    a.tdist, a.nrep, a.thin, a.n, a.n0, a.d = int(bool(tdist)), nrep, thin, n, n0, d
    a.gamma, a.q = gamma, q
    a.init = keep.add(i32(init))
    a.subspots, a.verbose = subspots, int(bool(verbose))
    a.jitter_scale, a.adapt_before, a.c = jitter_scale, adapt_before, c
    a.mu0, a.lambda0 = keep.add(f64(mu0)), keep.add(colmajor(lambda0))
    a.alpha, a.beta, a.thread_num, a.seed, a.device = alpha, beta, thread_num, seed, device
    if Y.shape != (n, d) or pos.shape != (n, 2) or len(np.asarray(init)) != n or len(sp) != n0 + 1:
        raise _lib.BayesSpaceError(_lib.ERR_INVALID, "Invalid arguments!")
    nsnap = nrep // thin + 1
    z = np.zeros((nsnap, n), dtype=np.int32) if keep_z else None
    mu = np.zeros((nsnap, q * d))
    lam = np.zeros((nsnap, d * d))
    w = np.zeros((nsnap, n)) if keep_weights else None
    Yo = np.zeros((nsnap, d, n)) if keep_Y else None
    Ym = np.zeros((d, n)) if mean_from is not None else None
    ych, pl = np.zeros(nrep), np.zeros(nrep)
    dfj = np.zeros(4 * n, dtype=np.int64)
    o = DeconvOut()
    o.z, o.mu, o.lambda_, o.weights, o.Y, o.Y_mean = ptr(z), ptr(mu), ptr(lam), ptr(w), ptr(Yo), ptr(Ym)
    o.mean_from = 0 if mean_from is None else int(mean_from)
    zm = np.zeros(n, dtype=np.int32) if mode_from is not None else None
    o.z_mode, o.mode_from = ptr(zm), 0 if mode_from is None else int(mode_from)
    o.Ychange, o.plogLik, o.df_j = ptr(ych), ptr(pl), ptr(dfj)
    check(lib().bsb200_deconv(C.byref(a), C.byref(o)))
    out = {"mu": mu, "lambda": lam.reshape(nsnap, d, d).transpose(0, 2, 1), "Ychange": ych,
           "plogLik": pl, "df_j": dfj.reshape(4, n).T.copy(),
           "_info": {"ncolours": o.ncolours, "setup_ms": o.setup_ms, "device_ms": o.device_ms, "rows_done": o.rows_done}}
    if keep_z:
        out["z"] = z
    if keep_weights:
        out["weights"] = w
    if keep_Y:
        out["Y"] = Yo.transpose(0, 2, 1)   # nsnap x n x d
    if mean_from is not None:
        out["Y_mean"] = Ym.T.copy()
    if zm is not None:
        out["z_mode"] = zm
    return out


class Chain:
    """A cluster sampler whose state stays resident in HBM between calls (bsb200_chain_*)."""

    def __init__(self, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, model="t",
                 precision="equal", seed=0, compat=COMPAT_REFERENCE, device=0, colour=None, groups=0, world=1, rank=0,
                 comm_id=None):
        keep = _Keep()
        a = _cluster_args(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, model, precision,
                          seed, compat, device, colour, keep, groups, world, rank, comm_id)
        self.n, self.d, self.q, self.nl, self.tdist = n, d, q, (q if precision == "variable" else 1), model == "t"
        self._h = C.c_void_p()
        check(lib().bsb200_chain_create(C.byref(a), C.byref(self._h)))

    def run(self, iters):
        ms = C.c_double(0)
        check(lib().bsb200_chain_run(self._h, int(iters), C.byref(ms)))
        return ms.value

    def profile(self, iters):
        sw, pm, hm, nl = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
        check(lib().bsb200_chain_profile(self._h, int(iters), C.byref(sw), C.byref(nl), C.byref(pm), C.byref(hm)))
        return {"sweep_ms": sw.value, "sweep_launches": int(nl.value), "param_ms": pm.value, "halo_ms": hm.value}

    def state(self):
        z = np.zeros(self.n, dtype=np.int32)
        w = np.zeros(self.n)
        mu = np.zeros((self.q, self.d))
        lam = np.zeros((self.nl, self.d * self.d))
        done = C.c_int64(0)
        check(lib().bsb200_chain_state(self._h, ptr(z), ptr(w), ptr(mu), ptr(lam), None, C.byref(done)))
        pl = np.zeros(int(done.value) + 1)
        check(lib().bsb200_chain_state(self._h, None, None, None, None, ptr(pl), C.byref(done)))
        return {"z": z, "weights": w, "mu": mu, "lambda": lam.reshape(self.nl, self.d, self.d).transpose(0, 2, 1),
                "plogLik": pl, "iters_done": int(done.value)}

    def dry_sweep(self):
        n = self.n
        zn, acc = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        w, hp, hn, pr = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        check(lib().bsb200_chain_dry_sweep(self._h, ptr(zn), ptr(w), ptr(hp), ptr(hn), ptr(pr), ptr(acc)))
        return {"z_new": zn, "w": w, "h_prev": hp, "h_new": hn, "prob": pr, "accepted": acc}

    def close(self):
        if self._h:
            lib().bsb200_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- state-level probes ---------------------------------------------------------------------------
def eval_loglik(Y, mu, mat, w=None, form=0, compat=COMPAT_REFERENCE, device=0):
    """n x q log-densities: form 0 dmvnrm_arma_fast(y, mu_k, mat_k / w), form 1 dmvnrm_prec_arma_fast(y, mu_k, mat_k * w)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    mu = f64(mu)
    q = mu.shape[0]
    mat = np.asarray(mat, dtype=np.float64)
    if mat.ndim == 2:
        mat = mat[None]
    nl = mat.shape[0]
    mb = np.concatenate([colmajor(m) for m in mat])
    yb = colmajor(Y)
    wb = None if w is None else f64(w)
    out = np.zeros(n * q)
    check(lib().bsb200_eval_loglik(device, form, C.c_uint32(compat), ptr(yb), C.c_int64(n), d, q, ptr(mu), ptr(mb), nl,
                                   ptr(wb), ptr(out)))
    return out.reshape(q, n).T.copy()


def eval_suffstats(Y, z, w, mu, nl=1, device=0):
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    mu = f64(mu)
    q = mu.shape[0]
    yb, zb = colmajor(Y), i32(z)
    wb = None if w is None else f64(w)
    nk, sk, S = np.zeros(q), np.zeros(q * d), np.zeros(nl * d * d)
    check(lib().bsb200_eval_suffstats(device, ptr(yb), C.c_int64(n), d, q, ptr(zb), ptr(wb), ptr(mu), nl, ptr(nk),
                                      ptr(sk), ptr(S)))
    Sm = S.reshape(nl, d, d).transpose(0, 2, 1)
    return nk, sk.reshape(q, d), (Sm if nl > 1 else Sm[0])


def debug_rng(kind, seed, purpose, index, it, slot=0, shape=1.0, scale=1.0, count=1, device=0):
    need = {0: 2, 3: 4}.get(kind, count)
    out = np.zeros(need)
    check(lib().bsb200_debug_rng(device, kind, C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(index),
                                 C.c_uint32(it), C.c_uint32(slot), C.c_double(shape), C.c_double(scale), count, ptr(out)))
    return out
