"""ctypes front-end of the CPU oracle (oracle/ref_cpu.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under bayesspace_b200/ does.  See the header of ref_cpu.cpp for what is
restated and the parity status ("parity unpinned": the reference cannot run in this image).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

COMPAT_Q1 = 1
COMPAT_Q3 = 4
COMPAT_REFERENCE = COMPAT_Q1 | COMPAT_Q3

VARIANTS = {"iterate": 0, "iterate_vvv": 1, "iterate_t": 2, "iterate_t_vvv": 3}

P_MU, P_WISH_NORM, P_WISH_CHI, P_SPOT_U, P_SPOT_GAMMA, P_ERR, P_SPOT_YACC = 1, 2, 3, 4, 5, 6, 7


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "ref_cpu.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_radius_neighbors.restype = C.c_int64
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(a):
    """n x d array -> flat column-major float64 buffer (R's layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64)).ravel(order="F")


def csr_from_lists(df_j):
    ptr = np.zeros(len(df_j) + 1, dtype=np.int64)
    for i, v in enumerate(df_j):
        ptr[i + 1] = ptr[i] + len(v)
    idx = np.zeros(max(int(ptr[-1]), 1), dtype=np.int32)
    for i, v in enumerate(df_j):
        idx[ptr[i]:ptr[i + 1]] = np.asarray(v, dtype=np.int32)
    return ptr, idx


def lists_from_csr(ptr, idx):
    return [idx[ptr[i]:ptr[i + 1]].copy() for i in range(len(ptr) - 1)]


# ---------------------------------------------------------------------------------------------
def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def uniform2(seed, purpose, index, it, slot):
    o = np.zeros(2)
    lib().orc_uniform2(C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(index), C.c_uint32(it),
                       C.c_uint32(slot), _p(o, C.c_double))
    return o


def normal(seed, purpose, index, it, m0, count):
    o = np.zeros(count)
    lib().orc_normal(C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(index), C.c_uint32(it),
                     C.c_uint32(m0), C.c_int(count), _p(o, C.c_double))
    return o


def gamma(seed, purpose, index0, count, it, shape, scale):
    o = np.zeros(count)
    lib().orc_gamma(C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(index0), C.c_int(count),
                    C.c_uint32(it), C.c_double(shape), C.c_double(scale), _p(o, C.c_double))
    return o


def spot_draw(seed, index, it):
    """(proposal bits, u_accept, u of the first gamma attempt) of one spot and iteration: one Philox block."""
    o = np.zeros(3)
    lib().orc_spot_draw(C.c_uint64(seed), C.c_uint32(index), C.c_uint32(it), _p(o, C.c_double))
    return int(o[0]), float(o[1]), float(o[2])


def gamma_spot(seed, index, it, shape, scale):
    """t-model weight draw of one spot (Marsaglia-Tsang on the per-spot draw layout)."""
    f = lib().orc_gamma_spot
    f.restype = C.c_double
    return float(f(C.c_uint64(seed), C.c_uint32(index), C.c_uint32(it), C.c_double(shape), C.c_double(scale)))


def propose_label_bits(z_prev, q, bits):
    return int(lib().orc_propose_label_bits(int(z_prev), int(q), C.c_uint32(bits)))


def dmvnrm_arma_fast(x, mean, sigma, w=None, q1_compat=True):
    """log N(x; mean, sigma[/w]) exactly as src/cluster.cpp:83-106 evaluates it."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, d = x.shape
    out = np.zeros(n)
    xb, mb, sb = _colmajor(x), _f64(mean), _colmajor(sigma)
    wb = None if w is None else _f64(w)
    lib().orc_dmvnrm(0, _p(xb, C.c_double), C.c_int64(n), d, _p(mb, C.c_double), _p(sb, C.c_double),
                     _p(wb, C.c_double), int(q1_compat), _p(out, C.c_double))
    return out


def dmvnrm_prec_arma_fast(x, mean, lam, w=None):
    """log N(x; mean, (lam*w)^-1) as src/cluster.cpp:109-132."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n, d = x.shape
    out = np.zeros(n)
    xb, mb, sb = _colmajor(x), _f64(mean), _colmajor(lam)
    wb = None if w is None else _f64(w)
    lib().orc_dmvnrm(1, _p(xb, C.c_double), C.c_int64(n), d, _p(mb, C.c_double), _p(sb, C.c_double),
                     _p(wb, C.c_double), 1, _p(out, C.c_double))
    return out


def chol_upper(X):
    d = X.shape[0]
    xb = _colmajor(X)
    out = np.zeros(d * d)
    rc = lib().orc_chol_upper(_p(xb, C.c_double), d, _p(out, C.c_double))
    if rc:
        raise np.linalg.LinAlgError("not positive definite")
    return out.reshape(d, d, order="F")


def inv(X):
    d = X.shape[0]
    xb = _colmajor(X)
    out = np.zeros(d * d)
    rc = lib().orc_inv(_p(xb, C.c_double), d, _p(out, C.c_double))
    if rc:
        raise np.linalg.LinAlgError("singular")
    return out.reshape(d, d, order="F")


def suffstats(Y, z, w, mu):
    """n_k, s_k (q x d), pooled scatter S (d x d), per-cluster scatter (q x d x d) about mu_{z_j}."""
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    mu = _f64(mu)
    q = mu.shape[0]
    yb = _colmajor(Y)
    zb = np.ascontiguousarray(z, dtype=np.int32)
    wb = None if w is None else _f64(w)
    nk, sk, S, Sk = np.zeros(q), np.zeros(q * d), np.zeros(d * d), np.zeros(q * d * d)
    lib().orc_suffstats(_p(yb, C.c_double), C.c_int64(n), d, q, _p(zb, C.c_int32), _p(wb, C.c_double),
                        _p(mu, C.c_double), _p(nk, C.c_double), _p(sk, C.c_double), _p(S, C.c_double),
                        _p(Sk, C.c_double))
    return nk, sk.reshape(q, d), S.reshape(d, d, order="F"), np.stack(
        [Sk[k * d * d:(k + 1) * d * d].reshape(d, d, order="F") for k in range(q)])


def h_values(variant, Y, df_j, z, w, mu, sigma, gamma, z_new, compat=COMPAT_REFERENCE):
    """h(z_prev), h(z_new) per spot on a fixed state (the quantities the MH step compares)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    mu = _f64(mu)
    q = mu.shape[0]
    ptr, idx = csr_from_lists(df_j) if isinstance(df_j, list) else df_j
    sig = np.asarray(sigma, dtype=np.float64)
    if sig.ndim == 2:
        sig = sig[None]
    sb = np.concatenate([_colmajor(s) for s in sig])
    yb = _colmajor(Y)
    zb = np.ascontiguousarray(z, dtype=np.int32)
    znb = np.ascontiguousarray(z_new, dtype=np.int32)
    wb = _f64(np.ones(n) if w is None else w)
    hp, hn = np.zeros(n), np.zeros(n)
    v = VARIANTS[variant] if isinstance(variant, str) else variant
    lib().orc_h_values(v, C.c_uint32(compat), _p(yb, C.c_double), C.c_int64(n), d, q,
                       _p(ptr, C.c_int64), _p(idx, C.c_int32), _p(zb, C.c_int32), _p(wb, C.c_double),
                       _p(mu, C.c_double), _p(sb, C.c_double), C.c_double(gamma), _p(znb, C.c_int32),
                       _p(hp, C.c_double), _p(hn, C.c_double))
    return hp, hn


_ERR = {1: "Invalid arguments!", 2: "Error in finding neighbors of subspots!", 3: "mu update failed",
        4: "scatter inverse failed", 5: "Wishart draw failed (empty cluster / df <= 0)",
        6: "NA in probability vector", 7: "adaptive Cholesky failed", 8: "degree > 4"}


def scan_fixed(variant, Y, df_j, gamma, q, init, mu, lam, niter, seed=0, order=None, compat=COMPAT_REFERENCE):
    """niter z (and w) scans of a cluster sampler with FIXED mu (q x d) and lambda (d x d, or q x d x d for the VVV
    samplers): the label-update Markov kernel alone.  Returns (histogram over the q^n label configurations, coded
    sum_j (z_j - 1) q^j; mean of the weights)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, d = Y.shape
    v = VARIANTS[variant] if isinstance(variant, str) else variant
    ptr, idx = df_j if isinstance(df_j, tuple) else csr_from_lists(df_j)
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    lam = np.asarray(lam, dtype=np.float64)
    lam3 = lam if lam.ndim == 3 else lam[None]
    lb = np.concatenate([_colmajor(m) for m in lam3])
    sb = np.concatenate([_colmajor(np.linalg.inv(m)) for m in lam3])
    yb, mub, zb = _colmajor(Y), _f64(mu).ravel(), np.ascontiguousarray(init, dtype=np.int32)
    ob = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    hist = np.zeros(q ** n, dtype=np.int64)
    wm = np.zeros(n)
    rc = lib().orc_scan_fixed(v, C.c_uint32(compat), _p(yb, C.c_double), C.c_int64(n), d, q, _p(ptr, C.c_int64),
                              _p(idx, C.c_int32), C.c_double(gamma), _p(zb, C.c_int32), _p(mub, C.c_double),
                              _p(lb, C.c_double), _p(sb, C.c_double), C.c_uint64(seed), _p(ob, C.c_int32), int(niter),
                              _p(hist, C.c_int64), _p(wm, C.c_double))
    if rc:
        raise RuntimeError("oracle error %d" % rc)
    return hist, wm


def map_subspot2ref(subspot_coords, ref_coords):
    """src/compare_enhanced.cpp:47-92 (single thread): rows (i, j, squared distance), all exact ties."""
    s = np.asarray(subspot_coords, dtype=np.float64)
    r = np.asarray(ref_coords, dtype=np.float64)
    ns, k = s.shape
    nr = r.shape[0]
    cap = 4 * ns + 64
    out = np.zeros((cap, 3))
    f = lib().orc_map_subspot2ref
    f.restype = C.c_int64
    rows = f(_p(_colmajor(s), C.c_double), C.c_int64(ns), _p(_colmajor(r), C.c_double), C.c_int64(nr), k,
             _p(out, C.c_double), C.c_int64(cap))
    if rows < 0:
        raise RuntimeError("too many ties")
    return out[:rows].copy()


def compute_corr(m1, m2):
    """src/compare_enhanced.cpp:94-132: matrix (row index, Pearson correlation of the two rows)."""
    a = np.asarray(m1, dtype=np.float64)
    b = np.asarray(m2, dtype=np.float64)
    if a.shape != b.shape:
        raise RuntimeError("Input matrices should be of the same shape.")
    n, p = a.shape
    cor = np.zeros(n)
    lib().orc_compute_corr(_p(_colmajor(a), C.c_double), _p(_colmajor(b), C.c_double), C.c_int64(n), C.c_int64(p),
                           _p(cor, C.c_double))
    return np.c_[np.arange(n, dtype=np.float64), cor]


def _run_cluster(variant, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta,
                 seed=0, order=None, compat=COMPAT_REFERENCE):
    Y = np.asarray(Y, dtype=np.float64)
    assert Y.shape == (n, d)
    vvv = variant in (1, 3)
    tdist = variant >= 2
    nsnap = nrep // thin + 1
    nl = q if vvv else 1
    ptr, idx = csr_from_lists(df_j) if isinstance(df_j, list) else df_j
    yb = _colmajor(Y)
    ib = np.ascontiguousarray(init, dtype=np.int32)
    m0 = _f64(mu0)
    l0 = _colmajor(lambda0)
    ob = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    z = np.zeros((nsnap, n), dtype=np.int32)
    mu = np.zeros((nsnap, q * d))
    lam = np.zeros((nsnap, nl, d * d))
    w = np.zeros((nsnap, n)) if tdist else None
    pl = np.zeros(nrep)
    rc = lib().orc_iterate(variant, _p(yb, C.c_double), C.c_int64(n), d, q, _p(ptr, C.c_int64),
                           _p(idx, C.c_int32), nrep, thin, C.c_double(gamma), _p(ib, C.c_int32),
                           _p(m0, C.c_double), _p(l0, C.c_double), C.c_double(alpha), C.c_double(beta),
                           C.c_uint64(seed), _p(ob, C.c_int32), C.c_uint32(compat), _p(z, C.c_int32),
                           _p(mu, C.c_double), _p(lam, C.c_double), _p(w, C.c_double), _p(pl, C.c_double))
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    lam_m = lam.reshape(nsnap, nl, d, d).transpose(0, 1, 3, 2)   # col-major -> [i, j]
    out = {"z": z, "mu": mu, "lambda": lam_m if vvv else lam_m[:, 0], "plogLik": pl}
    if tdist:
        out["weights"] = w
    return out


def iterate(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """src/cluster.cpp:217-321"""
    return _run_cluster(0, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_vvv(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """src/cluster.cpp:323-444"""
    return _run_cluster(1, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_t(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """src/cluster.cpp:446-568"""
    return _run_cluster(2, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def iterate_t_vvv(Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw):
    """src/cluster.cpp:570-710"""
    return _run_cluster(3, Y, df_j, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, **kw)


def parse_neighbors(strings):
    """convert_neighbors (src/cluster.cpp:134-145): "1-based,comma" strings / None -> 0-based CSR."""
    n0 = len(strings)
    arr = (C.c_char_p * n0)(*[None if s is None else s.encode() for s in strings])
    ptr = np.zeros(n0 + 1, dtype=np.int64)
    rc = lib().orc_parse_neighbors(arr, C.c_int64(n0), _p(ptr, C.c_int64), None)
    if rc:
        raise ValueError("Unconvertable value")
    idx = np.zeros(max(int(ptr[-1]), 1), dtype=np.int32)
    lib().orc_parse_neighbors(arr, C.c_int64(n0), _p(ptr, C.c_int64), _p(idx, C.c_int32))
    return ptr, idx


def find_subspot_neighbors(spot_csr, positions, dist):
    """src/cluster.cpp:161-215"""
    sp, si = spot_csr
    n0 = len(sp) - 1
    pos = _colmajor(positions)
    n = np.asarray(positions).shape[0]
    ptr = np.zeros(n + 1, dtype=np.int64)
    cap = max(64 * n, 64)
    idx = np.zeros(cap, dtype=np.int32)
    rc = lib().orc_subspot_neighbors(_p(sp, C.c_int64), _p(si, C.c_int32), C.c_int64(n0),
                                     _p(pos, C.c_double), C.c_int64(n), C.c_double(dist),
                                     _p(ptr, C.c_int64), _p(idx, C.c_int32), C.c_int64(cap))
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    return ptr, idx[:max(int(ptr[-1]), 1)]


def lattice_neighbors(array_row, array_col, platform):
    """.find_neighbors (R/spatialCluster.R:227-302); platform 'Visium' | 'VisiumHD' | 'ST'."""
    if platform not in ("Visium", "VisiumHD", "ST"):
        raise ValueError('.find_neighbors: Unsupported platform "%s".' % platform)
    r = np.ascontiguousarray(array_row, dtype=np.int32)
    c = np.ascontiguousarray(array_col, dtype=np.int32)
    n = len(r)
    ptr = np.zeros(n + 1, dtype=np.int64)
    idx = np.zeros(max(6 * n, 1), dtype=np.int32)
    lib().orc_lattice_neighbors(_p(r, C.c_int32), _p(c, C.c_int32), C.c_int64(n),
                                0 if platform == "Visium" else 1, _p(ptr, C.c_int64), _p(idx, C.c_int32))
    return ptr, idx[:max(int(ptr[-1]), 1)]


def find_neighbors(positions, radius, method="manhattan"):
    """find_neighbors (R/utils.R:13-26) -> list of 0-based ascending neighbour arrays."""
    pos = _colmajor(positions)
    n = np.asarray(positions).shape[0]
    m = {"manhattan": 0, "euclidean": 1}[method]
    ptr = np.zeros(n + 1, dtype=np.int64)
    nnz = lib().orc_radius_neighbors(_p(pos, C.c_double), C.c_int64(n), C.c_double(radius), m,
                                     _p(ptr, C.c_int64), None)
    idx = np.zeros(max(int(nnz), 1), dtype=np.int32)
    lib().orc_radius_neighbors(_p(pos, C.c_double), C.c_int64(n), C.c_double(radius), m,
                               _p(ptr, C.c_int64), _p(idx, C.c_int32))
    return lists_from_csr(ptr, idx)


def adaptive_mcmc(acc, target, it, A, u):
    d = A.shape[0]
    ab, ub = _colmajor(A), _f64(u)
    out = np.zeros(d * d)
    rc = lib().orc_adaptive_mcmc(C.c_double(acc), C.c_double(target), C.c_int64(it), _p(ab, C.c_double),
                                 _p(ub, C.c_double), d, _p(out, C.c_double))
    if rc:
        raise np.linalg.LinAlgError("adaptive chol failed")
    return out.reshape(d, d, order="F")


def iterate_deconv(subspot_positions, dist, spot_neighbors, Y, tdist, nrep, thin, n, n0, d, gamma, q,
                   init, subspots, verbose, jitter_scale, adapt_before, c, mu0, lambda0, alpha, beta,
                   thread_num=1, seed=0, order=None, keep_Y=True):
    """src/cluster.cpp:742-1141.  spot_neighbors: list of "a,b,c" strings / None (NA).
    Y (n x d) is NOT mutated here (the reference mutates its argument, quirk Q10)."""
    sp, si = parse_neighbors(spot_neighbors)
    pos = _colmajor(subspot_positions)
    yb = _colmajor(Y).copy()
    ib = np.ascontiguousarray(init, dtype=np.int32)
    m0, l0 = _f64(mu0), _colmajor(lambda0)
    ob = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    nsnap = nrep // thin + 1
    z = np.zeros((nsnap, n), dtype=np.int32)
    mu = np.zeros((nsnap, q * d))
    lam = np.zeros((nsnap, d * d))
    w = np.zeros((nsnap, n))
    Yo = np.zeros((nsnap, d, n)) if keep_Y else None   # each snapshot col-major n x d
    ych, pl = np.zeros(nrep), np.zeros(nrep)
    dfj = np.zeros(4 * n, dtype=np.int64)
    rc = lib().orc_iterate_deconv(
        _p(pos, C.c_double), C.c_double(dist), _p(sp, C.c_int64), _p(si, C.c_int32), _p(yb, C.c_double),
        int(bool(tdist)), nrep, thin, C.c_int64(n), C.c_int64(n0), d, C.c_double(gamma), q,
        _p(ib, C.c_int32), subspots, C.c_double(jitter_scale), adapt_before, C.c_double(c),
        _p(m0, C.c_double), _p(l0, C.c_double), C.c_double(alpha), C.c_double(beta), thread_num,
        C.c_uint64(seed), _p(ob, C.c_int32), _p(z, C.c_int32), _p(mu, C.c_double), _p(lam, C.c_double),
        _p(w, C.c_double), _p(Yo, C.c_double), _p(ych, C.c_double), _p(pl, C.c_double), _p(dfj, C.c_int64))
    if rc:
        raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
    out = {"z": z, "mu": mu, "lambda": lam.reshape(nsnap, d, d).transpose(0, 2, 1), "weights": w,
           "Ychange": ych, "plogLik": pl, "df_j": dfj.reshape(4, n).T.copy()}
    if keep_Y:
        out["Y"] = Yo.transpose(0, 2, 1)   # nsnap x n x d
    return out


def mode_rows(zs):
    """R's Mode applied to every column of the kept label rows: apply(zs, 2, Mode) with
    Mode(x) = ux[which.max(tabulate(match(x, ux)))], ux = unique(x)   (R/utils.R:63-66, R/spatialCluster.R:204-213):
    the most frequent label; among equally frequent ones the label that appears FIRST in the column."""
    zs = np.asarray(zs)
    out = np.zeros(zs.shape[1], dtype=np.int32)
    for j in range(zs.shape[1]):
        x = zs[:, j]
        ux = []
        for v in x:                      # unique(): order of first appearance
            if v not in ux:
                ux.append(v)
        counts = [int(np.sum(x == v)) for v in ux]    # tabulate(match(x, ux))
        out[j] = ux[int(np.argmax(counts))]           # which.max: first maximum
    return out
