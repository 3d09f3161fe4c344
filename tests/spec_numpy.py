"""A second, independent restatement of the four cluster samplers, written in numpy directly from the reference's
lines (/root/reference/src/cluster.cpp) with numpy.linalg doing what Armadillo does there -- `inv`, `chol` per spot,
exactly the reference's cost profile, none of the algebraic shortcuts of the CUDA path and none of oracle/ref_cpu.cpp's
code.  Test infrastructure: tests/test_oracle.py runs it beside the C++ oracle on the same keyed variates; the two
must produce the same chain.  (The variates themselves come from the oracle's Philox primitives, which are pinned
separately by Random123 known answers and distribution tests.)

    iterate        src/cluster.cpp:217-321      iterate_vvv    :323-444
    iterate_t      :446-568                     iterate_t_vvv  :570-710
    dmvnrm_arma_fast :83-106, inplace_tri_mat_mult_t :68-78
    rmvnorm / rwish (RcppDist), sample (Rcpp sugar): SURVEY.md App. C
"""
import numpy as np

P_MU, P_WISH_NORM, P_WISH_CHI, P_SPOT_U, P_SPOT_GAMMA = 1, 2, 3, 4, 5
LOG2PI = np.log(2.0 * np.pi)


def chol_upper(S):
    """arma::chol: upper R with R^T R = S."""
    return np.linalg.cholesky(S).T


def dmvnrm_arma_fast(x, mean, sigma):
    """:83-106 with logd = true, for one row x.  rooti = inv(trimatu(chol(sigma))); z = rooti * (x - mean) through
    inplace_tri_mat_mult_t (row i of the UPPER triangle times z[i:]) -- the reference's form (quirk Q1)."""
    d = x.shape[0]
    rooti = np.linalg.inv(np.triu(chol_upper(sigma)))
    z = x - mean
    for i in range(d):                       # :71-76, in place: entries j >= i are still the old ones
        z[i] = np.dot(rooti[i, i:], z[i:])
    return np.sum(np.log(np.diag(rooti))) - 0.5 * d * LOG2PI - 0.5 * np.dot(z, z)


def rmvnorm1(O, seed, index, it, mean, var):
    """RcppDist rmvnorm(1, mean, var) = z chol(var) + mean, z ~ N(0, I) row."""
    z = O.normal(seed, P_MU, index, it, 0, mean.shape[0])
    return z @ chol_upper(var) + mean


def rwish(O, seed, index, it, df, S):
    """RcppDist rwish(int df, S): Bartlett factor A (lower; N(0,1) below the diagonal, sqrt(chisq(df - i)) on it),
    result (A^T chol(S))^T (A^T chol(S))."""
    d = S.shape[0]
    df = int(df)                              # the C++ parameter is an int (quirk Q4)
    A = np.zeros((d, d))
    for i in range(1, d):
        A[i, :i] = O.normal(seed, P_WISH_NORM, index, it, i * (i - 1) // 2, i)
    for i in range(d):
        A[i, i] = np.sqrt(O.gamma(seed, P_WISH_CHI, index * 1024 + i, 1, it, (df - i) / 2.0, 2.0)[0])
    B = A.T @ chol_upper(S)
    return B.T @ B


def sample_without(q, z_prev, u):
    """sample(qvec[qvec != z_prev], 1): element floor(len * u) of the remaining labels, ascending."""
    rest = [k for k in range(1, q + 1) if k != z_prev]
    return rest[min(int(np.floor(len(rest) * u)), len(rest) - 1)]


def sample_without_bits(q, z_prev, bits):
    """The same with the uniform given as 32 random bits: u = bits / 2^32, in exact integer arithmetic."""
    rest = [k for k in range(1, q + 1) if k != z_prev]
    return rest[(int(bits) * len(rest)) >> 32]


def sample_two(z_prev, z_new, p, u):
    """sample({z_prev, z_new}, 1, TRUE, {1 - p, p}): Rcpp sugar sorts the probabilities in DECREASING order and
    inverts the cumulative sum with one uniform."""
    items = [(1.0 - p, z_prev), (p, z_new)]
    if p >= 0.5:                              # stable descending order: z_new first when p >= 1 - p ... ties: p == .5
        items = [(p, z_new), (1.0 - p, z_prev)]
    return items[0][1] if u <= items[0][0] else items[1][1]


def run(O, variant, Y, nbrs, nrep, thin, n, d, gamma, q, init, mu0, lambda0, alpha, beta, seed, order=None):
    """variant: 'iterate' | 'iterate_vvv' | 'iterate_t' | 'iterate_t_vvv'.  nbrs: list of 0-based neighbour arrays.
    order: visiting order of the z / w scan; None = the reference's j = 0..n-1 (the only departure from the reference
    this file knows: the CUDA path visits colour class by colour class)."""
    tdist, vvv = variant in ("iterate_t", "iterate_t_vvv"), variant.endswith("vvv")
    nsnap = nrep // thin + 1
    zs = np.zeros((nsnap, n), dtype=np.int64)
    mus = np.zeros((nsnap, q * d))
    lams = [None] * nsnap
    ws = np.zeros((nsnap, n))
    plogLik = np.full(nrep, np.nan)
    z = np.array(init, dtype=np.int64)
    w = np.ones(n)
    lam = [lambda0.copy() for _ in range(q)] if vvv else [lambda0.copy()]
    sig = [np.linalg.inv(lambda0) for _ in lam]
    zs[0], mus[0], ws[0] = z, np.tile(mu0, q), w
    lams[0] = [m.copy() for m in lam] if vvv else lam[0].copy()
    V = beta * np.eye(d)
    for i in range(1, nrep):
        # ---- mu (:247-258, :482-492)
        mu = np.zeros((q, d))
        for k in range(1, q + 1):
            idx = np.where(z == k)[0]
            wk = w[idx] if tdist else np.ones(len(idx))
            n_i = wk.sum() if tdist else len(idx)
            ysum = (Y[idx] * wk[:, None]).sum(axis=0)
            L = lam[k - 1] if vvv else lam[0]
            var = np.linalg.inv(lambda0 + n_i * L)
            mean = var @ (lambda0 @ mu0 + L @ ysum)
            mu[k - 1] = rmvnorm1(O, seed, k - 1, i, mean, var)
        # ---- lambda (:260-270, :382-390, :500-505, :635-644)
        mu_long = mu[z - 1]
        R = Y - mu_long
        if vvv:
            for k in range(1, q + 1):
                idx = np.where(z == k)[0]
                wk = w[idx] if tdist else np.ones(len(idx))
                S = (R[idx] * wk[:, None]).T @ R[idx]
                lam[k - 1] = rwish(O, seed, k - 1, i, len(idx) + alpha, np.linalg.inv(V + S))
                sig[k - 1] = np.linalg.inv(lam[k - 1])
        else:
            S = (R * (w if tdist else np.ones(n))[:, None]).T @ R
            lam[0] = rwish(O, seed, 0, i, n + alpha, np.linalg.inv(V + S))
            sig[0] = np.linalg.inv(lam[0])
        # ---- z (and w) scan (:272-306, :392-430, :507-553, :646-695)
        w_alpha = float((d + 4) // 2)             # integer division in the reference (quirk Q2)
        plj = np.zeros(n)
        for j in (range(n) if order is None else order):
            z_prev = int(z[j])
            kp = z_prev - 1 if vvv else 0
            if tdist:
                r = Y[j] - mu_long[j]
                w_beta = 2.0 / (r @ lam[kp] @ r + 4.0)
                w[j] = O.gamma_spot(seed, j, i, w_alpha, w_beta)
            prop_bits, u_acc, _ = O.spot_draw(seed, j, i)
            z_new = sample_without_bits(q, z_prev, prop_bits)
            kn = z_new - 1 if vvv else 0
            scale = w[j] if tdist else 1.0
            h_prev = dmvnrm_arma_fast(Y[j].copy(), mu[z_prev - 1], sig[kp] / scale)
            h_new = dmvnrm_arma_fast(Y[j].copy(), mu[z_new - 1], sig[kn] / scale)
            nb = np.asarray(nbrs[j], dtype=np.int64)
            if nb.size:
                if variant == "iterate_vvv":      # :404 compares the neighbour INDICES with the label (quirk Q3)
                    h_prev += gamma / nb.size * 2 * np.sum(nb == z_prev)
                else:
                    h_prev += gamma / nb.size * 2 * np.sum(z[nb] == z_prev)
                h_new += gamma / nb.size * 2 * np.sum(z[nb] == z_new)
            p = min(np.exp(h_new - h_prev), 1.0)
            z[j] = sample_two(z_prev, z_new, p, u_acc)
            plj[j] = h_prev
        plogLik[i] = plj.sum()
        if (i + 1) % thin == 0:
            r_ = (i + 1) // thin
            zs[r_], mus[r_], ws[r_] = z, mu.reshape(-1), w
            lams[r_] = [m.copy() for m in lam] if vvv else lam[0].copy()
    return {"z": zs, "mu": mus, "lambda": lams, "weights": ws, "plogLik": plogLik}


# ---- iterate_deconv (src/cluster.cpp:742-1141), adaptive_mcmc (:49-65), dmvnrm_prec_arma_fast (:109-132) -----
P_ERR, P_SPOT_YACC = 6, 7


def dmvnrm_prec_arma_fast(x, mean, lam):
    """:109-132 with logd = true, one row: rooti = trimatu(chol(lambda)); z = rooti (x - mean)."""
    d = x.shape[0]
    rooti = np.triu(chol_upper(lam))
    z = x - mean
    for i in range(d):
        z[i] = np.dot(rooti[i, i:], z[i:])
    return np.sum(np.log(np.diag(rooti))) - 0.5 * d * LOG2PI - 0.5 * np.dot(z, z)


def adaptive_mcmc(acc, target, it, A, u):
    """:49-65  chol(A (I + eta (acc - target) u^T u / |u|^2) A^T, "lower"), eta = min(1, d it^(-2/3))."""
    d = u.shape[0]
    eta = min(1.0, d * float(it) ** (-2.0 / 3.0))
    M = A @ (np.eye(d) + eta * (acc - target) * np.outer(u, u) / np.dot(u, u)) @ A.T
    return np.linalg.cholesky(M)


def run_deconv(O, Y, sub_nbrs, tdist, nrep, thin, n, n0, d, gamma, q, init, S, jitter_scale, adapt_before, c,
               mu0, lambda0, alpha, beta, seed, order=None):
    """Subspot s = r * n0 + j0.  sub_nbrs: 0-based neighbour arrays of the subspots.  order: visiting order of the
    parent spots in the sequential accept / w / z section (None = the reference's j0 = 0..n0-1)."""
    Y = np.array(Y, dtype=np.float64)          # the reference mutates the caller's matrix; here a private copy
    Y0 = Y[:n0].copy()                          # :777
    nsnap = nrep // thin + 1
    zs = np.zeros((nsnap, n), dtype=np.int64)
    mus = np.zeros((nsnap, q * d))
    lams, Ys = [None] * nsnap, [None] * nsnap
    ws = np.zeros((nsnap, n))
    Ychange, plogLik = np.full(nrep, np.nan), np.full(nrep, np.nan)
    z = np.array(init, dtype=np.int64)
    w = np.ones(n)
    lam = lambda0.copy()
    zs[0], mus[0], ws[0], lams[0], Ys[0] = z, np.tile(mu0, q), w, lam.copy(), Y.copy()
    V = beta * np.eye(d)
    w_alpha = float((d + 4) // 2)               # :809, integer division
    jitter_scale = jitter_scale / d             # :814
    adaptive = jitter_scale == 0.0
    A = [np.eye(d) for _ in range(n)] if adaptive else None
    acc_n, rej_n = np.zeros(n0, dtype=np.int64), np.zeros(n0, dtype=np.int64)
    rows = np.arange(S) * n0
    for i in range(1, nrep):
        # ---- mu, lambda, jitter draws (:880-908)
        mu = np.zeros((q, d))
        for k in range(1, q + 1):
            idx = np.where(z == k)[0]
            ysum = (Y[idx] * w[idx][:, None]).sum(axis=0)
            var = np.linalg.inv(lambda0 + w[idx].sum() * lam)
            mean = var @ (lambda0 @ mu0 + lam @ ysum)
            mu[k - 1] = rmvnorm1(O, seed, k - 1, i, mean, var)
        mu_long = mu[z - 1]
        R = Y - mu_long
        lam = rwish(O, seed, 0, i, n + alpha, np.linalg.inv(V + (R * w[:, None]).T @ R))
        sd = 1.0 if adaptive else np.sqrt(jitter_scale)
        error = np.stack([O.normal(seed, P_ERR, s, i, 0, d) for s in range(n)]) * sd
        # ---- proposals and their joint acceptance probability per parent spot (:913-962)
        Y_new = np.zeros_like(Y)
        accp = np.zeros(n0)
        for j0 in range(n0):
            ss = rows + j0
            e = error[ss].copy()
            if adaptive:
                for r in range(S):
                    e[r] = A[ss[r]] @ e[r]
            e -= e.sum(axis=0) / S                  # zero-sum jitter: the parent's mean is preserved
            Yp, Yn = Y[ss], Y[ss] + e
            p_prev = p_new = 0.0
            for r in range(S):
                L = lam * w[ss[r]]
                p_prev += dmvnrm_prec_arma_fast(Yp[r].copy(), mu_long[ss[r]], L) - c * np.sum((Yp[r] - Y0[j0]) ** 2)
                p_new += dmvnrm_prec_arma_fast(Yn[r].copy(), mu_long[ss[r]], L) - c * np.sum((Yn[r] - Y0[j0]) ** 2)
            accp[j0] = min(np.exp(p_new - p_prev), 1.0)
            Y_new[ss] = Yn
        # ---- accept / reject, RAM update, w, z (:969-1053)
        updates, ll = 0, np.zeros(n)
        for j0 in (range(n0) if order is None else order):
            ss = rows + j0
            u_y = O.uniform2(seed, P_SPOT_YACC, j0, i, 0)[0]
            if sample_two(0, 1, accp[j0], u_y) == 1:
                Y[ss] = Y_new[ss]
                acc_n[j0] += 1
                updates += 1
            else:
                rej_n[j0] += 1
            for r in range(S):
                s = ss[r]
                if adaptive and i > 10 and (adapt_before == 0 or i <= adapt_before):
                    A[s] = adaptive_mcmc(acc_n[j0] / float(acc_n[j0] + rej_n[j0]), 0.234, i, A[s], error[s])
                if tdist:
                    rr = Y[s] - mu_long[s]
                    w[s] = O.gamma(seed, P_SPOT_GAMMA, s, 1, i, w_alpha, 2.0 / (rr @ lam @ rr + 4.0))[0]
                z_prev = int(z[s])
                u_prop, u_acc = O.uniform2(seed, P_SPOT_U, s, i, 0)
                z_new = sample_without(q, z_prev, u_prop)
                l_prev = dmvnrm_prec_arma_fast(Y[s].copy(), mu[z_prev - 1], lam * w[s])
                l_new = dmvnrm_prec_arma_fast(Y[s].copy(), mu[z_new - 1], lam * w[s])
                h_prev, h_new = l_prev, l_new
                nb = np.asarray(sub_nbrs[s], dtype=np.int64)
                if nb.size:
                    h_prev += gamma / nb.size * 2 * np.sum(z[nb] == z_prev)
                    h_new += gamma / nb.size * 2 * np.sum(z[nb] == z_new)
                p = np.exp(h_new - h_prev)
                yes = 1 if p >= 1 else sample_two(0, 1, p, u_acc)     # :1037-1046: no draw when p >= 1
                if yes == 1:
                    z[s] = z_new
                ll[s] = l_new if yes == 1 else l_prev                 # likelihood only, of the resulting state
        Ychange[i] = updates / float(n0)
        plogLik[i] = ll.sum()
        if (i + 1) % thin == 0:
            r_ = (i + 1) // thin
            zs[r_], mus[r_], ws[r_], lams[r_], Ys[r_] = z, mu.reshape(-1), w, lam.copy(), Y.copy()
    return {"z": zs, "mu": mus, "lambda": lams, "weights": ws, "Y": Ys, "Ychange": Ychange, "plogLik": plogLik}


# ---- neighbour construction, restated from the R / C++ sources with pandas / numpy --------------------------
def find_neighbors_lattice(array_row, array_col, platform):
    """.find_neighbors (R/spatialCluster.R:227-302): cross-join the spots with the platform's offsets, left-merge the
    candidate coordinates with the spot table, order by (primary, neighbour), split by primary, drop the NAs.
    Returns the 0-based lists (df_j) and the 1-based comma strings / None (sce$spot.neighbors)."""
    import pandas as pd
    if platform == "Visium":
        offsets = pd.DataFrame({"x.offset": [-2, 2, -1, 1, -1, 1], "y.offset": [0, 0, -1, -1, 1, 1]})
    elif platform in ("VisiumHD", "ST"):
        offsets = pd.DataFrame({"x.offset": [0, 1, 0, -1], "y.offset": [-1, 0, 1, 0]})
    else:
        raise ValueError('.find_neighbors: Unsupported platform "%s".' % platform)
    n = len(array_row)
    spots = pd.DataFrame({"spot.idx": np.arange(1, n + 1), "array_col": np.asarray(array_col, dtype=np.int64),
                          "array_row": np.asarray(array_row, dtype=np.int64)})
    cand = spots.merge(offsets, how="cross")
    cand["x.pos"] = cand["array_col"] + cand["x.offset"]
    cand["y.pos"] = cand["array_row"] + cand["y.offset"]
    nb = cand.merge(spots, how="left", left_on=["x.pos", "y.pos"], right_on=["array_col", "array_row"],
                    suffixes=(".primary", ".neighbor"))
    nb = nb.sort_values(["spot.idx.primary", "spot.idx.neighbor"], kind="stable")
    df_j, strings = [], []
    groups = {k: v["spot.idx.neighbor"].dropna().astype(np.int64).to_numpy() for k, v in nb.groupby("spot.idx.primary")}
    for j in range(1, n + 1):
        x = groups.get(j, np.zeros(0, dtype=np.int64))
        strings.append(None if len(x) == 0 else ",".join(str(int(v)) for v in x))
        df_j.append(x - 1)
    return df_j, strings


def find_neighbors_radius(positions, radius, method="manhattan"):
    """find_neighbors (R/utils.R:13-26): dense distance matrix, 0 < dist <= radius, which() order (ascending)."""
    P = np.asarray(positions, dtype=np.float64)
    diff = P[:, None, :] - P[None, :, :]
    dist = np.abs(diff).sum(axis=2) if method == "manhattan" else np.sqrt((diff ** 2).sum(axis=2))
    keep = (dist <= radius) & (dist > 0)
    return [np.where(keep[i])[0] for i in range(P.shape[0])]


def find_subspot_neighbors(spot_nbrs, positions, dist):
    """src/cluster.cpp:161-215.  spot_nbrs: 0-based neighbour arrays of the n0 spots; subspot s = r * n0 + spot.
    Candidates: for every candidate spot (the spot's neighbours in list order, then the spot itself) its S subspots
    r = 0..S-1 (the column-major flattening of the S x (k+1) index matrix); kept where 1e-6 < Manhattan < dist."""
    P = np.asarray(positions, dtype=np.float64)
    n0 = len(spot_nbrs)
    S = P.shape[0] // n0
    if abs(P.shape[0] / n0 - S) > 1e-6:
        raise RuntimeError("Invalid arguments!")
    out = []
    for s in range(P.shape[0]):
        spot = s % n0
        cand_spots = list(spot_nbrs[spot]) + [spot]
        cand = np.array([r * n0 + c for c in cand_spots for r in range(S)], dtype=np.int64)
        man = np.abs(P[cand] - P[s]).sum(axis=1)
        keep = (man < dist) & (man > 1e-6)
        if keep.sum() < 2:
            raise RuntimeError("Error in finding neighbors of subspots!")
        out.append(cand[keep])
    return out
