"""Host-side mirror of the two helpers of src/compare_enhanced.cpp (internal functions of the reference,
R/spatialEnhance.R:570-585), bound to libbsb200.so.  Same names, argument meaning and returned matrices."""
import ctypes as C

import numpy as np

from ._lib import BayesSpaceError, ERR_INVALID, check, colmajor, lib, ptr


def map_subspot2ref(subspot_coords, ref_coords, thread_num=1, device=0):
    """map_subspot2ref (src/compare_enhanced.cpp:47-92): matrix with one row (subspot i, reference j, squared distance)
    per nearest reference spot of every subspot, 0-based ids as in the reference; exact ties give several rows."""
    s = np.asarray(subspot_coords, dtype=np.float64)
    r = np.asarray(ref_coords, dtype=np.float64)
    if s.ndim != 2 or r.ndim != 2 or s.shape[1] != r.shape[1]:
        raise BayesSpaceError(ERR_INVALID, "Invalid arguments!")
    ns, k = s.shape
    nr = r.shape[0]
    sb, rb = colmajor(s), colmajor(r)
    row_ptr = np.zeros(ns + 1, dtype=np.int64)
    best = np.zeros(ns)
    nrows = C.c_int64(0)
    L = lib()
    check(L.bsb200_map_subspot2ref(device, ptr(sb), C.c_int64(ns), ptr(rb), C.c_int64(nr), k, ptr(row_ptr), None,
                                   C.c_int64(0), ptr(best), C.byref(nrows)))
    idx = np.zeros(max(int(nrows.value), 1), dtype=np.int32)
    check(L.bsb200_map_subspot2ref(device, ptr(sb), C.c_int64(ns), ptr(rb), C.c_int64(nr), k, ptr(row_ptr), ptr(idx),
                                   C.c_int64(len(idx)), ptr(best), C.byref(nrows)))
    counts = np.diff(row_ptr)
    out = np.zeros((int(nrows.value), 3))
    out[:, 0] = np.repeat(np.arange(ns), counts)
    out[:, 1] = idx[:int(nrows.value)]
    out[:, 2] = np.repeat(best, counts)
    return out


def compute_corr(m1, m2, thread_num=1, device=0):
    """compute_corr (src/compare_enhanced.cpp:94-132): matrix (row index, Pearson correlation of the two rows)."""
    a = np.asarray(m1, dtype=np.float64)
    b = np.asarray(m2, dtype=np.float64)
    if a.shape != b.shape or a.ndim != 2:
        raise BayesSpaceError(ERR_INVALID, "Input matrices should be of the same shape.")
    nrow, ncol = a.shape
    ab, bb = colmajor(a), colmajor(b)
    cor = np.zeros(nrow)
    check(lib().bsb200_compute_corr(device, ptr(ab), ptr(bb), C.c_int64(nrow), C.c_int64(ncol), ptr(cor)))
    return np.c_[np.arange(nrow, dtype=np.float64), cor]
