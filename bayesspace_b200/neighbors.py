"""Neighbour construction behind the samplers: thin ctypes mirrors of the reference's helpers.

Same names, argument meaning and error behaviour as the reference (file:line under /root/reference):
  find_neighbors          R/utils.R:13-26
  _find_neighbors         R/spatialCluster.R:227-302   (.find_neighbors)
  convert_neighbors       src/cluster.cpp:134-145
  find_subspot_neighbors  src/cluster.cpp:161-215
  convert_neighbors2mtx   src/cluster.cpp:147-159
All work happens in libbsb200.so (csrc/graph_host.cpp); adjacency is CSR (int64 ptr, int32 idx).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, colmajor, i32, lib, ptr

PLATFORMS = {"Visium": 0, "VisiumHD": 1, "ST": 1}


def _two_pass(call, n):
    """Runs a CSR-producing ABI call with the two-pass sizing protocol."""
    ptr_arr = np.zeros(n + 1, dtype=np.int64)
    nnz = C.c_int64(0)
    check(call(ptr(ptr_arr), None, C.c_int64(0), C.byref(nnz)))
    idx = np.zeros(max(int(nnz.value), 1), dtype=np.int32)
    check(call(ptr(ptr_arr), ptr(idx), C.c_int64(int(nnz.value)), C.byref(nnz)))
    return ptr_arr, idx[:int(nnz.value)] if nnz.value else idx[:0]


def csr_to_lists(csr):
    p, idx = csr
    return [idx[p[j]:p[j + 1]].copy() for j in range(len(p) - 1)]


def lists_to_csr(df_j):
    p = np.zeros(len(df_j) + 1, dtype=np.int64)
    for j, v in enumerate(df_j):
        p[j + 1] = p[j] + len(v)
    idx = np.zeros(int(p[-1]), dtype=np.int32)
    for j, v in enumerate(df_j):
        idx[p[j]:p[j + 1]] = np.asarray(v, dtype=np.int32)
    return p, idx


def as_csr(df_j):
    if isinstance(df_j, tuple):
        return np.ascontiguousarray(df_j[0], dtype=np.int64), i32(df_j[1])
    return lists_to_csr(df_j)


def find_neighbors(positions, radius, method="manhattan"):
    """List df_j with df_j[i] = 0-based ascending neighbours of i: 0 < dist <= radius."""
    if method not in ("manhattan", "euclidean"):
        raise ValueError("'arg' should be one of \"manhattan\", \"euclidean\"")
    pos = np.asarray(positions, dtype=np.float64)
    n = pos.shape[0]
    buf = colmajor(pos)
    m = 0 if method == "manhattan" else 1
    L = lib()
    csr = _two_pass(lambda p, i, cap, nnz: L.bsb200_neighbors_radius(ptr(buf), C.c_int64(n), C.c_double(radius),
                                                                     m, p, i, cap, nnz), n)
    return csr_to_lists(csr)


def _find_neighbors(array_row, array_col, platform, as_lists=False):
    """.find_neighbors: returns (spot_neighbors strings, df_j) like the reference returns
    (sce$spot.neighbors, df_j); df_j is CSR unless as_lists."""
    if platform not in PLATFORMS:
        raise ValueError('.find_neighbors: Unsupported platform "%s".' % platform)
    r, c = i32(array_row), i32(array_col)
    n = r.shape[0]
    L = lib()
    csr = _two_pass(lambda p, i, cap, nnz: L.bsb200_neighbors_lattice(ptr(r), ptr(c), C.c_int64(n),
                                                                      PLATFORMS[platform], p, i, cap, nnz), n)
    p, idx = csr
    strings = [None if p[j] == p[j + 1] else ",".join(str(int(v) + 1) for v in idx[p[j]:p[j + 1]])
               for j in range(n)] if n <= 200000 else None
    return strings, (csr_to_lists(csr) if as_lists else csr)


def convert_neighbors(spot_neighbors):
    """"1-based,comma" strings / None (NA) -> 0-based CSR."""
    n0 = len(spot_neighbors)
    arr = (C.c_char_p * n0)(*[None if s is None else str(s).encode() for s in spot_neighbors])
    L = lib()
    return _two_pass(lambda p, i, cap, nnz: L.bsb200_parse_neighbor_strings(arr, C.c_int64(n0), p, i, cap, nnz), n0)


def find_subspot_neighbors(spot_csr, subspot_positions, dist):
    sp, si = as_csr(spot_csr)
    n0 = len(sp) - 1
    pos = np.asarray(subspot_positions, dtype=np.float64)
    n = pos.shape[0]
    buf = colmajor(pos)
    L = lib()
    return _two_pass(lambda p, i, cap, nnz: L.bsb200_subspot_neighbors(ptr(buf), C.c_int64(n), C.c_int64(n0), ptr(sp),
                                                                       ptr(si), C.c_double(dist), p, i, cap, nnz), n)


def convert_neighbors2mtx(csr):
    p, idx = as_csr(csr)
    n = len(p) - 1
    out = np.zeros(4 * n, dtype=np.int64)
    check(lib().bsb200_neighbors_to_matrix4(ptr(p), ptr(idx), C.c_int64(n), ptr(out)))
    return out.reshape(4, n).T.copy()


def colour_graph(csr):
    p, idx = as_csr(csr)
    n = len(p) - 1
    col = np.zeros(n, dtype=np.int32)
    nc = C.c_int32(0)
    check(lib().bsb200_colour_graph(ptr(p), ptr(idx), C.c_int64(n), ptr(col), C.byref(nc)))
    return col, int(nc.value)


def colour_lattice(array_row, array_col, platform):
    if platform not in PLATFORMS:
        raise ValueError('.find_neighbors: Unsupported platform "%s".' % platform)
    r, c = i32(array_row), i32(array_col)
    col = np.zeros(r.shape[0], dtype=np.int32)
    nc = C.c_int32(0)
    check(lib().bsb200_colour_lattice(ptr(r), ptr(c), C.c_int64(r.shape[0]), PLATFORMS[platform], ptr(col), C.byref(nc)))
    return col, int(nc.value)
