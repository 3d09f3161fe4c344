"""ctypes loader of libbsb200.so (the C ABI in include/bsb200.h).

There is no fallback of any kind: if the library has not been built, importing a sampler raises,
and if no sm_100 device is present every sampler / probe call raises BayesSpaceError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsb200.so")

OK = 0
ERR_INVALID, ERR_NEIGHBORS, ERR_PARSE, ERR_BUFFER, ERR_CUDA, ERR_NUMERIC, ERR_UNSUPPORTED, ERR_CANCELLED = \
    -1, -2, -3, -4, -5, -6, -7, -8
COMPAT_Q1, COMPAT_Q3 = 1, 4
COMPAT_REFERENCE = COMPAT_Q1 | COMPAT_Q3


class BayesSpaceError(RuntimeError):
    """Raised with the message the reference would have raised as an R error."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class ClusterArgs(C.Structure):
    _fields_ = [("model", C.c_int32), ("precision", C.c_int32), ("n", C.c_int64), ("d", C.c_int32),
                ("q", C.c_int32), ("nrep", C.c_int32), ("thin", C.c_int32), ("gamma", C.c_double),
                ("alpha", C.c_double), ("beta", C.c_double), ("Y", C.c_void_p), ("nbr_ptr", C.c_void_p),
                ("nbr_idx", C.c_void_p), ("init", C.c_void_p), ("mu0", C.c_void_p), ("lambda0", C.c_void_p),
                ("seed", C.c_uint64), ("compat", C.c_uint32), ("device", C.c_int32), ("colour", C.c_void_p),
                ("ncolours", C.c_int32), ("groups", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32),
                ("comm_id", C.c_void_p)]


class ClusterOut(C.Structure):
    _fields_ = [("z", C.c_void_p), ("mu", C.c_void_p), ("lambda_", C.c_void_p), ("weights", C.c_void_p),
                ("plogLik", C.c_void_p), ("rows_done", C.c_int32), ("ncolours", C.c_int32),
                ("setup_ms", C.c_double), ("device_ms", C.c_double), ("own_lo", C.c_int64), ("own_hi", C.c_int64),
                ("z_mode", C.c_void_p), ("mode_from", C.c_int32)]


class DeconvArgs(C.Structure):
    _fields_ = [("subspot_positions", C.c_void_p), ("dist", C.c_double), ("spot_nbr_ptr", C.c_void_p),
                ("spot_nbr_idx", C.c_void_p), ("Y", C.c_void_p), ("tdist", C.c_int32), ("nrep", C.c_int32),
                ("thin", C.c_int32), ("n", C.c_int64), ("n0", C.c_int64), ("d", C.c_int32), ("gamma", C.c_double),
                ("q", C.c_int32), ("init", C.c_void_p), ("subspots", C.c_int32), ("verbose", C.c_int32),
                ("jitter_scale", C.c_double), ("adapt_before", C.c_int32), ("c", C.c_double), ("mu0", C.c_void_p),
                ("lambda0", C.c_void_p), ("alpha", C.c_double), ("beta", C.c_double), ("thread_num", C.c_int32),
                ("seed", C.c_uint64), ("device", C.c_int32), ("reserved", C.c_int32)]


class DeconvOut(C.Structure):
    _fields_ = [("z", C.c_void_p), ("mu", C.c_void_p), ("lambda_", C.c_void_p), ("weights", C.c_void_p),
                ("Y", C.c_void_p), ("Y_mean", C.c_void_p), ("mean_from", C.c_int32), ("mode_from", C.c_int32),
                ("z_mode", C.c_void_p), ("Ychange", C.c_void_p),
                ("plogLik", C.c_void_p), ("df_j", C.c_void_p), ("rows_done", C.c_int32), ("ncolours", C.c_int32),
                ("setup_ms", C.c_double), ("device_ms", C.c_double)]


# every symbol include/bsb200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "bsb200_version", "bsb200_last_error", "bsb200_cancel", "bsb200_set_interrupt_poll", "bsb200_device_count",
    "bsb200_neighbors_lattice", "bsb200_neighbors_radius", "bsb200_parse_neighbor_strings",
    "bsb200_subspot_neighbors", "bsb200_neighbors_to_matrix4", "bsb200_colour_graph", "bsb200_colour_lattice",
    "bsb200_cluster", "bsb200_comm_unique_id", "bsb200_shard_plan", "bsb200_chain_create", "bsb200_chain_run", "bsb200_chain_profile", "bsb200_chain_state",
    "bsb200_chain_destroy", "bsb200_kernel_launches", "bsb200_release_memory", "bsb200_eval_loglik", "bsb200_eval_suffstats",
    "bsb200_chain_dry_sweep", "bsb200_debug_rng", "bsb200_debug_rendezvous", "bsb200_deconv",
    "bsb200_set_snapshot_sink", "bsb200_chain_file_open", "bsb200_chain_file_close", "bsb200_map_subspot2ref", "bsb200_compute_corr", "bsb200_init_cluster",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libbsb200.so has not been built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C bayesspace_b200/csrc -j8`); bayesspace_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.bsb200_last_error.restype = C.c_char_p
        L.bsb200_kernel_launches.restype = C.c_int64
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise BayesSpaceError(rc, lib().bsb200_last_error().decode())


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def colmajor(a):
    """2-D array -> flat float64 buffer in R's column-major layout (a view when `a` already is F-ordered)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2 and a.flags.f_contiguous:
        return a.reshape(-1, order="F")
    return np.asfortranarray(a).ravel(order="F")


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count():
    return int(lib().bsb200_device_count())


def kernel_launches():
    return int(lib().bsb200_kernel_launches())


def release_memory(device=0):
    """Hand the device and page-locked memory cached from destroyed chains back to the driver."""
    check(lib().bsb200_release_memory(int(device)))
