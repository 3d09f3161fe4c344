"""Chain persistence (R/mcmcChain.R:48-73, :181-223) without libhdf5: the samplers stream their snapshot rows into a
directory of raw files + meta.json (bayesspace_b200/csrc/chain_file.cpp).  `chain_file(dir)` wraps a sampler call,
`read_chain(dir)` maps the files back to the matrices .clean_chain builds (one row per kept iteration, columns =
the parameter flattened row-major, "dims" = its per-iteration shape)."""
import contextlib
import json
import os

import numpy as np

from ._lib import check, lib


@contextlib.contextmanager
def chain_file(directory):
    """with chain_file("/path/chain"): iterate_deconv(..., keep_Y=False)  -- rows are written as they are produced."""
    check(lib().bsb200_chain_file_open(os.fsencode(directory)))
    try:
        yield directory
    finally:
        check(lib().bsb200_chain_file_close())


def read_chain(directory):
    """dict param -> (rows x cols array, dims); memory-mapped, so a 7 GB Y dataset is not loaded unless it is touched."""
    meta = json.load(open(os.path.join(directory, "meta.json")))
    assert meta["format"] == "bsb200-chain-1"
    out = {}
    for name, p in meta["params"].items():
        dt = np.float64 if p["dtype"] == "float64" else np.int32
        a = np.memmap(os.path.join(directory, p["file"]), dtype=dt, mode="r", shape=(p["rows"], p["cols"]))
        out[name] = (a, tuple(p["dims"]))
    return out
