"""Generates tests/golden/golden_chains.npz + golden_meta.json from the CPU oracle.

The reference (R + Rcpp) cannot run in this image, so these vectors are NOT reference outputs; they pin
the oracle (and, through the GPU tests, the CUDA path) against drift.  Visiting order = colour classes
in sequence (closed-form lattice colouring), i.e. the scan the CUDA path performs.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from bayesspace_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def lattice_colour(w):
    if w.platform == "Visium":
        return (((w.array_col.astype(np.int64) - 3 * w.array_row.astype(np.int64)) // 2) % 3).astype(np.int32)
    return ((w.array_row.astype(np.int64) + w.array_col) % 2).astype(np.int32)


def main():
    cases, arrays = [], {}
    for workload in ("tiny_hex", "small_sq"):
        w = synth.make_workload(workload)
        g = O.lattice_neighbors(w.array_row, w.array_col, w.platform)
        order = np.lexsort((np.arange(w.n), lattice_colour(w))).astype(np.int32)
        for sampler in ("iterate", "iterate_vvv", "iterate_t", "iterate_t_vvv"):
            key = "%s_%s" % (workload, sampler)
            case = dict(key=key, workload=workload, sampler=sampler, nrep=25, thin=5, seed=20261020)
            out = getattr(O, sampler)(w.Y, g, case["nrep"], case["thin"], w.n, w.d, w.gamma, w.q, w.init, w.mu0,
                                      w.lambda0, w.alpha, w.beta, seed=case["seed"], order=order)
            arrays[key + "_order"] = order
            arrays[key + "_z"] = out["z"].astype(np.int8)
            arrays[key + "_mu"] = out["mu"]
            arrays[key + "_lambda"] = out["lambda"]
            arrays[key + "_plogLik"] = out["plogLik"]
            if "weights" in out:
                arrays[key + "_weights"] = out["weights"]
            cases.append(case)
    np.savez_compressed(os.path.join(HERE, "golden_chains.npz"), **arrays)
    json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/ref_cpu.cpp (CPU restatement; the "
               "reference itself cannot run in this image)", "cases": cases},
              open(os.path.join(HERE, "golden_meta.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
