"""Development driver: BASELINE config C3 (iterate_deconv on the Visium lattice) for timing / ncu.  argv: iterations, jitter_scale."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bayesspace_b200 as B
import os
if os.environ.get('AB_LIB'):   # A/B of two builds on one box
    import bayesspace_b200._lib as _L
    _L.LIB_PATH = os.path.abspath(os.environ['AB_LIB'])
import bench
from bayesspace_b200 import synth
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 500
js = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
w, csr, _ = bench.build_workload("c2")
pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium", jitter_prior=0.3)
args = lambda nrep: (dc["positions2"], dc["dist"], csr, dc["Y2"], True, nrep, 100, dc["n"], dc["n0"], w.d, 2.0, w.q,
                     dc["init1"], dc["subspots"], False, js, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
B.iterate_deconv(*args(21), seed=3, keep_Y=False)
out = B.iterate_deconv(*args(iters + 1), seed=3, keep_Y=False, mean_from=1)
print("c3 jitter_scale %g: %.2f us / iteration, Ychange %.3f" % (js, 1e3 * out["_info"]["device_ms"] / iters, np.nanmean(out["Ychange"])))
