"""Small driver used while developing the sweep kernels: one short chain of a named workload."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bayesspace_b200 as bsb
from bayesspace_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "small_sq"
fname = sys.argv[2] if len(sys.argv) > 2 else "iterate"
nrep = int(sys.argv[3]) if len(sys.argv) > 3 else 5
w = synth.make_workload(name)
_, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
out = getattr(bsb, fname)(w.Y, csr, nrep, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta, seed=5,
                          colour=(col, nc))
print(name, fname, "ok", out["plogLik"][-1], out["z"][-1][:10])
