"""Development driver: runs a chain's iterations as plain launches (bsb200_chain_profile), no CUDA graph."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bayesspace_b200 as bsb
from bayesspace_b200 import synth

name, model, iters = sys.argv[1], sys.argv[2], int(sys.argv[3])
w = synth.make_workload(name)
_, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
ch = bsb.Chain(w.Y, csr, 1000, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta, model=model,
               precision="equal", seed=5, colour=(col, nc))
print("created", flush=True)
for i in range(iters):
    p = ch.profile(1)
    print(i, p, flush=True)
st = ch.state()
print("ok", st["plogLik"][-1])
