"""Development driver: wall clock, set-up and device time of repeated bsb200_cluster() calls at Visium size (C2 lattice);
BSB200_TIMING=1 adds the library's marks."""
import sys, time; sys.path.insert(0,'.')
import numpy as np
import bayesspace_b200 as B, bench
w, csr, colour = bench.build_workload('c2')
for rep in range(4):
    t0=time.perf_counter()
    out = B.cluster(w.Y, w.q, csr, init=w.init, model='normal', precision='equal', mu0=w.mu0, lambda0=w.lambda0, gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=1000, thin=100, seed=7, colour=colour)
    print('wall %.1f ms setup %.1f device %.1f' % (1e3*(time.perf_counter()-t0), out['_info']['setup_ms'], out['_info']['device_ms']), flush=True)
