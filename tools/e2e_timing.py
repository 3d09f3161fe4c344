"""Development driver: wall-clock split of one bsb200_cluster() call with host buffers (BSB200_TIMING=1 prints the set-up phases)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bayesspace_b200 as B
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w, csr, colour = bench.build_workload(name)
for rep in range(2):
    t0 = time.perf_counter()
    out = B.cluster(w.Y, w.q, csr, init=w.init, model=w.model, precision=w.precision, mu0=w.mu0, lambda0=w.lambda0,
                    gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=iters + 1, thin=100, seed=7, colour=colour)
    t = time.perf_counter() - t0
    print("rep %d: wall %.3f s, setup %.1f ms, device %.1f ms, rest %.1f ms" %
          (rep, t, out["_info"]["setup_ms"], out["_info"]["device_ms"], 1e3 * t - out["_info"]["setup_ms"] - out["_info"]["device_ms"]), flush=True)
