"""Development driver: where the wall clock of the bench's end-to-end call goes on the Python side
(argument marshalling, output allocation, the C call itself).  BSB200_TIMING=1 adds the library's set-up marks."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bayesspace_b200 as B
from bayesspace_b200 import samplers
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w, csr, colour = bench.build_workload(name)
print("Y flags: C", w.Y.flags.c_contiguous, "F", w.Y.flags.f_contiguous, w.Y.dtype, flush=True)
orig_args, orig_zeros = samplers._cluster_args, np.zeros
tm = {}
def timed_args(*a, **k):
    t = time.perf_counter(); r = orig_args(*a, **k); tm["args"] = time.perf_counter() - t; tm["t_args_end"] = time.perf_counter(); return r
samplers._cluster_args = timed_args
real_lib = samplers.lib
class L:
    def __init__(s, l): s.l = l
    def __getattr__(s, k):
        f = getattr(s.l, k)
        if k != "bsb200_cluster": return f
        def g(*a):
            tm["alloc"] = time.perf_counter() - tm["t_args_end"]
            t = time.perf_counter(); r = f(*a); tm["call"] = time.perf_counter() - t; tm["t_call_end"] = time.perf_counter(); return r
        return g
samplers.lib = lambda: L(real_lib())
for rep in range(2):
    t0 = time.perf_counter()
    out = B.cluster(w.Y, w.q, csr, init=w.init, model=w.model, precision=w.precision, mu0=w.mu0, lambda0=w.lambda0,
                    gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=iters + 1, thin=100, seed=7, colour=colour)
    t1 = time.perf_counter()
    print("rep %d: wall %.1f ms = args %.1f + output alloc %.1f + C call %.1f (setup %.1f, device %.1f, other %.1f) + after %.1f" %
          (rep, 1e3 * (t1 - t0), 1e3 * tm["args"], 1e3 * tm["alloc"], 1e3 * tm["call"], out["_info"]["setup_ms"],
           out["_info"]["device_ms"], 1e3 * tm["call"] - out["_info"]["setup_ms"] - out["_info"]["device_ms"],
           1e3 * (t1 - tm["t_call_end"])), flush=True)
