import sys, time; sys.path.insert(0,'.')
import numpy as np
import bayesspace_b200 as B
import os
if os.environ.get('AB_LIB'):   # A/B of two builds on one box
    import bayesspace_b200._lib as _L
    _L.LIB_PATH = os.path.abspath(os.environ['AB_LIB'])
import bench
name = sys.argv[1]; iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w, csr, colour = bench.build_workload(name)
ch = B.Chain(*bench.chain_args(w, csr, 100000, 100), model=w.model, precision=w.precision, seed=1, colour=colour)
ch.run(warm)
ms = ch.run(iters)
p = ch.profile(iters)
print(name, 'n', w.n, 'graph ms/iter %.4f' % (ms/iters), 'sweep ms/launch %.4f' % (p['sweep_ms']/p['sweep_launches']), 'param ms/iter %.4f' % (p['param_ms']/iters), 'launches/iter', p['sweep_launches']/iters)
ch.close()
