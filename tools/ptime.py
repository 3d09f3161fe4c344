import sys, ctypes as C; sys.path.insert(0,'.')
import numpy as np
import bayesspace_b200 as B, bench
from bayesspace_b200 import _lib
import os
if os.environ.get('AB_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['AB_LIB'])
for name in ['c2','c4','c5']:
    w, csr, colour = bench.build_workload(name)
    ch = B.Chain(*bench.chain_args(w, csr, 100000, 100), model=w.model, precision=w.precision, seed=1, colour=colour)
    ch.run(20)
    t = (C.c_longlong*16)()
    _lib.lib().bsb200_debug_param_timing(t)
    t=list(t); print(name, 'reduce',t[1]-t[0],'plog',t[2]-t[1],'phase1',t[3]-t[2],'phase2',t[4]-t[3],'tail',t[5]-t[4],'total',t[5]-t[0], 'p2 steps', [t[6]-t[3]]+[t[i+1]-t[i] for i in range(6,13)])
    ch.close()
