"""Development driver: statistics-only pass of the sweep kernel (bsb200_eval_suffstats) at a given size."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bayesspace_b200 as bsb
from bayesspace_b200.samplers import eval_suffstats
n, d, q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(1)
Y = rng.normal(size=(n, d))
z = rng.integers(1, q + 1, size=n).astype(np.int32)
w = rng.gamma(2.0, 0.5, size=n)
mu = rng.normal(size=(q, d))
nk, sk, S = eval_suffstats(Y, z, w, mu, nl=1)
rnk = np.array([w[z == k + 1].sum() for k in range(q)])
rsk = np.array([(Y[z == k + 1] * w[z == k + 1, None]).sum(axis=0) for k in range(q)])
R = Y - mu[z - 1]
rS = (R * w[:, None]).T @ R
print(n, d, q, "nk err", np.abs(nk - rnk).max() / rnk.max(), "sk err", np.abs(sk - rsk).max() / np.abs(rsk).max(),
      "S err", np.abs(S - rS).max() / np.abs(rS).max())
