"""bayesspace_b200 -- B200-native replacement of BayesSpace's MCMC sampler hot path.

Only the path is here: the C ABI library (csrc/ -> libbsb200.so, include/bsb200.h), the host-side
mirror of the reference's sampler interface (samplers.py), the neighbour construction behind it
(neighbors.py) and the synthetic inputs used by tests and bench (synth.py).
"""
from . import chain_file as _chain_file, enhance_utils, neighbors, samplers, synth  # noqa: F401
from .chain_file import chain_file, read_chain  # noqa: F401
from .enhance_utils import compute_corr, map_subspot2ref  # noqa: F401
from ._lib import COMPAT_Q1, COMPAT_Q3, COMPAT_REFERENCE, BayesSpaceError, device_count, kernel_launches, release_memory  # noqa: F401
from .neighbors import (colour_graph, colour_lattice, convert_neighbors, convert_neighbors2mtx,  # noqa: F401
                        find_neighbors, find_subspot_neighbors)
from .samplers import (Chain, cluster, init_cluster, iterate, iterate_deconv, iterate_t, iterate_t_vvv, iterate_vvv)  # noqa: F401
