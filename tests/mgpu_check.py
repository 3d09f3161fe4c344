"""Run under `python -m torch.distributed.run --nproc-per-node W` on a box with W GPUs.

Every rank runs the same chain twice on its own GPU -- alone (world = 1) and as rank r of a spot-sharded chain
over W GPUs (labels of boundary spots exchanged after every colour phase, per-shard statistics allgathered) --
with the same number of virtual shards, once per transport (peer-memory stores, NCCL).  The two must agree BIT FOR BIT: labels, weights (own spots), mu,
lambda and plogLik (north_star: "1/2/4/8-GPU chains are bit-exact").
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesspace_b200 as B  # noqa: E402
from bayesspace_b200 import samplers, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    rank, W = dist.get_rank(), dist.get_world_size()
    # last case: a proper FOUR-colouring of the square lattice, (row % 2) * 2 + col % 2 -- the boundary row of a strip
    # then holds only two of the four colours, so every colour phase has label traffic in ONE direction only between
    # two neighbouring ranks (the one-sided legs the symmetric handshake of the peer transport exists for)
    cases = [("small_sq", (64, 64), "iterate", 16, None), ("small_hex", (60, 40), "iterate_t", 64, None),
             ("small_sq", (48, 64), "iterate_t_vvv", 8, None), ("small_hex", (40, 30), "iterate_vvv", 8, None),
             ("small_sq", (64, 48), "iterate_t", 16, "four")]
    transports = os.environ.get("BSB200_CHECK_TRANSPORTS", "peer,nccl").split(",")
    for name, shape, fname, groups, colouring in [c for c in cases for _ in transports]:
        transport = transports[0]
        transports = transports[1:] + transports[:1]
        os.environ["BSB200_COMM"] = transport   # read when the id is made: peer-memory stores (default) or NCCL
        w = synth.make_workload(name, n_override=shape)
        _, csr = B.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
        colour = B.colour_lattice(w.array_row, w.array_col, w.platform)
        if colouring == "four":
            colour = (((w.array_row % 2) * 2 + w.array_col % 2).astype(np.int32), 4)
        args = (w.Y, csr, 24, 4, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
        f = getattr(B, fname)
        ids = [samplers.comm_unique_id() if rank == 0 else None]   # one communicator id per sharded chain
        dist.broadcast_object_list(ids, src=0)
        alone = f(*args, seed=99, colour=colour, groups=groups, device=local)
        shard = f(*args, seed=99, colour=colour, groups=groups, device=local, world=W, rank=rank, comm_id=ids[0])
        lo, hi = shard["_info"]["own"]
        assert hi > lo and (hi - lo) < w.n
        assert np.array_equal(shard["z"][:, lo:hi], alone["z"][:, lo:hi]), (fname, "labels")
        if "weights" in alone:
            assert np.array_equal(shard["weights"][1:, lo:hi], alone["weights"][1:, lo:hi]), (fname, "weights")
        assert np.array_equal(shard["mu"], alone["mu"]), (fname, "mu")
        assert np.array_equal(shard["lambda"], alone["lambda"]), (fname, "lambda")
        assert np.array_equal(shard["plogLik"][1:], alone["plogLik"][1:]), (fname, "plogLik")
        dist.barrier()
        if rank == 0:
            print("bit-exact", transport, fname, name, shape, "world", W, "plogLik[-1] = %.6f" % alone["plogLik"][-1], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
