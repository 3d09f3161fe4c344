"""Run under `python -m torch.distributed.run --nproc-per-node W` with the gloo backend (CPU only).

Checks the host-side logic of a spot-sharded chain across real processes: every rank computes ITS share
(bsb200_shard_plan: own range, ghosts, halo send / receive lists), then
  * the own ranges tile [0, n) in rank order,
  * every send list equals the peer's receive list (same spots, same order),
  * ghosts are exactly the out-of-range neighbours of own spots,
  * the halo protocol itself: labels of own spots are changed colour by colour, the boundary labels travel
    over gloo send/recv following the lists, and afterwards every ghost equals its owner's label.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesspace_b200 as B  # noqa: E402
from bayesspace_b200 import samplers, synth  # noqa: E402


def split(plan, key, ncol, W):
    cnt, ids, out, o = plan[key + "_counts"], plan[key + "_ids"], {}, 0
    for c in range(ncol):
        for p in range(W):
            out[(c, p)] = ids[o:o + cnt[c, p]]
            o += cnt[c, p]
    return out


def check_rendezvous(rank, W):
    """The peer transport's host rendezvous (POSIX shared memory named by the communicator id): several rounds of
    allgather across real processes, with ranks arriving at different times."""
    import time
    os.environ.pop("BSB200_COMM", None)
    ids = [samplers.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0][:8] == b"BSBPEER1"
    mine = (np.arange(1000) * (rank + 3) % 251).astype(np.uint8)
    time.sleep(0.2 * rank)
    got = samplers.debug_rendezvous(ids[0], W, rank, mine, rounds=5)
    for p in range(W):
        expect = ((np.arange(1000) * (p + 3) % 251) + 4).astype(np.uint8)
        assert np.array_equal(got[p], expect), (rank, p)
    # a second communicator in the same processes gets its own segment
    ids = [samplers.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    got = samplers.debug_rendezvous(ids[0], W, rank, np.full(7, rank, np.uint8), rounds=1)
    assert np.array_equal(got, np.repeat(np.arange(W, dtype=np.uint8)[:, None], 7, axis=1))
    if rank == 0:
        assert not [f for f in os.listdir("/dev/shm") if f.startswith("bsb200-")], "rendezvous segment left behind"


def main():
    dist.init_process_group("gloo")
    rank, W = dist.get_rank(), dist.get_world_size()
    check_rendezvous(rank, W)
    for name, groups in (("small_sq", 8), ("small_hex", 64)):
        w = synth.make_workload(name)
        _, csr = B.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
        col, ncol = B.colour_lattice(w.array_row, w.array_col, w.platform)
        plan = samplers.shard_plan(csr, col, ncol, groups, W, rank)
        plans = [None] * W
        dist.all_gather_object(plans, plan)
        # own ranges tile [0, n)
        assert plans[0]["lo"] == 0 and plans[-1]["hi"] == w.n
        assert all(plans[r]["hi"] == plans[r + 1]["lo"] for r in range(W - 1))
        # ghosts = out-of-range neighbours of own spots
        p, idx = csr
        nb = idx[p[plan["lo"]]:p[plan["hi"]]]
        expect = np.unique(nb[(nb < plan["lo"]) | (nb >= plan["hi"])])
        assert np.array_equal(plan["ghosts"], expect)
        # my send lists == the peers' receive lists
        mine = split(plan, "send", ncol, W)
        for peer in range(W):
            theirs = split(plans[peer], "recv", ncol, W)
            for c in range(ncol):
                assert np.array_equal(mine[(c, peer)], theirs[(c, rank)]), (name, rank, peer, c)
                assert np.all(col[mine[(c, peer)]] == c)
        # the protocol: colour by colour, change own labels, ship boundary labels, compare ghosts with owners
        rng = np.random.default_rng(7)          # same stream on every rank: the "true" global label field
        z_true = w.init.copy()
        z_local = {int(j): int(z_true[j]) for j in list(range(plan["lo"], plan["hi"])) + list(plan["ghosts"])}
        recv = split(plan, "recv", ncol, W)
        for c in range(ncol):
            upd = rng.integers(1, w.q + 1, size=w.n)
            sel = col == c
            z_true[sel] = upd[sel]
            for j in range(plan["lo"], plan["hi"]):
                if col[j] == c:
                    z_local[j] = int(upd[j])
            reqs, bufs = [], {}
            for peer in range(W):
                if peer == rank:
                    continue
                if len(mine[(c, peer)]):
                    t = torch.tensor([z_local[int(j)] for j in mine[(c, peer)]], dtype=torch.uint8)
                    reqs.append(dist.isend(t, peer))
                if len(recv[(c, peer)]):
                    bufs[peer] = torch.zeros(len(recv[(c, peer)]), dtype=torch.uint8)
                    reqs.append(dist.irecv(bufs[peer], peer))
            for r in reqs:
                r.wait()
            for peer, t in bufs.items():
                for j, v in zip(recv[(c, peer)], t.tolist()):
                    z_local[int(j)] = v
            assert all(z_local[int(j)] == z_true[j] for j in plan["ghosts"] if col[j] <= c)
        assert all(z_local[j] == z_true[j] for j in z_local)
    dist.barrier()
    if rank == 0:
        print("mrank_plan_check ok, world", W)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
