"""Multi-rank tests: host logic on CPU (gloo, world_size 2 and 4) and, with >= 2 GPUs, bit-exactness of a
spot-sharded chain against the single-GPU chain."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 4])
def test_shard_plan_and_halo_protocol_gloo(bsb, world):
    r = _torchrun("mrank_plan_check.py", world, 29517 + world)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mrank_plan_check ok" in r.stdout


def test_shard_plan_rejects_bad_worlds(bsb):
    from bayesspace_b200 import samplers, synth
    w = synth.make_workload("small_sq")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    col, nc = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    with pytest.raises(bsb.BayesSpaceError, match="must divide"):
        samplers.shard_plan(csr, col, nc, 8, 3, 0)
    with pytest.raises(bsb.BayesSpaceError):
        samplers.shard_plan(csr, col, nc, 8, 2, 2)
    one = samplers.shard_plan(csr, col, nc, 8, 1, 0)
    assert (one["lo"], one["hi"], len(one["ghosts"])) == (0, w.n, 0)


@pytest.mark.gpu
def test_sharded_chain_is_bit_exact(bsb):
    ngpu = bsb.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_multi_rank.py -m gpu)")
    world = 4 if ngpu >= 4 else 2
    r = _torchrun("mgpu_check.py", world, 29541, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("bit-exact") == 4
