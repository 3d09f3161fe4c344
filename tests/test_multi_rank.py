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
    assert r.stdout.count("bit-exact") == 10   # 5 cases x 2 transports


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """The driver launches `bench.py --impl reference` like the GPU arm (torchrun for N > 1): rank 0 alone times the
    CPU restatement and prints ONE JSON line with the contract's keys; the other ranks exit 0 without work."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29567", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "2", "--warmup", "1", "--workload", "c2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["steps"] == 2 and j["warmup"] == 1
    assert j["metric"] == "spot-updates/s" and j["unit"] == "spot-updates/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(j["e2e"]["value"] - j["value"]) < 1e-9 * j["value"]


def test_bench_gpu_arm_refuses_to_run_without_a_gpu(bsb):
    """No CPU fallback: on a box without a usable GPU the product arm of bench.py stops with a clear message."""
    if bsb.device_count() > 0:
        pytest.skip("this box has a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--workload", "c2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
