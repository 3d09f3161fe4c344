#!/usr/bin/env python
"""bench.py -- spot-updates/s of the MCMC sampler hot path on B200 (BASELINE.json metric).

A "step" is one MCMC sweep (one pass of the hot path over every spot of the synthetic lattice):
k_param + one chromatic z/w sweep launch per colour class.  Default workload = BASELINE config C5
(10M-spot square lattice, t model, q=7, d=15): the largest single-GPU configuration, the only one
whose state (1.2 GB of PCs) exceeds the 126 MB L2, i.e. the one on which "% of HBM peak" -- part of the
metric -- is defined.  C2 (Visium, 4,992 spots) and C4 (700k bins) are reported in `other_workloads`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|c2] [--impl reference]

N > 1 is launched by the driver through torch.distributed.run, one rank per GPU: ONE chain is spot-sharded over
the N GPUs in row strips (labels of boundary spots stored into the neighbour GPUs' ghost slots after every colour
phase, per-shard statistics stored into every peer once per iteration, both by the library's own kernels over
NVLink peer memory inside the captured graph; BSB200_COMM=nccl selects the NCCL send/recv + allgather baseline):
total work fixed, "scaling": "strong".
The sharded chain is bit-identical to the 1-GPU chain (tests/mgpu_check.py).  `--replicas` runs one independent
chain per GPU instead (the qTune / multi-chain mode: no data-path collective, weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def load_synth():
    """bayesspace_b200/synth.py (pure numpy) loaded by path: the reference arm must not import the product package,
    which would map libbsb200.so into the process."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bsb200_synth", os.path.join(ROOT, "bayesspace_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


ALGO_NOTE = "8d + 4 (z read) + 4 (z write) + 4*deg (neighbour z) + 8 (w write, t model) bytes per spot-update"


def algorithmic_bytes(d, mean_deg, tdist):
    return 8.0 * d + 4 + 4 + 4.0 * mean_deg + (8 if tdist else 0)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of the job's GPUs while the timed region runs.

    NVML is polled in-process (pynvml, the library nvidia-smi itself reads) from a thread that is started BEFORE
    the warm-up, so that no NVML initialisation or nvidia-smi process start-up lands inside the timed region
    (with 8 ranks each spawning nvidia-smi the driver-wide locks it takes doubled the measured step time).
    Only samples taken between mark() and stop() are reported.  Falls back to one nvidia-smi -lms process."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpus, period_s=0.05):
        self.gpus, self.rows, self.proc, self.period = list(gpus), [], None, period_s
        self.stop_flag, self.first, self.t, self.mode, self.t0 = threading.Event(), threading.Event(), None, None, 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [self._nvml_handle(pynvml, g) for g in self.gpus]
            self.mode = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(map(str, self.gpus)),
                                              "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.mode = "nvidia-smi"
                self.t = threading.Thread(target=self._read_smi, daemon=True)
                self.t.start()
            except Exception:
                self.mode = None
        if self.mode:
            self.first.wait(timeout=10.0)
        return self

    @staticmethod
    def _nvml_handle(nv, cuda_index):
        """NVML enumerates every GPU of the box; CUDA ordinals follow CUDA_VISIBLE_DEVICES when it is set."""
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if cuda_index < len(vis):
            v = vis[cuda_index]
            if v.isdigit():
                return nv.nvmlDeviceGetHandleByIndex(int(v))
            try:
                return nv.nvmlDeviceGetHandleByUUID(v if isinstance(v, bytes) else v.encode())
            except Exception:
                return nv.nvmlDeviceGetHandleByUUID(v)
        return nv.nvmlDeviceGetHandleByIndex(cuda_index)

    def _poll_nvml(self):
        nv = self.nv
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        mx = [nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in self.handles]
        while not self.stop_flag.is_set():
            for h, m in zip(self.handles, mx):
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    self.rows.append((time.perf_counter(), float(sm), float(m), [n for b, n in bits if r & b]))
                except Exception:
                    pass
            self.first.set()
            self.stop_flag.wait(self.period)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                try:
                    self.rows.append((time.perf_counter(), float(f[0]), float(f[1]),
                                      [n for n, v in zip(self.NAMES, f[2:6]) if v == "Active"]))
                except ValueError:
                    pass
                self.first.set()

    def mark(self):
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.mode:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        self.t.join(timeout=2)
        rows = [r for r in self.rows if r[0] >= self.t0] or self.rows[-len(self.gpus):]
        sm, mx, reasons = [r[1] for r in rows], [r[2] for r in rows], set()
        for r in rows:
            reasons.update(r[3])
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.mode, "gpus": self.gpus}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep launch from the committed ncu --set full capture."""
    for rnd in ("r2", "r1"):
        p = os.path.join(ROOT, "profiles", "%s_sweep_%s_ncu.json" % (rnd, name))
        try:
            return float(json.load(open(p))["traffic_bytes_per_launch"]), "profiles/%s_sweep_%s_ncu.json" % (rnd, name)
        except Exception:
            pass
    return None, None


def profile_warm(ch, K):
    """Per-launch CUDA events (bsb200_chain_profile) right behind a few untimed sweeps: the GPU drops its clocks in the
    host-side gap after the timed region, and a 50-sweep profile started cold once read 0.345 ms per launch where the
    timed region itself and every other run had 0.327."""
    ch.run(min(K, 20))
    iters = min(K, 100)
    prof = ch.profile(iters)
    prof["iters"] = iters
    return prof


def build_workload(name):
    import bayesspace_b200 as B
    from bayesspace_b200 import synth
    w = synth.make_workload(name)
    _, csr = B.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    colour = B.colour_lattice(w.array_row, w.array_col, w.platform)
    return w, csr, colour


def chain_args(w, csr, nrep, thin):
    return (w.Y, csr, nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)


def subsample(w, csr, n_s):
    """First n_s spots (whole lattice rows) of the workload with edges leaving the sample dropped."""
    n_s = int(min(w.n, max(w.cols * 4, (n_s // w.cols) * w.cols)))
    ptr = csr[0][:n_s + 1]
    idx = csr[1][:ptr[-1]]
    keep = idx < n_s
    cs = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
    init = w.init[:n_s].copy()
    for k in range(1, w.q + 1):
        if not np.any(init == k):
            init[k - 1] = k
    return n_s, (cs[ptr], idx[keep].copy()), init


def cpu_baseline(w, csr, budget_s=15.0, sweeps=2):
    """Oracle (CPU restatement of the reference, single thread like the reference's cluster samplers)
    on a bounded sample: the first rows of the same lattice."""
    from oracle import oracle as O
    fn = {("normal", "equal"): O.iterate, ("normal", "variable"): O.iterate_vvv, ("t", "equal"): O.iterate_t,
          ("t", "variable"): O.iterate_t_vvv}[(w.model, w.precision)]

    def run(n_s, nsw):
        n_s, g, init = subsample(w, csr, n_s)
        Y = w.Y[:n_s]
        t0 = time.perf_counter()
        fn(Y, g, nsw + 1, 100, n_s, w.d, w.gamma, w.q, init, Y.mean(axis=0), w.lambda0, w.alpha, w.beta, seed=1)
        return n_s, time.perf_counter() - t0

    n0, t = run(min(w.n, 4096), 1)
    rate = n0 / t
    n_s = int(min(w.n, max(4096, rate * budget_s / sweeps)))
    n_s, t = run(n_s, sweeps)
    return {"value": n_s * sweeps / t, "unit": "spot-updates/s", "cores": 1, "kind": "port",
            "sample": "%d sweeps over the first %d spots (%d lattice rows) of the same workload, %.1f s" %
                      (sweeps, n_s, n_s // w.cols, t)}, n_s


def reference_workload(name):
    """The same synthetic workload built WITHOUT the product: lattice from the oracle's .find_neighbors restatement."""
    from oracle import oracle as O
    synth = load_synth()
    w = synth.make_workload(name)
    csr = O.lattice_neighbors(w.array_row, w.array_col, w.platform)
    return w, csr


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, csr = reference_workload(args.workload)
    K, W = max(1, args.steps), max(0, args.warmup)
    # bounded sample sized so that K + W sweeps take about a minute on one host core
    base, n_cal = cpu_baseline(w, csr, budget_s=3.0, sweeps=1)
    rate = base["value"]
    budget = 60.0
    n_s = int(min(w.n, max(4 * w.cols, rate * budget / (K + W))))
    from oracle import oracle as O
    fn = {("normal", "equal"): O.iterate, ("normal", "variable"): O.iterate_vvv, ("t", "equal"): O.iterate_t,
          ("t", "variable"): O.iterate_t_vvv}[(w.model, w.precision)]
    n_s, g, init = subsample(w, csr, n_s)
    Y = w.Y[:n_s]
    a = lambda nsw: (Y, g, nsw + 1, 1000000, n_s, w.d, w.gamma, w.q, init, Y.mean(axis=0), w.lambda0, w.alpha, w.beta)
    if W:
        fn(*a(W), seed=1)
    t0 = time.perf_counter()
    fn(*a(K), seed=2)
    t = time.perf_counter() - t0
    val = n_s * K / t
    sample = "%d sweeps over the first %d spots (%d lattice rows) of %s" % (K, n_s, n_s // w.cols, args.workload)
    assert "bayesspace_b200" not in sys.modules, "the reference arm must not load the product"
    line = {"impl": "reference", "metric": "spot-updates/s", "value": val, "unit": "spot-updates/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * t / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, args.workload, sample=sample),
            "cpu_baseline": {"value": val, "unit": "spot-updates/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "spot-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference's sampler (oracle/ref_cpu.cpp, g++ -O2, 1 thread as in the "
                    "reference's cluster samplers); the R/Rcpp reference itself cannot be built in this image"}
    print(json.dumps(line), flush=True)


def workload_config(w, name, **extra):
    desc = {"c5": "C5: 10M-spot square lattice (3200 x 3125), t model, shared precision, q=7, d=15",
            "c4": "C4: Visium-HD scale 837 x 837 square bins, t model, q=15, d=20",
            "c2": "C2: Visium 78 x 64 hex lattice (4,992 spots), t model, q=7, d=15"}.get(name, name)
    cfg = {"workload": desc, "n_spots": int(w.n), "d": int(w.d), "q": int(w.q), "model": w.model,
           "precision": w.precision, "gamma": w.gamma, "l2_policy": "inputs larger than L2" if w.n * w.d * 8 > 126e6
           else "state fits in L2 (latency-bound size; no flush: a chain re-reads its own state by construction)"}
    cfg.update(extra)
    return cfg


def time_workload(B, w, csr, colour, K, W, device, with_profile=True):
    thin = 100
    nrep = W + 2 * K + 64
    t0 = time.perf_counter()
    ch = B.Chain(*chain_args(w, csr, nrep, thin), model=w.model, precision=w.precision, seed=149, device=device,
                 colour=colour)
    setup_s = time.perf_counter() - t0
    ch.run(W)
    l0 = B.kernel_launches()
    ms = ch.run(K)
    launches = B.kernel_launches() - l0
    prof = profile_warm(ch, K) if with_profile else None
    ch.close()
    return ms, launches, prof, setup_s


def deconv_workload(B, device, peak, with_cpu):
    """BASELINE config C3: iterate_deconv on the Visium lattice (4,992 parents x 6 subspots, t model)."""
    from bayesspace_b200 import synth
    w, csr, _ = build_workload("c2")
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium", jitter_prior=0.3)
    args = lambda nrep: (dc["positions2"], dc["dist"], csr, dc["Y2"], True, nrep, 100, dc["n"], dc["n0"], w.d, 2.0, w.q,
                         dc["init1"], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    B.iterate_deconv(*args(101), seed=3, device=device, keep_Y=False)            # warm-up (graph, caches)
    iters = 2000
    out = B.iterate_deconv(*args(iters + 1), seed=3, device=device, keep_Y=False, mean_from=1)
    ms = out["_info"]["device_ms"]
    bytes_per = 16.0 * w.d + 8.0 * w.d / dc["subspots"] + 8 + 4 * 3 + 16
    achieved = bytes_per * dc["n"] * iters / (ms * 1e-3) / 1e9
    # Robust Adaptive Metropolis variant (jitter.scale = 0): + 2 * 8 * d(d+1)/2 bytes per subspot-update (RAM factors)
    a2 = list(args(501))
    a2[15] = 0.0
    out2 = B.iterate_deconv(*a2, seed=3, device=device, keep_Y=False)
    ms2 = out2["_info"]["device_ms"]
    res = {"roofline": {"bound": "hbm", "kernel": "k_deconv<%d, fixed jitter> (whole iteration: k_param + one launch for all "
                        "colour classes)" % w.d, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_subspot_update": bytes_per,
                        "formula": "16d (Y read + write) + 8d/S (parent row) + 8 (z) + 4*deg + 16 (w) bytes per subspot-update",
                        "note": "3.6 MB of state: resident in L2, the iteration is latency bound (84 CTAs per colour class, "
                                "one warp per SMSP); microseconds per iteration is the meaningful figure"},
           "adaptive": {"us_per_iteration": 1e3 * ms2 / 500, "Ychange_mean": float(np.nanmean(out2["Ychange"])),
                        "algorithmic_bytes_per_subspot_update": bytes_per + 2 * 8 * w.d * (w.d + 1) / 2},
           "config": {"workload": "C3: iterate_deconv, 4,992 Visium spots x 6 subspots = 29,952 subspots, t model, q=7, "
                                  "d=15, jitter.scale=5 (fixed proposal)", "n_subspots": dc["n"]},
           "subspot_updates_per_s": dc["n"] * iters / (ms * 1e-3), "sweeps_per_s": iters / (ms * 1e-3),
           "us_per_iteration": 1e3 * ms / iters, "Ychange_mean": float(np.nanmean(out["Ychange"])),
           "hbm_frac_whole_iteration": bytes_per * dc["n"] * iters / (ms * 1e-3) / 1e9 / peak}
    if with_cpu:
        from oracle import oracle as O
        strings = synth.neighbor_strings(*csr)
        for threads in (1, 8):
            a = list(args(4))
            a[2] = strings
            a[-1] = threads
            t0 = time.perf_counter()
            O.iterate_deconv(*a, seed=3, keep_Y=False)
            t = time.perf_counter() - t0
            res["cpu_baseline_%dthread" % threads] = {"value": dc["n"] * 3 / t, "unit": "subspot-updates/s", "cores": threads,
                                                      "kind": "port", "sample": "3 sweeps of the same workload, %.1f s" % t}
    return res


def bitexact_check(B, dist, rank, world, local):
    """Short sharded-vs-alone comparison on every rank before the timed runs: the spot-sharded chain must reproduce
    the 1-GPU chain bit for bit (labels, weights, mu, lambda, plogLik) -- the north_star's level-1 claim."""
    from bayesspace_b200 import samplers, synth
    ok = True
    for name, shape, fname in (("small_sq", (96, 64), "iterate_t"), ("small_hex", (64, 48), "iterate")):
        w = synth.make_workload(name, n_override=shape)
        _, csr = B.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
        colour = B.colour_lattice(w.array_row, w.array_col, w.platform)
        a = (w.Y, csr, 16, 4, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
        ids = [samplers.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        f = getattr(B, fname)
        alone = f(*a, seed=99, colour=colour, groups=64, device=local)
        shard = f(*a, seed=99, colour=colour, groups=64, device=local, world=world, rank=rank, comm_id=ids[0])
        lo, hi = shard["_info"]["own"]
        ok = ok and np.array_equal(shard["z"][:, lo:hi], alone["z"][:, lo:hi])
        if "weights" in alone:
            ok = ok and np.array_equal(shard["weights"][1:, lo:hi], alone["weights"][1:, lo:hi])
        for key in ("mu", "lambda"):
            ok = ok and np.array_equal(shard[key], alone[key])
        ok = ok and np.array_equal(shard["plogLik"][1:], alone["plogLik"][1:])
    import torch
    t = torch.tensor([1 if ok else 0], dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def qtune_sweep(B, dist, rank, world, local, with_oracle):
    """BASELINE configs[4], second half: qTune's q sweep (R/qTune.R:102-119) -- cluster() for q = 2..15, normal model,
    equal precision, nrep = 1000, burn.in = 100 on the C2 lattice, ONE CHAIN PER GPU with no communication (the
    reference fans the q values out over BiocParallel workers).  Returns wall time, mean plogLik per q, and for one q
    a comparison with the CPU oracle run in the same visiting order."""
    from bayesspace_b200 import synth
    w = synth.make_workload("c2")
    _, csr = B.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    colour = B.colour_lattice(w.array_row, w.array_col, w.platform)
    qs, nrep, burn = list(range(2, 16)), 1000, 100
    mine = [q for i, q in enumerate(qs) if i % world == rank]
    res = {}
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for q in mine:
        init = ((w.truth - 1) % q + 1).astype(np.int32)      # every label present; cluster() is given an init as in qTune
        out = B.cluster(w.Y, q, csr, init=init, model="normal", precision="equal", mu0=w.mu0, lambda0=w.lambda0,
                        gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=nrep, thin=100, seed=1000 + q, device=local,
                        colour=colour)
        res[q] = float(np.mean(out["plogLik"][burn:nrep]))   # mean(plogLik[(burn.in+1):nrep]), R/qTune.R:113
    wall = time.perf_counter() - t0
    if dist:
        import torch
        t = torch.tensor([wall], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
        allres = [None] * world
        dist.all_gather_object(allres, res)
        res = {k: v for r in allres for k, v in r.items()}
    out = {"config": "qTune q sweep: q = 2..15, normal model, equal precision, nrep = 1000, burn.in = 100, C2 lattice "
                     "(4,992 spots, d = 15), one chain per GPU, no communication",
           "n_gpus": world, "wall_s": wall, "chains": len(qs), "sweeps_per_s_all_chains": len(qs) * (nrep - 1) / wall,
           "mean_plogLik": {str(q): res[q] for q in sorted(res)}}
    if with_oracle and rank == 0:
        from oracle import oracle as O
        q, n2 = 4, 60
        init = ((w.truth - 1) % q + 1).astype(np.int32)
        order = np.lexsort((np.arange(w.n), colour[0])).astype(np.int32)
        a = (w.Y, csr, n2, 10, w.n, w.d, w.gamma, q, init, w.mu0, w.lambda0, w.alpha, w.beta)
        got = B.iterate(*a, seed=1004, device=local, colour=colour)
        ref = O.iterate(*a, seed=1004, order=order)
        out["oracle_check"] = {"q": q, "nrep": n2, "labels_identical": bool(np.array_equal(got["z"], ref["z"])),
                               "plogLik_max_rel_err": float(np.max(np.abs(got["plogLik"][1:] - ref["plogLik"][1:]) /
                                                                   np.abs(ref["plogLik"][1:])))}
    return out


def sharded_extra(B, dist, rank, world, local, name, K, W):
    """One more configuration spot-sharded over the job's GPUs (C4 beside the default C5): ms per sweep, max over ranks."""
    from bayesspace_b200 import samplers
    w, csr, colour = build_workload(name)
    ids = [samplers.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ch = B.Chain(*chain_args(w, csr, W + 2 * K + 64, 100), model=w.model, precision=w.precision, seed=149, device=local,
                 colour=colour, world=world, rank=rank, comm_id=ids[0])
    ch.run(W)
    import torch
    dist.barrier()
    torch.cuda.synchronize()
    ms = ch.run(K)
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    prof = profile_warm(ch, K)
    ch.close()
    mean_deg = len(csr[1]) / w.n
    per_launch = prof["sweep_ms"] / max(1, prof["sweep_launches"])
    peak, _ = measured_peak()
    return {"config": workload_config(w, name), "n_gpus": world, "ms_per_step": ms / K,
            "spot_updates_per_s": w.n * K / (ms * 1e-3), "sweep_launch_ms": per_launch,
            "param_kernel_ms_per_iter": prof["param_ms"] / max(1, prof["iters"]),
            "halo_exchange_ms_per_iter": prof["halo_ms"] / max(1, prof["iters"]),
            "hbm_frac_sweep_kernel": (algorithmic_bytes(w.d, mean_deg, w.model == "t") * (w.n / world) / colour[1]) /
                                     (per_launch * 1e-3) / 1e9 / peak}


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        if os.environ.get("BSB200_COMM") == "nccl":
            # the communicator lines stay checkable, but in a file: NCCL's banner on stdout would precede the JSON line
            os.environ.setdefault("NCCL_DEBUG", "INFO")
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(ROOT, "gpurun_out", "nccl_%h_%p.log"))
        else:
            os.environ["NCCL_DEBUG"] = "WARN"   # torch's own NCCL must not print a banner before the JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")   # plumbing only (ids, barriers, max of the timings); data path = library's NCCL
    import bayesspace_b200 as B
    from bayesspace_b200 import samplers
    if B.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; bayesspace_b200 has no CPU fallback")
    K, W = max(1, args.steps), max(3, args.warmup)
    w, csr, colour = build_workload(args.workload)
    mean_deg = len(csr[1]) / w.n
    Bspot = algorithmic_bytes(w.d, mean_deg, w.model == "t")
    peak, peak_src = measured_peak()

    thin = 100
    nrep = W + 2 * K + 64
    sharded = world > 1 and not args.replicas

    def comm_id():
        if not sharded:
            return None
        ids = [samplers.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    bitexact = bitexact_check(B, dist, rank, world, local) if sharded else None
    shard_kw = dict(world=world, rank=rank, comm_id=comm_id()) if sharded else {}
    ch = B.Chain(*chain_args(w, csr, nrep, thin), model=w.model, precision=w.precision,
                 seed=149 + (0 if sharded else rank), device=local, colour=colour, **shard_kw)
    # rank 0 watches every GPU of the job; the sampler is up (first sample in) before the warm-up starts
    clocks = ClockSampler(range(world) if rank == 0 else []).start() if rank == 0 else None
    ch.run(W)
    if dist:
        import torch
        dist.barrier()
        torch.cuda.synchronize()
    if clocks:
        clocks.mark()
    l0 = B.kernel_launches()
    ms = ch.run(K)   # CUDA events on the library's stream around the K captured iterations
    launches = B.kernel_launches() - l0
    if dist:
        import torch
        torch.cuda.synchronize()
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    prof = profile_warm(ch, K)
    ck = clocks.stop() if clocks else None
    ch.close()
    # ---- end to end through the C ABI with host buffers (bsb200_cluster: H2D, chain, snapshots D2H) ------
    e2e_iters = args.e2e_iters
    if dist:
        dist.barrier()
    e2e_kw = dict(world=world, rank=rank, comm_id=comm_id()) if sharded else {}
    t0 = time.perf_counter()
    out = B.cluster(w.Y, w.q, csr, init=w.init, model=w.model, precision=w.precision, mu0=w.mu0, lambda0=w.lambda0,
                    gamma=w.gamma, alpha=w.alpha, beta=w.beta, nrep=e2e_iters + 1, thin=100, seed=7, device=local,
                    colour=colour, **e2e_kw)
    e2e_s = time.perf_counter() - t0
    if dist:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    extra = {}
    if sharded and args.others:
        extra["c4"] = sharded_extra(B, dist, rank, world, local, "c4", 300, 10)
    if args.others:
        extra["qtune"] = qtune_sweep(B, dist if world > 1 else None, rank, world, local, not args.no_cpu)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    total_units = w.n if sharded or world == 1 else world * w.n
    n_dev = w.n // world if sharded else w.n          # spot-updates one GPU does per sweep
    value = total_units * K / (ms * 1e-3)
    sweep_ms_per_launch = prof["sweep_ms"] / max(1, prof["sweep_launches"])
    bytes_per_launch = Bspot * n_dev / colour[1]
    achieved = bytes_per_launch / (sweep_ms_per_launch * 1e-3) / 1e9
    line = {"metric": "spot-updates/s", "value": value, "unit": "spot-updates/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak" if (world > 1 and not sharded) else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, args.workload, parallelism=(
                "one chain spot-sharded over %d GPUs in row strips; halo labels and shard statistics travel as %s"
                % (world, "NCCL send/recv + allgather" if os.environ.get("BSB200_COMM") == "nccl" else
                   "stores into peer GPU memory from the library's own kernels (CUDA IPC over NVLink, epoch flags)")
                if sharded else ("1 chain per GPU (replicas, no collective)" if world > 1 else "single chain, 1 GPU")),
                sweeps_per_s=(1 if sharded or world == 1 else world) * K / (ms * 1e-3), colours=colour[1]),
            "roofline": {"bound": "hbm", "kernel": "k_sweep_fast<%d, t, NH=%d>" % (w.d, (w.q + 7) // 8), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(args.workload)[0],
                         "traffic_source": "%s (ncu --set full, bytes per launch)" % ncu_traffic(args.workload)[1],
                         "algorithmic_bytes_per_spot_update": Bspot, "bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": sweep_ms_per_launch, "formula": ALGO_NOTE,
                         "param_kernel_ms_per_iter": prof["param_ms"] / max(1, prof["iters"]),
                         "halo_exchange_ms_per_iter": prof["halo_ms"] / max(1, prof["iters"])},
            "gpu_launches": int(launches), "clocks": ck}
    if bitexact is not None:
        line["bitexact"] = bitexact   # sharded chain == 1-GPU chain, bit for bit, checked on every rank before timing
    nsnap = (e2e_iters + 1) // 100
    h2d = w.n * w.d * 8 + w.n * 4 + len(csr[1]) * 4 + (w.n + 1) * 8
    d2h = nsnap * w.n * (4 + (8 if w.model == "t" else 0)) + (e2e_iters + 1) * 8
    line["e2e"] = {"value": total_units * e2e_iters / e2e_s, "unit": "spot-updates/s",
                   "h2d_bytes_per_step": h2d / e2e_iters, "d2h_bytes_per_step": d2h / e2e_iters,
                   "call": "bsb200_cluster(nrep=%d, thin=100) with host buffers: wall clock of the whole call "
                           "(validation, staged H2D of pageable inputs, spot order and neighbour lists on the device, %d sweeps, %d snapshots D2H)" %
                           (e2e_iters + 1, e2e_iters, nsnap),
                   "wall_s": e2e_s, "setup_ms": out["_info"]["setup_ms"], "device_ms": out["_info"]["device_ms"]}
    # ---- CPU baseline beside it (rank 0, bounded sample) ----------------------------------------------------
    if not args.no_cpu and world == 1:
        cb, _ = cpu_baseline(w, csr)
        line["cpu_baseline"] = cb
    # ---- the other sizes the metric is quoted on -------------------------------------------------------------
    if args.others and world == 1:
        others = {}
        for name in [x for x in ("c2", "c4") if x != args.workload]:
            w2, csr2, col2 = build_workload(name)
            k2 = 2000 if name == "c2" else 200
            ms2, _, prof2, _ = time_workload(B, w2, csr2, col2, k2, 20, local)
            b2 = algorithmic_bytes(w2.d, len(csr2[1]) / w2.n, True)
            per_launch = prof2["sweep_ms"] / max(1, prof2["sweep_launches"])
            others[name] = {"config": workload_config(w2, name), "spot_updates_per_s": w2.n * k2 / (ms2 * 1e-3),
                            "sweeps_per_s": k2 / (ms2 * 1e-3), "us_per_iteration": 1e3 * ms2 / k2,
                            "hbm_frac_sweep_kernel": (b2 * w2.n / col2[1]) / (per_launch * 1e-3) / 1e9 / peak}
            if not args.no_cpu:
                others[name]["cpu_baseline"], _ = cpu_baseline(w2, csr2, budget_s=8.0)
        others["c3"] = deconv_workload(B, local, peak, not args.no_cpu)
        others.update(extra)
        line["other_workloads"] = others
    elif extra:
        line["other_workloads"] = extra
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c4", "c2"])
    ap.add_argument("--e2e-iters", type=int, default=1000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: one independent chain per GPU instead of sharding")
    ap.add_argument("--no-others", dest="others", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
