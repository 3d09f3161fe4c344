"""Device-side post-sampling reductions: the per-spot label mode the R callers compute with
apply(zs, 2, Mode), against the restatement of R's Mode (oracle.mode_rows) applied to the very label rows the same
call returned -- and to the oracle's rows, which are identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(bsb, name):
    from bayesspace_b200 import synth
    w = synth.make_workload(name)
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    colour = bsb.colour_lattice(w.array_row, w.array_col, w.platform)
    return w, csr, colour


@pytest.mark.parametrize("name,fname,burn_rows,nrep,thin", [
    ("small_hex", "iterate_t", 1, 7, 1), ("small_sq", "iterate", 3, 64, 4), ("tiny_hex", "iterate_t_vvv", 1, 7, 1),
    ("small_sq", "iterate_vvv", 2, 7, 1)])
def test_label_mode_matches_R_Mode(bsb, oracle, name, fname, burn_rows, nrep, thin):
    w, csr, colour = _setup(bsb, name)   # the short unthinned chains still move: spots with tied label counts exist
    args = (w.Y, csr, nrep, thin, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    got = getattr(bsb, fname)(*args, seed=21, colour=colour, mode_from=burn_rows)
    expect = oracle.mode_rows(got["z"][burn_rows:])
    assert np.array_equal(got["z_mode"], expect)
    ties = 0
    for j in range(w.n):
        c = np.bincount(got["z"][burn_rows:, j], minlength=w.q + 1)
        ties += int(np.sum(c == c.max()) > 1)
    if thin == 1:
        assert ties > 0, "the case should exercise Mode()'s tie rule"
    # without the snapshot matrices: same mode, same parameters, nothing else returned
    lean = getattr(bsb, fname)(*args, seed=21, colour=colour, mode_from=burn_rows, keep_z=False, keep_weights=False)
    assert "z" not in lean and "weights" not in lean
    assert np.array_equal(lean["z_mode"], expect) and np.array_equal(lean["mu"], got["mu"])
    assert np.array_equal(lean["plogLik"][1:], got["plogLik"][1:])


def test_label_mode_no_row_kept(bsb):
    w, csr, colour = _setup(bsb, "tiny_hex")
    args = (w.Y, csr, 20, 5, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    got = bsb.iterate_t(*args, seed=1, colour=colour, mode_from=99)
    assert np.all(got["z_mode"] == 0)
    one = bsb.iterate_t(*args, seed=1, colour=colour, mode_from=4)   # a single kept row is returned as is
    assert np.array_equal(one["z_mode"], one["z"][4])


def test_deconv_label_mode_and_mean(bsb, oracle):
    from bayesspace_b200 import synth
    w = synth.make_workload("small_hex")
    strings, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium")
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], True, 60, 5, dc["n"], dc["n0"], w.d, 2.0, w.q,
            dc["init1"], dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    got = bsb.iterate_deconv(*args, seed=8, mean_from=3, mode_from=3)
    assert np.array_equal(got["z_mode"], oracle.mode_rows(got["z"][3:]))
    assert np.allclose(got["Y_mean"], got["Y"][3:].mean(axis=0), rtol=1e-12, atol=1e-12)


def test_interrupt_poll_stops_the_call(bsb):
    """Rcpp::checkUserInterrupt() analogue (src/cluster.cpp:243-244): the host registers a poll function, the library
    calls it after every batch of at most `thin` iterations, and a non-zero return ends the call with
    BSB200_ERR_CANCELLED ("Stopping..."), the snapshot rows reached so far being complete."""
    import ctypes as C
    from bayesspace_b200 import _lib, synth
    w = synth.make_workload("small_hex")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    calls = []
    POLL = C.CFUNCTYPE(C.c_int, C.c_void_p)

    def poll(_user):
        calls.append(1)
        return 1 if len(calls) >= 3 else 0

    cb = POLL(poll)
    L = _lib.lib()
    L.bsb200_set_interrupt_poll.argtypes = [POLL, C.c_void_p]
    L.bsb200_set_interrupt_poll.restype = None
    L.bsb200_set_interrupt_poll(cb, None)
    try:
        with pytest.raises(bsb.BayesSpaceError, match="Stopping") as ei:
            bsb.iterate_t(w.Y, csr, 1000, 10, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta, seed=4)
        assert ei.value.code == -8 and len(calls) == 3
    finally:
        L.bsb200_set_interrupt_poll(POLL(0), None)
    out = bsb.iterate_t(w.Y, csr, 40, 10, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta, seed=4)
    assert out["_info"]["rows_done"] == 5          # the poll is gone: the next call runs to the end
