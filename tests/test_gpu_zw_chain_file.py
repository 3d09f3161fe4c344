"""Chain persistence (R/mcmcChain.R:48-73, :181-223): rows streamed to disk by the host layer equal the
matrices .clean_chain / .flatten_matrix_list build from the returned lists."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cluster_chain_file_equals_the_returned_lists(bsb, tmp_path):
    from bayesspace_b200 import synth
    w = synth.make_workload("small_hex")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    args = (w.Y, csr, 60, 10, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    ref = bsb.iterate_t(*args, seed=3)
    d = str(tmp_path / "chain")
    with bsb.chain_file(d):
        out = bsb.iterate_t(*args, seed=3, keep_z=False, keep_weights=False)     # snapshots never materialised
    assert "z" not in out and "weights" not in out
    ch = bsb.read_chain(d)
    rows = 60 // 10 + 1
    assert ch["z"][0].shape == (rows, w.n) and ch["z"][1] == (w.n,) and ch["z"][0].dtype == np.int32
    assert np.array_equal(ch["z"][0], ref["z"]) and np.array_equal(ch["weights"][0], ref["weights"])
    assert np.array_equal(ch["mu"][0], ref["mu"]) and ch["mu"][1] == (w.q, w.d)
    # lambda rows = as.vector(t(lambda_r)) (R/mcmcChain.R:206-211)
    assert np.array_equal(ch["lambda"][0], ref["lambda"].transpose(0, 2, 1).reshape(rows, -1)) and ch["lambda"][1] == (w.d, w.d)
    pl = ch["plogLik"][0][:, 0]
    assert pl.shape == (60,) and np.isnan(pl[0]) and np.array_equal(pl[1:], ref["plogLik"][1:])
    meta = json.load(open(os.path.join(d, "meta.json")))
    assert meta["format"] == "bsb200-chain-1" and set(meta["params"]) == {"z", "weights", "mu", "lambda", "plogLik"}


def test_deconv_chain_file_holds_Y_rows_row_major(bsb, tmp_path):
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    strings, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    pos = np.c_[w.array_col.astype(float), w.array_row * np.sqrt(3.0)]
    dc = synth.deconv_inputs(w.Y, pos, 1.0, np.sqrt(3.0), w.truth, platform="Visium")
    args = (dc["positions2"], dc["dist"], strings, dc["Y2"], True, 30, 5, dc["n"], dc["n0"], w.d, 2.0, w.q, dc["init1"],
            dc["subspots"], False, 5.0, 0, dc["c"], w.mu0, w.lambda0, w.alpha, w.beta, 1)
    ref = bsb.iterate_deconv(*args, seed=9)
    d = str(tmp_path / "enh")
    with bsb.chain_file(d):
        bsb.iterate_deconv(*args, seed=9, keep_Y=False, keep_z=False, keep_weights=False)
    ch = bsb.read_chain(d)
    rows, n = 30 // 5 + 1, dc["n"]
    assert np.array_equal(ch["z"][0], ref["z"]) and np.array_equal(ch["weights"][0], ref["weights"])
    # Y row r = as.vector(t(Y_r)): spot-major rows of d values; dims = (n, d)
    assert ch["Y"][1] == (n, w.d) and np.array_equal(np.asarray(ch["Y"][0]).reshape(rows, n, w.d), ref["Y"])
    assert np.array_equal(ch["Ychange"][0][1:, 0], ref["Ychange"][1:]) and np.array_equal(ch["mu"][0], ref["mu"])
    # a second open while one is active is refused; after close the sink is gone
    from bayesspace_b200 import _lib
    assert _lib.lib().bsb200_chain_file_close() != 0
    out = bsb.iterate_deconv(*args, seed=9)
    assert np.array_equal(out["z"], ref["z"])
