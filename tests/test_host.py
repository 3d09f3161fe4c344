"""CPU tests of the product's host layer: the C-ABI library loads and exports every symbol the header
declares, and the host-only entry points (neighbour construction, colouring, argument checks) agree with
the oracle.  No compute kernel is called here (there is no GPU in this container)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(bsb):
    from bayesspace_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "bsb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(bsb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = _lib.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == declared
    assert L.bsb200_version() == 100


def test_no_fallback_without_a_gpu(bsb):
    """The product path must fail loudly when it cannot run on a B200."""
    if bsb.device_count() > 0:
        pytest.skip("a GPU is present")
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    with pytest.raises(bsb.BayesSpaceError) as ei:
        bsb.iterate_t(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0, w.alpha, w.beta)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)
    with pytest.raises(bsb.BayesSpaceError):
        bsb.samplers.eval_loglik(w.Y, w.mu_true, np.eye(w.d))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesspace_b200")
    for dirpath, _, files in os.walk(pkg):
        if "obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\".*oracle|liboracle", txt, flags=re.M), f


def test_find_neighbors_reference_known_answer(bsb):
    """tests/testthat/test-find_neighbors.R:1-12, through the C ABI."""
    positions = np.c_[np.arange(1, 6), np.arange(5, 0, -1)]
    df_j = bsb.find_neighbors(positions, 1, "manhattan")
    assert all(len(v) == 0 for v in df_j)
    df_j = bsb.find_neighbors(positions, 4, "manhattan")
    assert list(df_j[0]) == [1, 2] and list(df_j[2]) == [0, 1, 3, 4]
    with pytest.raises(ValueError):
        bsb.find_neighbors(positions, 1, "chebyshev")


def test_radius_neighbors_match_oracle(bsb, oracle):
    rng = np.random.default_rng(2)
    pos = rng.uniform(0, 30, size=(700, 2))
    pos[10] = pos[11]   # coincident points: dist 0 is not a neighbour
    for method, radius in [("manhattan", 2.5), ("euclidean", 1.7), ("euclidean", 0.0)]:
        a, b = bsb.find_neighbors(pos, radius, method), oracle.find_neighbors(pos, radius, method)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert 11 not in bsb.find_neighbors(pos, 2.5, "manhattan")[10]


@pytest.mark.parametrize("platform", ["Visium", "ST", "VisiumHD"])
def test_lattice_neighbors_match_oracle(bsb, oracle, platform):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(3)
    r, c = synth.hex_lattice(30, 24) if platform == "Visium" else synth.square_lattice(20, 25)
    keep = rng.random(len(r)) < 0.8            # ragged tissue mask, spots in shuffled order
    perm = rng.permutation(int(keep.sum()))
    r, c = r[keep][perm], c[keep][perm]
    strings, csr = bsb.neighbors._find_neighbors(r, c, platform)
    optr, oidx = oracle.lattice_neighbors(r, c, platform)
    assert np.array_equal(csr[0], optr) and np.array_equal(csr[1], oidx[:len(csr[1])])
    lists = bsb.neighbors.csr_to_lists(csr)
    assert all(np.all(np.diff(v) > 0) for v in lists if len(v) > 1)       # ascending (R/spatialCluster.R:263-267)
    j = int(np.argmax(np.diff(csr[0]) == 0)) if np.any(np.diff(csr[0]) == 0) else None
    if j is not None:
        assert strings[j] is None                                           # isolated spot -> NA
    k = int(np.argmax(np.diff(csr[0]) > 0))
    assert strings[k] == ",".join(str(v + 1) for v in lists[k])             # 1-based strings (:276-286)
    back = bsb.convert_neighbors(strings)
    assert np.array_equal(back[0], csr[0]) and np.array_equal(back[1], csr[1])
    with pytest.raises(ValueError, match="Unsupported platform"):
        bsb.neighbors._find_neighbors(r, c, "Slide-seq")


def test_lattice_neighbors_duplicate_and_sparse_coordinates(bsb, oracle):
    r = np.array([0, 0, 0, 1, 1000000, 0], dtype=np.int32)
    c = np.array([0, 1, 1, 0, 5, 2], dtype=np.int32)      # spot 1 and 2 share a coordinate
    _, csr = bsb.neighbors._find_neighbors(r, c, "ST")
    optr, oidx = oracle.lattice_neighbors(r, c, "ST")
    assert np.array_equal(csr[0], optr) and np.array_equal(csr[1], oidx[:len(csr[1])])
    assert list(bsb.neighbors.csr_to_lists(csr)[0]) == [1, 2, 3]


def test_visium_graph_facts(bsb):
    """SURVEY.md App. E."""
    from bayesspace_b200 import synth
    r, c = synth.hex_lattice(78, 64)
    _, csr = bsb.neighbors._find_neighbors(r, c, "Visium")
    deg = np.diff(csr[0])
    assert len(csr[1]) == 2 * 14693
    assert {int(k): int(v) for k, v in zip(*np.unique(deg, return_counts=True))} == {2: 2, 3: 78, 4: 124, 5: 76, 6: 4712}
    col, nc = bsb.colour_lattice(r, c, "Visium")
    assert nc == 3 and sorted(np.bincount(col)) == [1638, 1677, 1677]
    p, idx = csr
    assert all(col[idx[e]] != col[j] for j in range(len(r)) for e in range(p[j], p[j + 1]))


def test_subspot_neighbors_match_oracle(bsb, oracle):
    from bayesspace_b200 import synth
    for platform, k in [("Visium", None), ("ST", 3), ("VisiumHD", 2)]:
        r, c = synth.hex_lattice(12, 10) if platform == "Visium" else synth.square_lattice(7, 9)
        _, csr = bsb.neighbors._find_neighbors(r, c, platform)
        pos = np.c_[c.astype(float), r * (np.sqrt(3.0) if platform == "Visium" else 1.0)]
        dc = synth.deconv_inputs(np.random.default_rng(0).normal(size=(len(r), 3)), pos, 1.0, np.sqrt(3.0) if platform == "Visium" else 1.0,
                                 np.ones(len(r), dtype=np.int32), platform=platform, nsubspots_per_edge=k or 3)
        a = bsb.find_subspot_neighbors(csr, dc["positions2"], dc["dist"])
        b = oracle.find_subspot_neighbors(csr, dc["positions2"], dc["dist"])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1][:len(a[1])])
        m = bsb.convert_neighbors2mtx(a)
        assert m.shape == (dc["n"], 4) and m.max() <= dc["n"] and ((m > 0).sum(axis=1) == np.diff(a[0])).all()
        col, nc = bsb.colour_graph(a)
        assert nc == 2                                                      # subspot graphs are bipartite
    with pytest.raises(bsb.BayesSpaceError, match="Invalid arguments"):
        bsb.find_subspot_neighbors(csr, np.zeros((len(r) * 4 + 1, 2)), 1.0)
    with pytest.raises(bsb.BayesSpaceError, match="finding neighbors of subspots"):
        bsb.find_subspot_neighbors(csr, dc["positions2"] * 100.0, dc["dist"])


def test_parse_neighbor_strings(bsb):
    p, idx = bsb.convert_neighbors(["2,3", None, "1", "1,,2", ""])
    assert list(p) == [0, 2, 2, 3, 5, 5] and list(idx) == [1, 2, 0, 0, 1]
    with pytest.raises(bsb.BayesSpaceError) as ei:
        bsb.convert_neighbors(["1,x"])
    assert ei.value.code == -3 and "Unconvertable" in str(ei.value)
    with pytest.raises(bsb.BayesSpaceError, match="out of range"):
        bsb.convert_neighbors(["99999999999"])


def test_matrix4_rejects_degree_above_four(bsb):
    p = np.array([0, 5, 5, 5, 5, 5, 5], dtype=np.int64)
    idx = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    with pytest.raises(bsb.BayesSpaceError) as ei:
        bsb.convert_neighbors2mtx((p, idx))
    assert ei.value.code == -7


def test_colourings_are_proper(bsb):
    from bayesspace_b200 import synth
    rng = np.random.default_rng(7)

    def proper(csr, col):
        p, idx = csr
        return all(col[idx[e]] != col[j] for j in range(len(p) - 1) for e in range(p[j], p[j + 1]) if idx[e] != j)

    r, c = synth.square_lattice(13, 17)
    _, csr = bsb.neighbors._find_neighbors(r, c, "ST")
    col, nc = bsb.colour_lattice(r, c, "ST")
    g, ng = bsb.colour_graph(csr)
    assert nc == ng == 2 and proper(csr, col) and np.array_equal(col == col[0], g == g[0])
    # hex offsets on a filled rectangle (quirk Q13, qTune's default platform): greedy must still be proper
    _, csr13 = bsb.neighbors._find_neighbors(r, c, "Visium")
    g, ng = bsb.colour_graph(csr13)
    assert proper(csr13, g) and ng <= 4
    with pytest.raises(bsb.BayesSpaceError):
        bsb.colour_lattice(r, c, "Visium")       # col != row (mod 2): the closed form does not apply
    # random directed graph with self loops: colour on the symmetrised graph
    n = 300
    lists = [np.unique(np.r_[rng.integers(0, n, size=rng.integers(0, 5)), [j] if j % 7 == 0 else []]).astype(np.int32)
             for j in range(n)]
    csr_r = bsb.neighbors.lists_to_csr(lists)
    g, ng = bsb.colour_graph(csr_r)
    assert proper(csr_r, g) and g.min() == 0 and g.max() == ng - 1


def test_argument_checks_precede_the_device(bsb):
    """Errors the R wrapper / Rcpp glue raise are raised before any CUDA call."""
    from bayesspace_b200 import synth
    w = synth.make_workload("tiny_hex")
    _, csr = bsb.neighbors._find_neighbors(w.array_row, w.array_col, w.platform)
    with pytest.raises(bsb.BayesSpaceError, match="mu0"):
        bsb.iterate(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0[:-1], w.lambda0, w.alpha, w.beta)
    with pytest.raises(bsb.BayesSpaceError, match="lambda0"):
        bsb.iterate(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, w.init, w.mu0, w.lambda0[:-1], w.alpha, w.beta)
    bad = w.init.copy()
    bad[3] = 0
    with pytest.raises(bsb.BayesSpaceError, match="outside 1..q"):
        bsb.iterate(w.Y, csr, 10, 1, w.n, w.d, w.gamma, w.q, bad, w.mu0, w.lambda0, w.alpha, w.beta)
    Ybig = np.zeros((w.n, 40))
    with pytest.raises(bsb.BayesSpaceError) as ei:
        bsb.iterate(Ybig, csr, 10, 1, w.n, 40, w.gamma, w.q, w.init, np.zeros(40), np.eye(40), w.alpha, w.beta)
    assert ei.value.code == -7
    assert bsb.cluster(w.Y, 1, csr)["z"].shape == (1, w.n)      # q == 1 shortcut (R/spatialCluster.R:89-91)


def test_header_is_plain_c_and_a_c_program_can_bind_the_library(tmp_path):
    """The drop-in boundary is a C ABI: include/bsb200.h must compile as C99 (no C++/torch types), the ctypes mirror
    must agree with the header on the struct sizes, and a C program linked against libbsb200.so must be able to call it."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "bsb200.h"
int main(void) {
  bsb200_cluster_args a; bsb200_cluster_out o; bsb200_deconv_args da; bsb200_deconv_out dout;
  memset(&a, 0, sizeof a); memset(&o, 0, sizeof o); memset(&da, 0, sizeof da); memset(&dout, 0, sizeof dout);
  printf("%zu %zu %zu %zu\n", sizeof a, sizeof o, sizeof da, sizeof dout);
  printf("%d\n", bsb200_version());
  /* host-only entry point: neighbours of a 2 x 2 square lattice */
  int32_t row[4] = {1, 1, 2, 2}, col[4] = {1, 2, 1, 2}, idx[16];
  int64_t ptr[5], nnz = 0;
  int rc = bsb200_neighbors_lattice(row, col, 4, BSB200_PLATFORM_SQUARE, ptr, idx, 16, &nnz);
  printf("%d %lld\n", rc, (long long) nnz);
  /* invalid arguments are rejected before any device work, with a message */
  rc = bsb200_cluster(&a, &o);
  printf("%d %s\n", rc, bsb200_last_error());
  return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "bayesspace_b200")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lbsb200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    sizes = [int(v) for v in out[0].split()]
    from bayesspace_b200 import _lib
    assert sizes == [ctypes.sizeof(_lib.ClusterArgs), ctypes.sizeof(_lib.ClusterOut), ctypes.sizeof(_lib.DeconvArgs),
                     ctypes.sizeof(_lib.DeconvOut)]
    assert int(out[1]) > 0
    assert out[2] == "0 8"                     # 4 spots x 2 neighbours each
    rc, msg = out[3].split(" ", 1)
    assert int(rc) == _lib.ERR_INVALID and len(msg) > 0


# ---- the product's graph builders against restatements written from the R / C++ sources ---------------------
def _lists(ptr, idx):
    return [np.asarray(idx[ptr[j]:ptr[j + 1]]) for j in range(len(ptr) - 1)]


@pytest.mark.parametrize("platform", ["Visium", "ST"])
def test_lattice_neighbors_match_the_R_restatement(bsb, oracle, platform):
    """.find_neighbors restated with pandas merges exactly as R/spatialCluster.R:227-302 does it, on a lattice with
    holes and a duplicated coordinate: product (host code of libbsb200.so), oracle and restatement must agree on the
    lists, their order and the comma strings."""
    import spec_numpy
    from bayesspace_b200 import synth
    rng = np.random.default_rng(4)
    if platform == "Visium":
        r, c = synth.hex_lattice(14, 11)
    else:
        r, c = synth.square_lattice(12, 13)
    keep = rng.random(len(r)) > 0.15                      # holes in the tissue
    r, c = r[keep], c[keep]
    r, c = np.append(r, r[7]), np.append(c, c[7])         # two barcodes on one coordinate
    r, c = np.append(r, r.max() + 40), np.append(c, c.max() + 40)   # a spot without neighbours (NA entry)
    perm = rng.permutation(len(r))                        # spots are not stored in lattice order
    r, c = r[perm].astype(np.int32), c[perm].astype(np.int32)
    want, want_str = spec_numpy.find_neighbors_lattice(r, c, platform)
    strings, (ptr, idx) = bsb.neighbors._find_neighbors(r, c, platform)
    optr, oidx = oracle.lattice_neighbors(r, c, platform)
    got, ogot = _lists(ptr, idx), _lists(optr, oidx)
    assert len(got) == len(want) == len(ogot)
    for j in range(len(want)):
        assert np.array_equal(got[j], want[j]) and np.array_equal(ogot[j], want[j]), j
    assert list(strings) == want_str
    assert sum(s is None for s in want_str) >= 1            # isolated spots keep an NA entry / empty list


def test_radius_neighbors_match_the_R_restatement(bsb, oracle):
    import spec_numpy
    rng = np.random.default_rng(9)
    pos = np.round(rng.uniform(0, 12, size=(300, 2)), 1)     # ties at the radius occur on the 0.1 grid
    for method, radius in (("manhattan", 1.5), ("euclidean", 1.3), ("manhattan", 0.05)):
        want = spec_numpy.find_neighbors_radius(pos, radius, method)
        got = bsb.find_neighbors(pos, radius, method)
        ogot = oracle.find_neighbors(pos, radius, method)
        for j in range(len(want)):
            assert np.array_equal(np.asarray(got[j]), want[j]) and np.array_equal(np.asarray(ogot[j]), want[j]), (method, j)


@pytest.mark.parametrize("platform,k", [("Visium", None), ("ST", 3), ("ST", 2)])
def test_subspot_neighbors_match_the_cpp_restatement(bsb, oracle, platform, k):
    """find_subspot_neighbors (src/cluster.cpp:161-215) including the candidate ORDER, which fixes df_j's columns."""
    import spec_numpy
    from bayesspace_b200 import synth
    if platform == "Visium":
        r, c = synth.hex_lattice(9, 8)
        pos = np.c_[c.astype(float), r * np.sqrt(3.0)]
        xd, yd = 1.0, np.sqrt(3.0)
    else:
        r, c = synth.square_lattice(8, 9)
        pos = np.c_[c.astype(float), r.astype(float)]
        xd, yd = 1.0, 1.0
    _, (ptr, idx) = bsb.neighbors._find_neighbors(r, c, platform)
    dc = synth.deconv_inputs(np.random.default_rng(0).normal(size=(len(r), 2)), pos, xd, yd,
                             np.ones(len(r), dtype=np.int32), platform=platform,
                             nsubspots_per_edge=k or 3)
    want = spec_numpy.find_subspot_neighbors(_lists(ptr, idx), dc["positions2"], dc["dist"])
    sp, si = bsb.find_subspot_neighbors((ptr, idx), dc["positions2"], dc["dist"])
    op, oi = oracle.find_subspot_neighbors((ptr, idx), dc["positions2"], dc["dist"])
    got, ogot = _lists(sp, si), _lists(op, oi)
    for s in range(len(want)):
        assert np.array_equal(got[s], want[s]) and np.array_equal(ogot[s], want[s]), s
