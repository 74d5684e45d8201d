"""Pins the oracle (oracle/spmv_oracle.c) against (a) the committed golden
vectors, which are inputs/outputs of the reference's own tests and CPU code
(tests/golden/make_golden.py), and (b) the unmodified reference library
(oracle/_ref) on fresh seeded inputs when it is present."""
import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_known_answers(orc, golden):
    k = golden.known
    rp, ci, va = orc.csr_from_dense(k["readme_dense"])
    assert np.array_equal(orc.spmv_csr(3, rp, ci, va, np.ones(3, np.float32)), k["readme_y_ones"])
    w, ec, ev = orc.ell_from_csr(3, rp, ci, va)
    assert w == 2 and np.array_equal(ec, k["readme_ell_cols"])
    # design.md 3x4 layout example
    dense = np.zeros((3, 4), np.float32)
    dense[0, 0], dense[0, 2], dense[1, 1], dense[1, 2], dense[2, 3] = 1, 2, 3, 4, 5
    rp, ci, va = orc.csr_from_dense(dense)
    assert np.array_equal(rp, k["design_row_ptrs"]) and np.array_equal(ci, k["design_cols"])
    assert np.array_equal(va, k["design_values"])
    w, ec, ev = orc.ell_from_dense(dense)
    assert np.array_equal(ec, k["design_ell_cols"]) and np.array_equal(ev, k["design_ell_values"])
    # tests/test_spmv.cu:161-186 (5*2=10) and :188-218 ([3,0,7])
    rp, ci, va = orc.csr_from_dense(np.array([[5.0]], np.float32))
    assert orc.spmv_csr(1, rp, ci, va, np.array([2.0], np.float32))[0] == 10.0
    rp, ci, va = orc.csr_from_dense(k["unit_dense"])
    assert np.array_equal(orc.spmv_csr(3, rp, ci, va, np.ones(3, np.float32)), k["unit_y_ones"])


def test_golden_property_cases(orc, golden):
    g = golden.spmv
    for it in range(int(g["n_cases"][0])):
        p = f"c{it}_"
        dense, x = g[p + "dense"], g[p + "x"]
        rows, cols = dense.shape
        rp, ci, va = orc.csr_from_dense(dense)
        assert np.array_equal(rp, g[p + "row_ptrs"]) and np.array_equal(ci, g[p + "col_indices"])
        assert np.array_equal(bits(va), bits(g[p + "values"]))
        assert np.array_equal(bits(orc.spmv_csr(rows, rp, ci, va, x)), bits(g[p + "y_csr"]))
        s = orc.stats(rows, len(va), rp)
        exp = g[p + "stats"]
        assert (np.float32(s.avg), s.max, s.min, np.float32(s.skew)) == (np.float32(exp[0]), int(exp[1]), int(exp[2]), np.float32(exp[3]))
        assert list(orc.auto_config(rows, cols, len(va), rp)) == [int(g[p + "selector"][0]), int(g[p + "selector"][1]), bool(g[p + "selector"][2])]
        for w, ec, ev in (orc.ell_from_dense(dense), orc.ell_from_csr(rows, rp, ci, va)):
            assert w == int(g[p + "ell_width"][0])
            assert np.array_equal(ec, g[p + "ell_cols"]) and np.array_equal(bits(ev), bits(g[p + "ell_values"]))
        assert np.array_equal(bits(orc.spmv_ell(rows, w, ec, ev, x)), bits(g[p + "y_ell"]))
        assert np.array_equal(orc.csr_to_dense(rows, cols, rp, ci, va), dense)
        assert np.array_equal(orc.ell_to_dense(rows, cols, w, ec, ev), dense)


def test_golden_config1(orc, golden):
    g = golden.c1
    rp, ci, va, x = g["row_ptrs"], g["col_indices"], g["values"], g["x"]
    assert len(va) == 100232
    assert np.array_equal(bits(orc.spmv_csr(10000, rp, ci, va, x)), bits(g["y_csr"]))
    s = orc.stats(10000, len(va), rp)
    assert (s.max, s.min, np.float32(s.skew)) == (24, 1, np.float32(12.0))
    assert orc.auto_config(10000, 10000, len(va), rp) == (2, 256, False)  # MERGE_PATH, no texture


def test_f64_variant_consistent(orc, golden):
    g = golden.c1
    y64, scale = orc.spmv_csr_f64(10000, g["row_ptrs"], g["col_indices"], g["values"], g["x"])
    assert np.all(np.abs(y64 - g["y_csr"].astype(np.float64)) <= 1e-5 * scale + 1e-30)


def test_against_reference_library(orc, ref):
    """Fresh seeded inputs through the unmodified reference vs the restatement."""
    ref.rng_seed(1234)
    for _ in range(40):
        rows, cols = ref.rng_int(1, 120), ref.rng_int(1, 120)
        dense = ref.rng_dense(rows, cols, ref.rng_float(0.0, 0.5))
        x = ref.rng_vector(cols)
        h, st = ref.csr_from_dense(dense)
        r, c, nnz, rp, ci, va = ref.csr_fields(h)
        orp, oci, ova = orc.csr_from_dense(dense)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(bits(va), bits(ova))
        assert np.array_equal(bits(ref.spmv_cpu_csr(h, x, rows)), bits(orc.spmv_csr(rows, rp, ci, va, x)))
        a, mx, mn, sk = ref.stats(h)
        s = orc.stats(rows, nnz, rp)
        assert (np.float32(a), mx, mn, np.float32(sk)) == (np.float32(s.avg), s.max, s.min, np.float32(s.skew))
        assert ref.auto_config(h) == orc.auto_config(rows, cols, nnz, rp)
        he, _ = ref.ell_from_csr(h)
        _, _, w, ec, ev = ref.ell_fields(he)
        ow, oec, oev = orc.ell_from_csr(rows, rp, ci, va)
        assert w == ow and np.array_equal(ec, oec) and np.array_equal(bits(ev), bits(oev))
        assert np.array_equal(bits(ref.spmv_cpu_ell(he, x, rows)), bits(orc.spmv_ell(rows, w, ec, ev, x)))
        for _k in range(10):
            rr, cc = ref.rng_int(-1, rows), ref.rng_int(-1, cols)
            assert ref.L.ref_csr_get_element(h, rr, cc) == orc.csr_get_element(rows, cols, rp, ci, va, rr, cc)
            assert ref.L.ref_ell_get_element(he, rr, cc) == orc.ell_get_element(rows, cols, w, ec, ev, rr, cc)
        ref.L.ref_csr_destroy(h)
        ref.L.ref_ell_destroy(he)


def test_bandwidth_bytes(orc):
    # SURVEY 8d concrete figures
    assert orc.bytes_csr(16777216, 16777216, 83869696) == 872284164
    assert orc.bytes_ell(16777216, 16777216, 5) == 805306368
    assert orc.bytes_csr(16777216, 16777216, 268435456) == 2348810244
    assert orc.bytes_csr(10000, 10000, 100232) == 921860


def test_pagerank_restatement_invariants(orc, golden):
    """Properties the reference's own PageRank tests assert (tests/test_pagerank.cu:42-72)
    plus agreement between the literal fp32 and the f64-accumulator restatements."""
    g = golden.pr
    for it in range(int(g["n_cases"][0])):
        p = f"p{it}_"
        n = int(g[p + "n"][0])
        rp, ci, va = g[p + "row_ptrs"], g[p + "col_indices"], g[p + "values"]
        r32, it32, res32, conv32 = orc.pagerank_f32(n, n, rp, ci, va, 0.85, 1e-5, 50)
        assert np.all(r32 >= 0) and abs(float(r32.sum()) - 1.0) < 1e-4
        assert conv32 or it32 == 50
        if conv32:
            assert res32 < 1e-5
        r64, it64, l2, l1, conv64 = orc.pagerank_f64(n, n, rp, ci, va, 0.85, 1e-5, 50, fixed_it=it32)
        assert np.abs(r64.astype(np.float64) - r32).sum() < 1e-5
    # 3-cycle (tests/test_pagerank.cu:140-164): equal ranks, converged
    dense = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]], np.float32)
    rp, ci, va = orc.csr_from_dense(dense)
    r, iters, res, conv = orc.pagerank_f32(3, 3, rp, ci, va)
    assert conv and np.allclose(r, 1.0 / 3.0, atol=1e-4)


def test_merge_path_search_brute_force(orc):
    rng = np.random.default_rng(7)
    for _ in range(30):
        rows = int(rng.integers(1, 40))
        lens = rng.integers(0, 6, rows)
        rp = np.zeros(rows + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        # brute-force merge of row-end items and nz items (row end first on ties)
        path = [(0, 0)]
        r = z = 0
        while r < rows or z < nnz:
            if r < rows and (z >= nnz or rp[r + 1] <= z):
                r += 1
            else:
                z += 1
            path.append((r, z))
        for d in range(rows + nnz + 1):
            assert orc.merge_path_search(d, rp, rows, nnz) == path[d]
