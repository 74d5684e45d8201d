"""Host side of the product (csr_* / ell_* / stats / selector / serialisation /
API host SpMV) through the C ABI, bit-exact against the oracle and the golden
vectors.  Mirrors the reference's tests/test_csr.cpp, test_ell.cpp,
test_kernel_selector.cpp, test_common.cpp (CPU-only cases)."""
import os

import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def random_dense(rng, rows, cols, density):
    d = np.zeros((rows, cols), np.float32)
    mask = rng.random((rows, cols)) < density
    d[mask] = rng.uniform(-10, 10, mask.sum()).astype(np.float32)
    return d


def test_error_strings(sp):
    # reference tests/test_common.cpp:8-18, verbatim
    expect = {0: "Success", -1: "Invalid matrix/vector dimension", -2: "CUDA memory allocation failed",
              -3: "CUDA memory copy failed", -4: "CUDA kernel launch failed", -5: "Invalid sparse matrix format",
              -6: "File I/O error", -7: "Out of memory", -8: "Invalid argument", -99: "Unknown error"}
    for code, text in expect.items():
        assert sp.spmv_error_string(code) == text


def test_csr_round_trip_and_oracle(sp, orc):
    # reference tests/test_csr.cpp:18-43 (+ bit-exact arrays vs the oracle)
    rng = np.random.default_rng(42)
    for _ in range(60):
        rows, cols = int(rng.integers(1, 101)), int(rng.integers(1, 101))
        dense = random_dense(rng, rows, cols, rng.uniform(0.01, 0.5))
        A = sp.csr_create(0, 0, 0)
        assert sp.csr_from_dense(A, dense, rows, cols) == 0
        rp, ci, va = sp.csr_arrays(A)
        orp, oci, ova = orc.csr_from_dense(dense)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(bits(va), bits(ova))
        rc, back = sp.csr_to_dense(A)
        assert rc == 0 and np.array_equal(back, dense)
        for _k in range(20):  # test_csr.cpp:47-76
            r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
            assert sp.csr_get_element(A, r, c) == dense[r, c]
        assert sp.csr_get_element(A, -1, 0) == 0.0 and sp.csr_get_element(A, rows, 0) == 0.0
        s, o = sp.csr_compute_stats(A), orc.stats(rows, len(va), rp)
        assert (bits(s.avg_nnz_per_row), s.max_nnz_per_row, s.min_nnz_per_row, bits(s.skewness)) == \
               (bits(o.avg), o.max, o.min, bits(o.skew))
        assert np.array_equal(bits(sp.spmv_cpu_csr(A, np.ones(cols, np.float32))),
                              bits(orc.spmv_csr(rows, rp, ci, va, np.ones(cols, np.float32))))
        sp.csr_destroy(A)


def test_csr_edge_cases(sp):
    # reference tests/test_csr.cpp:130-166: empty, all-zero, 1x1; negative dims; bad args
    A = sp.csr_create(0, 0, 0)
    assert A and A.contents.nnz == 0 and not A.contents.values and A.contents.row_ptrs[0] == 0
    sp.csr_destroy(A)
    assert not sp.csr_create(-1, 0, 0) and not sp.csr_create(0, -1, 0) and not sp.csr_create(0, 0, -1)
    A = sp.csr_create(0, 0, 0)
    assert sp.csr_from_dense(A, np.zeros((4, 5), np.float32), 4, 5) == 0
    assert A.contents.nnz == 0 and not A.contents.values and list(sp.csr_arrays(A)[0]) == [0] * 5
    assert sp.csr_from_dense(A, np.array([[7.0]], np.float32), 1, 1) == 0
    assert A.contents.nnz == 1 and sp.csr_get_element(A, 0, 0) == 7.0
    assert sp.csr_from_dense(A, np.zeros(1, np.float32), 0, 1) == -8
    assert sp.csr_from_dense(A, np.zeros(1, np.float32), 1, -1) == -8
    # -0.0 is dropped, NaN and inf are kept (v != 0.0f)
    d = np.array([[-0.0, np.nan, np.inf, 0.0]], np.float32)
    assert sp.csr_from_dense(A, d, 1, 4) == 0 and list(sp.csr_arrays(A)[1]) == [1, 2]
    sp.csr_destroy(A)
    sp.csr_destroy(None)
    s = sp.csr_compute_stats(None)
    assert (s.avg_nnz_per_row, s.max_nnz_per_row, s.min_nnz_per_row, s.skewness) == (0.0, 0, 0, 0.0)


def test_ell_round_trip_padding_and_oracle(sp, orc):
    # reference tests/test_ell.cpp:19-108
    rng = np.random.default_rng(43)
    for _ in range(60):
        rows, cols = int(rng.integers(1, 101)), int(rng.integers(1, 101))
        dense = random_dense(rng, rows, cols, rng.uniform(0.01, 0.5))
        E = sp.ell_create(0, 0, 0)
        assert sp.ell_from_dense(E, dense, rows, cols) == 0
        ec, ev = sp.ell_arrays(E)
        w, oec, oev = orc.ell_from_dense(dense)
        assert E.contents.max_nnz_per_row == w
        assert np.array_equal(ec, oec) and np.array_equal(bits(ev), bits(oev))
        rc, back = sp.ell_to_dense(E)
        assert rc == 0 and np.array_equal(back, dense)
        row_nnz = (dense != 0).sum(axis=1)
        for r in range(rows):  # padding is col=-1, val=0 beyond the row's nnz
            for k in range(int(row_nnz[r]), w):
                assert ec[sp.ell_index(r, k, rows)] == -1 and ev[sp.ell_index(r, k, rows)] == 0.0
        A = sp.csr_create(0, 0, 0)
        sp.csr_from_dense(A, dense, rows, cols)
        E2 = sp.ell_create(0, 0, 0)
        assert sp.ell_from_csr(E2, A) == 0
        ec2, ev2 = sp.ell_arrays(E2)
        assert np.array_equal(ec, ec2) and np.array_equal(bits(ev), bits(ev2))
        x = rng.uniform(-10, 10, cols).astype(np.float32)
        assert np.array_equal(bits(sp.spmv_cpu_ell(E, x)), bits(orc.spmv_ell(rows, w, oec, oev, x)))
        for _k in range(10):
            r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
            assert sp.ell_get_element(E, r, c) == dense[r, c]
        for m in (E, E2):
            sp.ell_destroy(m)
        sp.csr_destroy(A)
    assert sp.ell_index(3, 2, 10) == 23
    E = sp.ell_create(4, 4, 2)
    assert list(sp.ell_arrays(E)[0]) == [-1] * 8 and not np.any(sp.ell_arrays(E)[1])
    sp.ell_destroy(E)
    assert not sp.ell_create(-1, 1, 1)


def test_serialisation_round_trip_and_byte_format(sp, tmp_path):
    # reference tests/test_csr.cpp:80-127, tests/test_ell.cpp:112-150 + the on-disk layout
    rng = np.random.default_rng(44)
    dense = random_dense(rng, 17, 23, 0.2)
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, dense, 17, 23)
    f = tmp_path / "a.bin"
    assert sp.csr_serialize(A, f) == 0
    rp, ci, va = sp.csr_arrays(A)
    raw = f.read_bytes()
    expect = np.array([17, 23, len(va)], np.int32).tobytes() + va.tobytes() + ci.tobytes() + rp.tobytes()
    assert raw == expect
    B = sp.csr_create(0, 0, 0)
    assert sp.csr_deserialize(B, f) == 0
    for a, b in zip(sp.csr_arrays(A), sp.csr_arrays(B)):
        assert a.tobytes() == b.tobytes()
    assert sp.csr_deserialize(B, tmp_path / "missing.bin") == -6
    (tmp_path / "short.bin").write_bytes(raw[: len(raw) // 2])
    assert sp.csr_deserialize(B, tmp_path / "short.bin") == -6
    E = sp.ell_create(0, 0, 0)
    sp.ell_from_csr(E, A)
    g = tmp_path / "e.bin"
    assert sp.ell_serialize(E, g) == 0
    ec, ev = sp.ell_arrays(E)
    assert g.read_bytes() == np.array([17, 23, E.contents.max_nnz_per_row], np.int32).tobytes() + ev.tobytes() + ec.tobytes()
    E2 = sp.ell_create(0, 0, 0)
    assert sp.ell_deserialize(E2, g) == 0
    assert sp.ell_arrays(E2)[0].tobytes() == ec.tobytes() and sp.ell_arrays(E2)[1].tobytes() == ev.tobytes()
    for m in (A, B):
        sp.csr_destroy(m)
    for m in (E, E2):
        sp.ell_destroy(m)


def test_serialisation_interchange_with_reference(sp, ref, tmp_path):
    """Files written by the reference load here and vice versa (bit-exact)."""
    rng = np.random.default_rng(45)
    dense = random_dense(rng, 31, 19, 0.3)
    h, _ = ref.csr_from_dense(dense)
    f = str(tmp_path / "ref.bin").encode()
    assert ref.L.ref_csr_serialize(h, f) == 0
    A = sp.csr_create(0, 0, 0)
    assert sp.csr_deserialize(A, f.decode()) == 0
    _, _, _, rp, ci, va = ref.csr_fields(h)
    mine = sp.csr_arrays(A)
    assert np.array_equal(mine[0], rp) and np.array_equal(mine[1], ci) and np.array_equal(bits(mine[2]), bits(va))
    g = tmp_path / "mine.bin"
    sp.csr_serialize(A, g)
    assert g.read_bytes() == open(f, "rb").read()
    sp.csr_destroy(A)
    ref.L.ref_csr_destroy(h)


def test_golden_cases_through_product(sp, golden):
    g = golden.spmv
    for it in range(int(g["n_cases"][0])):
        p = f"c{it}_"
        dense, x = g[p + "dense"], g[p + "x"]
        rows, cols = dense.shape
        A = sp.csr_create(0, 0, 0)
        sp.csr_from_dense(A, dense, rows, cols)
        rp, ci, va = sp.csr_arrays(A)
        assert np.array_equal(rp, g[p + "row_ptrs"]) and np.array_equal(ci, g[p + "col_indices"])
        assert np.array_equal(bits(va), bits(g[p + "values"]))
        assert np.array_equal(bits(sp.spmv_cpu_csr(A, x)), bits(g[p + "y_csr"]))
        c = sp.spmv_auto_config(A)
        assert [c.kernel_type, c.block_size, int(c.use_texture)] == list(g[p + "selector"])
        E = sp.ell_create(0, 0, 0)
        sp.ell_from_csr(E, A)
        assert E.contents.max_nnz_per_row == int(g[p + "ell_width"][0])
        assert np.array_equal(sp.ell_arrays(E)[0], g[p + "ell_cols"])
        assert np.array_equal(bits(sp.spmv_cpu_ell(E, x)), bits(g[p + "y_ell"]))
        sp.csr_destroy(A)
        sp.ell_destroy(E)


def test_selector_reference_cases_and_policy(sp, orc, golden):
    # reference tests/test_kernel_selector.cpp:53-137
    d = np.zeros((10, 10), np.float32)
    d[np.arange(10), 0] = 1.0
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, d, 10, 10)
    assert sp.spmv_auto_config(A).kernel_type == sp.SCALAR_CSR
    d = np.zeros((10, 10), np.float32)
    d[:, :5] = 1.0
    sp.csr_from_dense(A, d, 10, 10)
    assert sp.spmv_auto_config(A).kernel_type == sp.VECTOR_CSR
    d = np.zeros((12, 64), np.float32)
    d[0, :] = 1.0
    d[1:, 0] = 1.0
    d[1:, 1] = 1.0
    d[1:, 2] = 1.0
    d[1:, 3] = 1.0  # avg >= 4, skew = 64/5 >= 10
    sp.csr_from_dense(A, d, 12, 64)
    assert sp.spmv_auto_config(A).kernel_type == sp.MERGE_PATH
    c = sp.spmv_auto_config(A)
    assert c.block_size == 256 and 32 <= c.block_size <= 1024 and c.block_size % 32 == 0 and not c.use_texture
    sp.csr_destroy(A)
    # use_texture = num_cols > 10000 (strict), config 1 -> MERGE_PATH
    g = golden.c1
    A = sp.csr_from_arrays(10000, 10000, g["row_ptrs"], g["col_indices"], g["values"])
    c = sp.spmv_auto_config(A)
    assert (c.kernel_type, c.block_size, c.use_texture) == (sp.MERGE_PATH, 256, False)
    A.contents.num_cols = 10001
    assert sp.spmv_auto_config(A).use_texture
    sp.csr_destroy(A)
    # random shapes: reference policy == oracle decision, bit for bit; auto_config identical here (no outliers)
    rng = np.random.default_rng(46)
    for _ in range(200):
        rows = int(rng.integers(1, 300))
        lens = rng.integers(0, int(rng.integers(1, 40)), rows)
        if rng.random() < 0.3:
            lens[int(rng.integers(0, rows))] = int(rng.integers(50, 400))
        rp = np.zeros(rows + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        nnz, cols = int(rp[-1]), int(rng.integers(1, 20000))
        A = sp.csr_from_arrays(rows, cols, rp, np.zeros(nnz, np.int32), np.ones(nnz, np.float32))
        kt, bs, tex = orc.auto_config(rows, cols, nnz, rp)
        for c in (sp.spmv_reference_policy(A), sp.spmv_auto_config(A)):
            assert (c.kernel_type, c.block_size, c.use_texture) == (kt, bs, tex)
        sp.csr_destroy(A)


def test_selector_outlier_override(sp, orc):
    """The one documented divergence: avg < 4 with a row > 65536 nnz -> MERGE_PATH
    (the reference policy says SCALAR; BASELINE config 3)."""
    rows = 200000
    lens = np.full(rows, 3, np.int64)
    lens[1000] = 70000
    rp = np.zeros(rows + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    nnz = int(rp[-1])
    A = sp.csr_from_arrays(rows, rows, rp, np.zeros(nnz, np.int32), np.ones(nnz, np.float32))
    assert orc.auto_config(rows, rows, nnz, rp)[0] == sp.SCALAR_CSR
    assert sp.spmv_reference_policy(A).kernel_type == sp.SCALAR_CSR
    assert sp.spmv_auto_config(A).kernel_type == sp.MERGE_PATH
    sp.csr_destroy(A)


def test_bandwidth_model_and_json(sp, orc, ref):
    assert sp.csr_bytes(16777216, 16777216, 83869696) == orc.bytes_csr(16777216, 16777216, 83869696) == 872284164
    assert sp.ell_bytes(16777216, 16777216, 5) == orc.bytes_ell(16777216, 16777216, 5) == 805306368
    # elapsed 0 -> all-zero metrics, before any device query (reference tests/test_bandwidth.cu:100-113)
    A = sp.csr_create(4, 4, 0)
    m = sp.compute_bandwidth_csr(A, 0.0)
    assert (m.theoretical_bandwidth_gb_s, m.achieved_bandwidth_gb_s, m.efficiency) == (0.0, 0.0, 0.0)
    sp.csr_destroy(A)
    # JSON: byte-identical to the reference writer, and both readers agree (tests/test_benchmark.cu:65-103,151-170)
    r = sp.BenchmarkResult()
    r.name = b"CSR SpMV"
    vals = [1.234567, 456.789012, 321.5, 1.234567, 1.1, 1.4, 0.0123456]
    (r.execution_time_ms, r.gflops, r.bandwidth_gb_s, r.avg_time_ms, r.min_time_ms, r.max_time_ms,
     r.stddev_time_ms) = vals
    r.num_runs = 20
    text = sp.benchmark_to_json(r)
    import ctypes as C
    buf = C.create_string_buffer(4096)
    f7 = (C.c_float * 7)(*vals)
    n = ref.L.ref_benchmark_to_json(b"CSR SpMV", f7, 20, buf, 4096)
    assert n > 0 and buf.value.decode() == text
    for key in ("name", "execution_time_ms", "gflops", "bandwidth_gb_s", "avg_time_ms", "min_time_ms",
                "max_time_ms", "stddev_time_ms", "num_runs"):
        assert f'"{key}":' in text
    back = sp.benchmark_from_json(text)
    out7, nr = (C.c_float * 7)(), C.c_int(0)
    ref.L.ref_benchmark_from_json(text.encode(), out7, C.byref(nr))
    mine = [back.execution_time_ms, back.gflops, back.bandwidth_gb_s, back.avg_time_ms, back.min_time_ms,
            back.max_time_ms, back.stddev_time_ms]
    assert mine == list(out7) and back.num_runs == nr.value == 20


def test_top_k_host(sp, orc):
    # reference tests/test_pagerank.cu:81-137: descending, dominates all non-members; identical up to ties
    rng = np.random.default_rng(47)
    for _ in range(20):
        n = int(rng.integers(1, 400))
        ranks = rng.random(n).astype(np.float32)
        if n > 4:
            ranks[1] = ranks[3]  # a tie
        k = int(rng.integers(1, n + 5))
        ids, vals = sp.pagerank_top_k(ranks, k)
        oids, ovals = orc.top_k(ranks, k)
        assert len(ids) == min(k, n)
        assert np.array_equal(vals, ovals)  # same rank values position by position
        assert np.array_equal(ranks[ids], vals) and len(set(ids.tolist())) == len(ids)
        assert np.all(np.diff(vals) <= 0)
        rest = np.setdiff1d(np.arange(n), ids)
        if len(rest):
            assert ranks[rest].max() <= vals.min()
