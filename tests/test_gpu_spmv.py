"""Parity of the sm_100a SpMV kernels with the reference, through the C ABI.

Oracle = spmv_cpu_csr / spmv_cpu_ell restated in oracle/spmv_oracle.c (pinned
by tests/test_oracle_pinned.py) and its f64-accumulating variant for the
north_star tolerance |y - y_ref| <= 1e-5 * sum_j |a_ij x_j| per row.
SCALAR_CSR and ELL keep the reference CPU path's operation order, so they are
checked BIT-EXACT against it."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from gpu_helpers import GpuCSR, assert_within_tolerance, bits, run_csr

pytestmark = pytest.mark.gpu
KERNELS = ("SCALAR_CSR", "VECTOR_CSR", "MERGE_PATH")


def gen_mod():
    import gpu_spmv_b200.gen as gen
    return gen


def check_all_csr_kernels(sp, orc, dev, rows, cols, rp, ci, va, x, what=""):
    A = GpuCSR(sp, rows, cols, rp, ci, va)
    y_cpu = orc.spmv_csr(rows, A.rp, A.ci, A.va, x)
    y64, scale = orc.spmv_csr_f64(rows, A.rp, A.ci, A.va, x)
    for k, name in enumerate(KERNELS):
        y, res = run_csr(sp, A.mat, x, k, dev, rows)
        assert_within_tolerance(y, y64, scale, f"{what} {name}")
        if name == "SCALAR_CSR":
            assert np.array_equal(bits(y), bits(y_cpu)), f"{what}: SCALAR_CSR not bit-identical to spmv_cpu_csr"
        assert res.elapsed_ms >= 0
    A.close()


def check_ell(sp, orc, dev, rows, cols, rp, ci, va, x, what=""):
    A = sp.csr_from_arrays(rows, cols, rp, ci, va)
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr(E, A) == 0 and sp.ell_to_gpu(E) == 0
    w = E.contents.max_nnz_per_row
    ec, ev = sp.ell_arrays(E)
    d_x = torch.as_tensor(np.ascontiguousarray(x, np.float32)).to(dev)
    d_y = torch.full((max(rows, 1),), float("nan"), dtype=torch.float32, device=dev)
    res = sp.spmv_ell(E, d_x, d_y, None, cols)
    assert res.error_code == 0
    y = d_y[:rows].cpu().numpy()
    assert np.array_equal(bits(y), bits(orc.spmv_ell(rows, w, ec, ev, x))), f"{what}: ELL not bit-identical"
    true_nnz = int((ec >= 0).sum())
    if res.elapsed_ms > 0:
        assert abs(res.gflops - 2.0 * true_nnz / (res.elapsed_ms * 1e6)) <= 1e-3 * max(res.gflops, 1e-9)
    sp.csr_destroy(A)
    sp.ell_destroy(E)


def test_golden_property_cases(sp, orc, cuda, golden):
    """The inputs of the reference's SpMVPropertyTest.CSRCorrectness / ELLCorrectness
    (tests/test_spmv.cu:40-118) against the reference's own CPU outputs."""
    g = golden.spmv
    for it in range(int(g["n_cases"][0])):
        p = f"c{it}_"
        dense, x = g[p + "dense"], g[p + "x"]
        rows, cols = dense.shape
        A = sp.csr_create(0, 0, 0)
        sp.csr_from_dense(A, dense, rows, cols)
        assert sp.csr_to_gpu(A) == 0
        y64, scale = orc.spmv_csr_f64(rows, g[p + "row_ptrs"], g[p + "col_indices"], g[p + "values"], x)
        for k, name in enumerate(KERNELS):
            y, res = run_csr(sp, A, x, k, cuda, rows)
            assert_within_tolerance(y, y64, scale, f"case {it} {name}")
            if k == 0:
                assert np.array_equal(bits(y), bits(g[p + "y_csr"]))
                # the reference's own criterion (tests/test_spmv.cu:18-35): 1e-6 relative
                ref_y = g[p + "y_csr"]
                mx = np.maximum(np.abs(ref_y), np.abs(y))
                assert np.all((np.abs(ref_y - y) <= 1e-6 * mx) | (mx < 1e-10))
        E = sp.ell_create(0, 0, 0)
        sp.ell_from_dense(E, dense, rows, cols)
        assert sp.ell_to_gpu(E) == 0
        d_x = torch.as_tensor(x).to(cuda)
        d_y = torch.empty(rows, dtype=torch.float32, device=cuda)
        assert sp.spmv_ell(E, d_x, d_y, None, cols).error_code == 0
        assert np.array_equal(bits(d_y.cpu().numpy()), bits(g[p + "y_ell"]))
        sp.csr_destroy(A)
        sp.ell_destroy(E)


def test_config1_random_10k(sp, orc, cuda, golden):
    """BASELINE config 1 (reference generator, seed 42): every kernel, the selector's choice, ELL."""
    g = golden.c1
    rp, ci, va, x = g["row_ptrs"], g["col_indices"], g["values"], g["x"]
    A = GpuCSR(sp, 10000, 10000, rp, ci, va)
    y64, scale = orc.spmv_csr_f64(10000, rp, ci, va, x)
    cfg = sp.spmv_auto_config(A.mat)
    assert cfg.kernel_type == sp.MERGE_PATH
    for k in (0, 1, 2):
        y, res = run_csr(sp, A.mat, x, k, cuda, 10000)
        assert_within_tolerance(y, y64, scale, KERNELS[k])
        if k == 0:
            assert np.array_equal(bits(y), bits(g["y_csr"]))
        assert res.gflops > 0 and res.bandwidth_gb_s > 0
        assert abs(res.bandwidth_gb_s - orc.achieved_gbs(921860, res.elapsed_ms)) <= 1e-3 * res.bandwidth_gb_s
    A.close()
    check_ell(sp, orc, cuda, 10000, 10000, rp, ci, va, x, "config1")


def test_known_answers_and_edge_cases(sp, orc, cuda):
    # tests/test_spmv.cu:161-186 and :188-218 with the default (NULL) config
    A = GpuCSR(sp, 1, 1, [0, 1], [0], [5.0])
    y, _ = run_csr(sp, A.mat, np.array([2.0], np.float32), None, cuda, 1)
    assert y[0] == 10.0
    A.close()
    rp, ci, va = orc.csr_from_dense(np.array([[1, 2, 0], [0, 0, 0], [3, 0, 4]], np.float32))
    for k in (None, 0, 1, 2):
        A = GpuCSR(sp, 3, 3, rp, ci, va)
        y, _ = run_csr(sp, A.mat, np.ones(3, np.float32), k, cuda, 3)
        assert list(y) == [3.0, 0.0, 7.0]
        A.close()
    # unknown kernel type (ELL_KERNEL passed to spmv_csr) falls back to scalar (src/spmv_kernels.cu:287-288)
    A = GpuCSR(sp, 3, 3, rp, ci, va)
    y, _ = run_csr(sp, A.mat, np.ones(3, np.float32), sp.ELL_KERNEL, cuda, 3)
    assert list(y) == [3.0, 0.0, 7.0]
    # dimension mismatch is reported before anything runs
    d = torch.ones(8, device=cuda)
    assert sp.spmv_csr(A.mat, d, d, None, 4).error_code == -1
    A.close()
    # all-zero matrix (nnz == 0, rows > 0): csr_to_gpu leaves d_col_indices NULL, which the launcher
    # reports as INVALID_FORMAT exactly like the reference (src/spmv_kernels.cu:229-232) ...
    A = GpuCSR(sp, 37, 5, np.zeros(38, np.int32), [], [])
    assert sp.spmv_csr(A.mat, d, d, None, 5).error_code == -5
    A.close()
    # ... and with device arrays present every kernel must write zeros
    dummy_i, dummy_f = torch.zeros(4, dtype=torch.int32, device=cuda), torch.zeros(4, device=cuda)
    Z = sp.DeviceCSR(37, 5, torch.zeros(38, dtype=torch.int32, device=cuda), dummy_i, dummy_f, nnz=0)
    for k in (0, 1, 2):
        d_y = torch.full((37,), float("nan"), device=cuda)
        assert sp.spmv_csr(Z.ptr, torch.ones(5, device=cuda), d_y, sp.make_config(k), 5).error_code == 0
        assert not d_y.cpu().numpy().any()
    # empty matrix (0 x 0): success, nothing launched (documented divergence from the reference's KERNEL_LAUNCH)
    A = sp.csr_create(0, 0, 0)
    assert sp.csr_to_gpu(A) == 0
    d = torch.ones(1, device=cuda)
    r = sp.spmv_csr(A, d, d, None, 0)
    assert r.error_code in (0, -5)
    sp.csr_destroy(A)
    # ELL with zero width
    E = sp.ell_create(9, 9, 0)
    E.contents.d_values = d.data_ptr()
    E.contents.d_col_indices = d.data_ptr()
    d_y = torch.full((9,), float("nan"), device=cuda)
    assert sp.spmv_ell(E, torch.ones(9, device=cuda), d_y, None, 9).error_code == 0
    assert not d_y.cpu().numpy().any()
    E.contents.d_values = None
    E.contents.d_col_indices = None
    sp.ell_destroy(E)


@pytest.mark.parametrize("rows,cols,avg,skew,seed", [
    (1, 1, 1, 0.0, 1), (2, 7, 3, 0.0, 2), (255, 300, 2, 0.0, 3), (256, 256, 5, 0.0, 4), (1023, 999, 1, 0.0, 5),
    (1025, 4000, 3, 0.0, 6), (5000, 5000, 7, 0.0, 7), (4099, 6000, 12, 0.0, 8), (3001, 3001, 25, 0.0, 9),
    (2000, 9000, 50, 0.0, 10), (700, 20000, 150, 0.0, 11), (300, 30000, 700, 0.0, 12),
    (20000, 20000, 4, 0.05, 13), (50000, 50000, 2, 0.01, 14), (9000, 9000, 10, 0.2, 15), (100003, 65537, 3, 0.0, 16),
])
def test_random_shapes_all_kernels(sp, orc, cuda, rows, cols, avg, skew, seed):
    gen = gen_mod()
    rp, ci, va = gen.random_csr(rows, cols, avg, seed, "cpu", skew)
    x = gen.vector_pm1(cols, seed + 100, "cpu").numpy()
    check_all_csr_kernels(sp, orc, cuda, rows, cols, rp.numpy(), ci.numpy(), va.numpy(), x, f"random {rows}x{cols} avg {avg}")
    if avg <= 60:
        check_ell(sp, orc, cuda, rows, cols, rp.numpy(), ci.numpy(), va.numpy(), x, f"random {rows}x{cols}")


def test_outlier_rows_spanning_many_tiles(sp, orc, cuda):
    """Short rows plus rows far longer than a merge tile / product pass (config 3 in small),
    with empty rows before, between and after, and an outlier as the very last row."""
    gen = gen_mod()
    rows = 200000
    rp, ci, va = gen.short_rows_with_outliers_csr(rows, 43, "cpu", outlier_rows=[0, 777, 100000, rows - 1],
                                                  outlier_nnz=30011)
    x = gen.uniform_01_open_low(5, torch.arange(rows), 9).numpy()
    check_all_csr_kernels(sp, orc, cuda, rows, rows, rp.numpy(), ci.numpy(), va.numpy(), x, "outliers")
    A = sp.csr_from_arrays(rows, rows, rp.numpy(), ci.numpy(), va.numpy())
    assert sp.spmv_reference_policy(A).kernel_type == sp.SCALAR_CSR  # avg < 4 ...
    sp.csr_destroy(A)
    # one single huge row (all merge tiles carry into the same row), and huge first/last rows
    for lens in ([100000], [50000, 0, 0, 50000], [0, 0, 70000, 0], [3, 90000, 1]):
        lens = np.array(lens)
        rp = np.zeros(len(lens) + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        rng = np.random.default_rng(nnz)
        ci = np.sort(rng.integers(0, 5000, nnz)).astype(np.int32)
        va = rng.uniform(-1, 1, nnz).astype(np.float32)
        x = rng.uniform(-1, 1, 5000).astype(np.float32)
        check_all_csr_kernels(sp, orc, cuda, len(lens), 5000, rp, ci, va, x, f"huge rows {lens.tolist()}")


def test_unaligned_device_pointers(sp, orc, cuda):
    """d_values / d_col_indices / d_row_ptrs / y that are only 4-byte aligned take the
    scalar-load fall-backs (the API accepts any device pointer in the public struct fields)."""
    gen = gen_mod()
    rows = cols = 3000
    rp, ci, va = gen.random_csr(rows, cols, 6, 21, "cpu")
    x = gen.vector_pm1(cols, 22, "cpu")
    y64, scale = orc.spmv_csr_f64(rows, rp.numpy(), ci.numpy(), va.numpy(), x.numpy())
    pad = lambda t: torch.cat([torch.zeros(1, dtype=t.dtype), t]).to(cuda)[1:]  # noqa: E731
    d_rp, d_ci, d_va, d_x = pad(rp), pad(ci), pad(va), pad(x)
    assert d_va.data_ptr() % 16 == 4
    A = sp.DeviceCSR(rows, cols, d_rp, d_ci, d_va)
    d_y = torch.zeros(rows + 1, dtype=torch.float32, device=cuda)[1:]
    for k in (0, 1, 2):
        d_y.fill_(float("nan"))
        assert sp.spmv_csr(A.ptr, d_x, d_y, sp.make_config(k), cols).error_code == 0
        assert_within_tolerance(d_y.cpu().numpy(), y64, scale, f"unaligned {KERNELS[k]}")
    # ELL: odd row count (1 row/thread), even-not-multiple-of-4 (2 rows/thread), unaligned base
    for r in (3001, 3002, 3004):
        rp2, ci2, va2 = gen.random_csr(r, cols, 5, 23, "cpu")
        check_ell(sp, orc, cuda, r, cols, rp2.numpy(), ci2.numpy(), va2.numpy(), x.numpy(), f"ell rows {r}")
    w, ec, ev = orc.ell_from_csr(rows, rp.numpy(), ci.numpy(), va.numpy())
    d_ec, d_ev = pad(torch.as_tensor(ec)), pad(torch.as_tensor(ev))
    E = sp.DeviceELL(rows, cols, w, d_ec, d_ev)
    assert sp.spmv_ell(E.ptr, d_x, d_y, None, cols).error_code == 0
    assert np.array_equal(bits(d_y.cpu().numpy()), bits(orc.spmv_ell(rows, w, ec, ev, x.numpy())))


def test_laplacian_medium_vs_oracle(sp, orc, cuda):
    """BASELINE config 2 at 1024^2 (1 M rows): CSR (selector says VECTOR) and ELL vs the oracle."""
    gen = gen_mod()
    grid = 1024
    n = grid * grid
    rp, ci, va = gen.laplacian_2d_csr(grid, "cpu")
    x = gen.vector_pm1(n, 42, "cpu").numpy()
    A = sp.csr_from_arrays(n, n, rp.numpy(), ci.numpy(), va.numpy())
    cfg = sp.spmv_auto_config(A)
    assert cfg.kernel_type == sp.VECTOR_CSR and cfg.use_texture
    sp.csr_destroy(A)
    check_all_csr_kernels(sp, orc, cuda, n, n, rp.numpy(), ci.numpy(), va.numpy(), x, "laplacian 1024^2")
    check_ell(sp, orc, cuda, n, n, rp.numpy(), ci.numpy(), va.numpy(), x, "laplacian 1024^2")


def test_device_ell_assembly_matches_host(sp, orc, cuda):
    gen = gen_mod()
    rows, cols = 40001, 40001
    rp, ci, va = gen.random_csr(rows, cols, 6, 31, "cpu", 0.01)
    A = GpuCSR(sp, rows, cols, rp.numpy(), ci.numpy(), va.numpy())
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.mat) == 0
    w, ec, ev = orc.ell_from_csr(rows, A.rp, A.ci, A.va)
    assert E.contents.max_nnz_per_row == w
    # read the device arrays back through the API: host-allocated twin + ell_from_gpu
    H = sp.ell_create(rows, cols, w)
    H.contents.d_values, H.contents.d_col_indices = E.contents.d_values, E.contents.d_col_indices
    assert sp.ell_from_gpu(H) == 0
    hc, hv = sp.ell_arrays(H)
    assert np.array_equal(hc, ec) and np.array_equal(bits(hv), bits(ev))
    H.contents.d_values, H.contents.d_col_indices = None, None
    sp.ell_destroy(H)
    sp.ell_destroy(E)
    A.close()


def test_full_size_laplacian_properties(sp, cuda):
    """BASELINE config 2 at full size (4096^2 grid, 16.7 M rows), size-independent properties:
    A*1 has the closed form 4 - (#neighbours); SCALAR_CSR and ELL are bit-identical to each other
    (same per-row order); VECTOR / MERGE agree within tolerance; linearity A(2x) = 2 A x exactly."""
    gen = gen_mod()
    grid = 4096
    n = grid * grid
    rp, ci, va = gen.laplacian_2d_csr(grid, cuda)
    assert ci.numel() == 83869696
    A = sp.DeviceCSR(n, n, rp, ci, va)
    ones = torch.ones(n, device=cuda)
    y = torch.empty(n, device=cuda)
    i = torch.arange(n, device=cuda)
    gy, gx = i // grid, i % grid
    expect = 4.0 - ((gy > 0).float() + (gx > 0).float() + (gx < grid - 1).float() + (gy < grid - 1).float())
    for k in (0, 1, 2):
        y.fill_(float("nan"))
        assert sp.spmv_csr(A.ptr, ones, y, sp.make_config(k), n).error_code == 0
        assert torch.equal(y, expect), KERNELS[k]
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.ptr) == 0 and E.contents.max_nnz_per_row == 5
    x = gen.vector_pm1(n, 42, cuda)
    y_ell, y_sc, y_k = torch.empty(n, device=cuda), torch.empty(n, device=cuda), torch.empty(n, device=cuda)
    assert sp.spmv_ell(E, x, y_ell, None, n).error_code == 0
    assert sp.spmv_csr(A.ptr, x, y_sc, sp.make_config(0), n).error_code == 0
    assert torch.equal(y_ell, y_sc)
    scale = torch.empty(n, device=cuda)
    absA = sp.DeviceCSR(n, n, rp, ci, va.abs())
    assert sp.spmv_csr(absA.ptr, x.abs(), scale, sp.make_config(0), n).error_code == 0
    for k in (1, 2):
        assert sp.spmv_csr(A.ptr, x, y_k, sp.make_config(k), n).error_code == 0
        assert bool(((y_k - y_sc).abs() <= 1e-5 * scale).all()), KERNELS[k]
        y2 = torch.empty(n, device=cuda)
        assert sp.spmv_csr(A.ptr, x * 2, y2, sp.make_config(k), n).error_code == 0
        assert torch.equal(y2, y_k * 2)
    sp.ell_destroy(E)


def test_async_entry_points_and_launch_count(sp, orc, cuda):
    gen = gen_mod()
    rows = cols = 10000
    rp, ci, va = gen.random_csr(rows, cols, 8, 51, "cpu", 0.05)
    x = gen.vector_pm1(cols, 52, "cpu")
    A = GpuCSR(sp, rows, cols, rp.numpy(), ci.numpy(), va.numpy())
    y64, scale = orc.spmv_csr_f64(rows, A.rp, A.ci, A.va, x.numpy())
    d_x = x.to(cuda)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for k in (0, 1, 2):
            d_y = torch.full((rows,), float("nan"), device=cuda)
            before = sp.launch_count()
            assert sp.spmv_csr_async(A.mat, d_x, d_y, sp.make_config(k), stream.cuda_stream) == 0
            first = sp.launch_count() - before
            # merge-path: partition + tile + fix-up; row-owner kernels: one launch, plus a one-off scan of
            # the longest row the first time a row_ptrs array is seen (it picks the kernel variant)
            assert first == 3 if k == 2 else first in (1, 2)
            stream.synchronize()
            assert_within_tolerance(d_y.cpu().numpy(), y64, scale, f"async {KERNELS[k]}")
            before = sp.launch_count()
            assert sp.spmv_csr_async(A.mat, d_x, d_y, sp.make_config(k), stream.cuda_stream) == 0
            assert sp.launch_count() - before == (3 if k == 2 else 1)
            stream.synchronize()
    A.close()


def test_benchmark_harness_and_bandwidth(sp, cuda):
    """reference tests/test_benchmark.cu:17-61,106-149 and tests/test_bandwidth.cu:19-98"""
    gen = gen_mod()
    rows = cols = 50000
    rp, ci, va = gen.random_csr(rows, cols, 10, 61, "cpu")
    x = gen.vector_pm1(cols, 62, "cpu").numpy()
    A = GpuCSR(sp, rows, cols, rp.numpy(), ci.numpy(), va.numpy())
    for k in (0, 1, 2):
        r = sp.benchmark_csr(A.mat, x, sp.make_config(k), sp.make_bench_config(2, 7))
        assert r.name == b"CSR SpMV" and r.num_runs == 7 and r.execution_time_ms > 0
        assert r.min_time_ms <= r.avg_time_ms <= r.max_time_ms and r.stddev_time_ms >= 0
        assert r.gflops > 0 and r.bandwidth_gb_s > 0 and r.execution_time_ms == r.avg_time_ms
    gpu, cpu, speedup = sp.compare_gpu_cpu_csr(A.mat, x, sp.make_config(1), sp.make_bench_config(1, 3))
    assert cpu.name == b"CPU CSR SpMV" and cpu.num_runs == 3 and speedup > 0
    E = sp.ell_create(0, 0, 0)
    sp.ell_from_csr(E, A.mat)
    sp.ell_to_gpu(E)
    r = sp.benchmark_ell(E, x, sp.make_bench_config(1, 4))
    assert r.name == b"ELL SpMV" and r.num_runs == 4 and r.gflops > 0
    peak = sp.get_gpu_peak_bandwidth()
    assert 0 < peak < 10000
    m = sp.compute_bandwidth_csr(A.mat, 1.0)
    assert m.theoretical_bandwidth_gb_s == peak and 0 <= m.efficiency <= 1
    assert abs(m.achieved_bandwidth_gb_s - sp.csr_bytes(rows, cols, len(A.va)) / 1e9 / 1e-3) < 1e-3 * m.achieved_bandwidth_gb_s
    m = sp.compute_bandwidth_ell(E, 1.0)
    assert m.achieved_bandwidth_gb_s > 0
    sp.ell_destroy(E)
    A.close()


@pytest.mark.parametrize("env", [{"SPMV_B200_CSR_ROBUST": "0"}, {"SPMV_B200_CSR_ROBUST": "1"},
                                 {"SPMV_B200_CSR_NO_PIPE": "1"}, {"SPMV_B200_MERGE_TMA": "1"},
                                 {"SPMV_B200_ELL_VARIANT": "1"}, {"SPMV_B200_ELL_VARIANT": "3"}])
def test_every_kernel_variant_forced(cuda, env):
    """The launchers choose between kernel variants (fast / robust row-owner pipeline, register-staged
    fall-backs, TMA-staged merge tiles, one-shot / register-staged ELL).  Each is forced here through
    its tuning variable (read once per process, hence the subprocess) and must pass the same parity
    tests: random shapes, outlier rows, unaligned pointers, the Laplacian."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_spmv.py"), "-q", "-x", "-m", "gpu",
           "-k", "random_shapes or outlier_rows or unaligned or laplacian_medium or golden"]
    p = subprocess.run(cmd, env={**os.environ, **env}, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-3000:]


def test_products_in_place_kernel_is_bit_identical(sp, orc, cuda):
    """csr_prod_kernel (SCALAR / one-lane VECTOR on short-row matrices: TMA-staged window, products in
    place in stream order, then one thread per row front to back) against spmv_cpu_csr, bit for bit,
    on shapes that exercise the unstaged tail (nnz % 4 != 0), empty rows, rows == 1, a last window
    whose row_ptrs cannot be bulk-copied, and unaligned window starts."""
    gen = gen_mod()
    cases = [(1, 7, 3), (5, 5, 2), (257, 300, 3), (1000, 999, 5), (4099, 4099, 4), (70001, 70001, 6), (262144, 262144, 5)]
    for rows, cols, avg in cases:
        rp, ci, va = gen.random_csr(rows, cols, avg, seed=rows + avg, device="cpu")
        x = gen.vector_pm1(cols, 3, "cpu").numpy()
        A = GpuCSR(sp, rows, cols, rp.numpy(), ci.numpy(), va.numpy())
        y_cpu = orc.spmv_csr(rows, A.rp, A.ci, A.va, x)
        for kernel in (sp.SCALAR_CSR, sp.VECTOR_CSR):
            if kernel == sp.VECTOR_CSR and avg >= 6:
                continue  # several lanes per row: tolerance, covered elsewhere
            y, _ = run_csr(sp, A.mat, x, kernel, cuda, rows)
            assert np.array_equal(bits(y), bits(y_cpu)), (rows, cols, avg, kernel)
        A.close()
    # the Laplacian stencil (config 2 at 1024^2): the shape the kernel exists for
    grid = 1024
    n = grid * grid
    rp, ci, va = gen.laplacian_2d_csr(grid, "cpu")
    x = gen.vector_pm1(n, 42, "cpu").numpy()
    A = GpuCSR(sp, n, n, rp.numpy(), ci.numpy(), va.numpy())
    y, _ = run_csr(sp, A.mat, x, sp.SCALAR_CSR, cuda, n)
    assert np.array_equal(bits(y), bits(orc.spmv_csr(n, A.rp, A.ci, A.va, x)))
    A.close()


def test_pipelined_host_buffer_ell(sp, orc, cuda):
    """spmv_b200_spmv_ell_host: x and y in (pinned) HOST memory, upload / product / download pipelined
    over row chunks by the measured column range of every chunk.  Bit-identical to spmv_cpu_ell for a
    banded matrix (chunks run ahead of the upload), for a matrix whose rows read all of x (no run-ahead),
    for a shape the row-range kernel does not cover (rows % 4 != 0: single launch), and re-usable."""
    gen = gen_mod()

    def check(rows, cols, rp, ci, va, chunks, expect_ranged, expect_lookahead_small):
        x = gen.vector_pm1(cols, 11, "cpu").numpy()
        A = GpuCSR(sp, rows, cols, rp, ci, va)
        E = sp.ell_create(0, 0, 0)
        assert sp.ell_from_csr(E, A.mat) == 0 and sp.ell_to_gpu(E) == 0
        w = E.contents.max_nnz_per_row
        ec, ev = sp.ell_arrays(E)
        y_cpu = orc.spmv_ell(rows, w, ec, ev, x)
        plan = C.c_void_p()
        assert sp.lib.spmv_b200_ell_host_plan_create(E, chunks, C.byref(plan)) == 0
        n_chunks, ranged, look = C.c_int(), C.c_int(), C.c_int()
        assert sp.lib.spmv_b200_ell_host_plan_info(plan, C.byref(n_chunks), C.byref(ranged), C.byref(look)) == 0
        assert bool(ranged.value) == expect_ranged
        if expect_lookahead_small:
            assert look.value <= 1
        xh = torch.as_tensor(x).pin_memory()
        yh = torch.full((rows,), float("nan")).pin_memory()
        for rep in range(3):
            yh.fill_(float("nan"))
            assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
            assert np.array_equal(bits(yh.numpy()), bits(y_cpu)), (rows, cols, chunks, rep)
        # pageable buffers work too
        yp = np.full(rows, np.nan, np.float32)
        assert sp.lib.spmv_b200_spmv_ell_host(plan, x.ctypes.data, yp.ctypes.data) == 0
        assert np.array_equal(bits(yp), bits(y_cpu))
        sp.lib.spmv_b200_ell_host_plan_destroy(plan)
        sp.ell_destroy(E)
        A.close()

    grid = 512
    rp, ci, va = gen.laplacian_2d_csr(grid, "cpu")
    check(grid * grid, grid * grid, rp.numpy(), ci.numpy(), va.numpy(), 16, True, True)
    rp, ci, va = gen.random_csr(40000, 40000, 3, seed=5, device="cpu")
    check(40000, 40000, rp.numpy(), ci.numpy(), va.numpy(), 8, int(np.diff(rp.numpy()).max()) <= 8, False)
    rp, ci, va = gen.random_csr(30001, 50000, 2, seed=6, device="cpu")
    check(30001, 50000, rp.numpy(), ci.numpy(), va.numpy(), 4, False, False)


def test_gated_host_buffer_ell(sp, orc, cuda):
    """The gated form of spmv_b200_spmv_ell_host (one upload, one persistent kernel that consumes x while it
    arrives, downloads released by progress counters): taken by default for a matrix the TMA ring covers,
    bit-identical to spmv_cpu_ell over repeated calls (the sentinel refill between calls), with 1 .. many download
    chunks, with a row count that leaves a partial last window, and with an x that CONTAINS the sentinel bit
    pattern (accepted once the upload is complete: same bits as the device path spmv_ell)."""
    import os
    gen = gen_mod()
    for grid, chunks in ((512, 24), (510, 7), (256, 1), (1024, 200), (768, 0)):
        n = grid * grid
        rp, ci, va = gen.laplacian_2d_csr(grid, "cpu")
        A = GpuCSR(sp, n, n, rp.numpy(), ci.numpy(), va.numpy())
        E = sp.ell_create(0, 0, 0)
        assert sp.ell_from_csr(E, A.mat) == 0 and sp.ell_to_gpu(E) == 0
        w = E.contents.max_nnz_per_row
        ec, ev = sp.ell_arrays(E)
        if chunks:  # that many equal download chunks; 0: the default schedule (small first chunk, growing)
            os.environ["SPMV_B200_HOST_GATED_CHUNKS"] = str(chunks)
        plan = C.c_void_p()
        try:
            assert sp.lib.spmv_b200_ell_host_plan_create(E, 0, C.byref(plan)) == 0
        finally:
            os.environ.pop("SPMV_B200_HOST_GATED_CHUNKS", None)
        gated, down = C.c_int(), C.c_int()
        assert sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(gated), C.byref(down)) == 0
        assert gated.value == 1 and 1 <= down.value <= (chunks or 255)
        yh = torch.empty(n).pin_memory()
        for seed in (1, 2, 3):
            x = gen.vector_pm1(n, seed, "cpu").numpy()
            xh = torch.as_tensor(x).pin_memory()
            yh.fill_(float("nan"))
            assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
            assert np.array_equal(bits(yh.numpy()), bits(orc.spmv_ell(n, w, ec, ev, x))), (grid, chunks, seed)
        # x holding the sentinel pattern in places (and pageable buffers): same bits as the device path
        x = gen.vector_pm1(n, 9, "cpu").numpy()
        x.view(np.uint32)[::1001] = 0x7FA3C0DE
        x.view(np.uint32)[n - 1] = 0x7FA3C0DE
        d_x = torch.as_tensor(x).to(cuda)
        d_y = torch.empty(n, dtype=torch.float32, device=cuda)
        assert sp.spmv_ell(E, d_x, d_y, None, n).error_code == 0
        yp = np.full(n, 7.0, np.float32)
        assert sp.lib.spmv_b200_spmv_ell_host(plan, x.ctypes.data, yp.ctypes.data) == 0
        assert np.array_equal(bits(yp), bits(d_y.cpu().numpy()))
        assert sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(gated), None) == 0 and gated.value == 1
        sp.lib.spmv_b200_ell_host_plan_destroy(plan)
        sp.ell_destroy(E)
        A.close()


def test_gated_host_call_times_out_into_the_chunked_form(cuda):
    """The safety net of the gated host-buffer call: if x never arrives (here: the upload is suppressed by a test switch)
    the kernel's producers time out, everybody drains, and the SAME call repeats itself in the chunked form -- correct y,
    no hang, and the plan stays chunked afterwards.  Child process: the switches are read once."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import ctypes as C, sys, time, torch
sys.path.insert(0, %r)
from _load_pkg import load_pkg
sp = load_pkg()
import gpu_spmv_b200.gen as gen
dev = torch.device("cuda:0")
grid = 512
n = grid * grid
rp, ci, va = gen.laplacian_2d_csr(grid, dev)
x = gen.vector_pm1(n, 7, dev)
A = sp.DeviceCSR(n, n, rp, ci, va)
E = sp.ell_create(0, 0, 0)
assert sp.ell_from_csr_device(E, A.ptr) == 0
y = torch.empty(n, device=dev)
assert sp.spmv_ell(E, x, y, None, n).error_code == 0
plan = C.c_void_p()
assert sp.lib.spmv_b200_ell_host_plan_create(E, 0, C.byref(plan)) == 0
g = C.c_int()
sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(g), None)
assert g.value == 1
xh, yh = x.cpu().pin_memory(), torch.full((n,), float("nan")).pin_memory()
t0 = time.perf_counter()
assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
dt = time.perf_counter() - t0
assert torch.equal(yh.view(torch.int32), y.cpu().view(torch.int32)), "wrong y after the fall-back"
sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(g), None)
assert g.value == 0, "the plan should stay in the chunked form"
yh.fill_(float("nan"))
assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
assert torch.equal(yh.view(torch.int32), y.cpu().view(torch.int32))
print("fell back after %%.0f ms" %% (dt * 1e3))
""" % root
    env = dict(os.environ, SPMV_B200_HOST_GATED_TEST_STALL="1", SPMV_B200_HOST_GATED_TIMEOUT_MS="40")
    p = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and "fell back after" in p.stdout, p.stdout[-3000:]
    print(p.stdout.strip().splitlines()[-1])


def test_h2d_order_probe_and_l2_persistence_hooks(sp, orc, cuda):
    """The two design probes behind DESIGN 6 / 8 stay runnable: (i) spmv_b200_probe_h2d_order -- one host-to-device copy
    lands front to back (the premise of the gated host-buffer call; the gated kernel itself never relies on it);
    (ii) spmv_b200_set_l2_persistence -- a persisting L2 set-aside + access-policy window over x changes no bit of a product."""
    gen = gen_mod()
    n, samples = 1 << 22, 16
    xh = torch.rand(n).pin_memory()
    out = (C.c_longlong * (samples + 1))()
    for _ in range(2):  # the first copy from freshly pinned pages is slow
        assert sp.lib.spmv_b200_probe_h2d_order(xh.data_ptr(), n, samples, out, 1, 0) == 0
    arrivals = list(out)[:samples]
    assert min(arrivals) == 0 and out[samples] > 0
    print("arrival of every 1/16 of one H2D copy [ns]:", arrivals)  # front to back on B200 (profiles/r2_host_gated.txt); a premise of
    assert arrivals[-1] >= arrivals[0]                                   # the gated call's SPEED only, so not asserted pair by pair
    assert sp.lib.spmv_b200_probe_h2d_order(None, n, samples, out, 1, 0) != 0

    lim = [C.c_ulonglong() for _ in range(3)]
    assert sp.lib.spmv_b200_l2_persistence_limits(*[C.byref(v) for v in lim]) == 0
    max_aside, max_window, l2 = [v.value for v in lim]
    assert l2 > 0 and max_window > 0
    rows = 60000
    rp, ci, va = gen.random_csr(rows, rows, 7, seed=12, device="cpu")
    x = gen.vector_pm1(rows, 3, "cpu").numpy()
    A = GpuCSR(sp, rows, rows, rp.numpy(), ci.numpy(), va.numpy())
    d_x = torch.as_tensor(x).to(cuda)
    d_y = torch.empty(rows, dtype=torch.float32, device=cuda)
    stream = torch.cuda.Stream()
    cfg = sp.make_config(sp.MERGE_PATH)
    ys = []
    for window in (False, True, False):
        if window:
            assert sp.lib.spmv_b200_set_l2_persistence(C.c_void_p(stream.cuda_stream), C.c_void_p(d_x.data_ptr()), min(rows * 4, max_window),
                                                       1.0, min(max_aside, 8 << 20)) == 0
        else:
            assert sp.lib.spmv_b200_set_l2_persistence(C.c_void_p(stream.cuda_stream), None, 0, 0.0, 0) == 0
        d_y.fill_(float("nan"))
        torch.cuda.synchronize()
        assert sp.lib.spmv_b200_spmv_csr_async(A.mat, sp.dptr(d_x), sp.dptr(d_y), C.byref(cfg), C.c_void_p(stream.cuda_stream)) == 0
        stream.synchronize()
        ys.append(d_y.cpu().numpy().copy())
    assert np.array_equal(bits(ys[0]), bits(ys[1])) and np.array_equal(bits(ys[0]), bits(ys[2]))
    y64, scale = orc.spmv_csr_f64(rows, A.rp, A.ci, A.va, x)
    assert np.all(np.abs(ys[1].astype(np.float64) - y64) <= 1e-5 * scale + 1e-30)
    A.close()


def test_benchmark_csr_report_has_roofline_fields(sp, cuda):
    """spmv_b200_benchmark_csr_report: the reference's nine benchmark keys (src/benchmark.cu:187-202) in the
    reference's format, followed by the roofline figures (algorithmic bytes of src/bandwidth.cpp:34-42,
    effective GB/s, fraction of the peak) and the single-threaded CPU time next to them."""
    import json
    gen = gen_mod()
    rp, ci, va = gen.random_csr(30000, 30000, 12, seed=21, device="cpu")
    x = gen.vector_pm1(30000, 3, "cpu").numpy()
    A = GpuCSR(sp, 30000, 30000, rp.numpy(), ci.numpy(), va.numpy())
    buf = C.create_string_buffer(4096)
    bc = sp.make_bench_config(2, 5, True)
    n = sp.lib.spmv_b200_benchmark_csr_report(A.mat, x.ctypes.data_as(C.POINTER(C.c_float)), None, C.byref(bc), 6548.5, buf, 4096)
    assert n > 0, n
    rep = json.loads(buf.value.decode())
    keys = list(rep)
    assert keys[:9] == ["name", "execution_time_ms", "gflops", "bandwidth_gb_s", "avg_time_ms", "min_time_ms", "max_time_ms",
                        "stddev_time_ms", "num_runs"]
    assert rep["num_runs"] == 5 and rep["kernel"] in ("SCALAR_CSR", "VECTOR_CSR", "MERGE_PATH")
    assert rep["algorithmic_bytes"] == sp.csr_bytes(30000, 30000, ci.numel())
    assert rep["peak_gb_s"] == 6548.5 and 0.0 < rep["roofline_fraction"] < 1.5
    assert abs(rep["effective_gb_s"] - rep["algorithmic_bytes"] / 1e9 / (rep["avg_time_ms"] * 1e-3)) <= 1e-3 * rep["effective_gb_s"]
    assert rep["cpu_threads"] == 1 and rep["cpu_avg_time_ms"] > 0 and rep["speedup"] > 0
    # the reference's reader still parses the object (first `"key":` wins, src/benchmark.cu:215-237)
    back = sp.benchmark_from_json(buf.value.decode())
    assert back.num_runs == 5
    assert sp.lib.spmv_b200_benchmark_csr_report(A.mat, x.ctypes.data_as(C.POINTER(C.c_float)), None, C.byref(bc), 0.0, buf, 16) == -1
    A.close()
