"""The BASELINE configurations at their FULL sizes against the CPU oracle (not against the GPU
library itself): config 2 (16.7 M rows) bit for bit against the reference's own spmv_cpu_ell /
spmv_cpu_csr, config 3 (50 M rows with the real 1 M-nnz outlier rows) and config 4 (R-MAT 24,
268 M non-zeros) against the f64-accumulating row dot within north_star's tolerance
|y - y_ref| <= 1e-5 * sum_j |a_ij x_j| per row, config 5's recurrence (R-MAT 22, 5 fixed iterations)
against the f64-accumulator restatement within L1 1e-6.  The oracle passes take 0.1 - 3 s of CPU each.
Reference: src/spmv_cpu.cpp:6-32, src/pagerank.cu:93-150."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_helpers import assert_within_tolerance, bits

pytestmark = pytest.mark.gpu


def gen_mod():
    import gpu_spmv_b200.gen as gen
    return gen


def test_config2_full_size_bit_identical_to_reference_cpu(sp, orc, cuda):
    """ELL kernel, SCALAR_CSR and VECTOR_CSR (the selector's choice: the short-row ring) on the
    4096 x 4096 Laplacian == spmv_cpu_ell / spmv_cpu_csr of the unmodified reference (oracle/_ref when it
    is built, else the oracle port -- the two are pinned to each other in test_oracle_pinned.py)."""
    from oracle_binding import Ref
    gen = gen_mod()
    dev = torch.device("cuda:0")
    n = 4096 * 4096
    rp, ci, va = gen.laplacian_2d_csr(4096, dev)
    x = gen.vector_pm1(n, 42, dev)
    rp_n, ci_n, va_n, x_n = rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy(), x.cpu().numpy()
    if Ref.available():
        ref = Ref()
        h, keep = ref.csr_wrap(n, n, rp_n, ci_n, va_n)
        he, st = ref.ell_from_csr(h)
        assert st == 0
        y_cpu = np.empty(n, np.float32)
        fp = C.POINTER(C.c_float)
        ref.L.ref_time_spmv_cpu_ell(he, x_n.ctypes.data_as(fp), y_cpu.ctypes.data_as(fp), 1)
        ref.L.ref_ell_destroy(he)
        ref.L.ref_csr_destroy(h)
    else:
        w, ec, ev = orc.ell_from_csr(n, rp_n, ci_n, va_n)
        y_cpu = orc.spmv_ell(n, w, ec, ev, x_n)
    assert np.array_equal(bits(orc.spmv_csr(n, rp_n, ci_n, va_n, x_n)), bits(y_cpu))  # CSR and ELL agree on the CPU
    A = sp.DeviceCSR(n, n, rp, ci, va)
    y = torch.empty(n, device=dev)
    for kernel in (sp.SCALAR_CSR, sp.VECTOR_CSR):
        y.fill_(float("nan"))
        assert sp.spmv_csr(A.ptr, x, y, sp.make_config(kernel), n).error_code == 0
        assert np.array_equal(bits(y.cpu().numpy()), bits(y_cpu)), sp.KERNEL_NAMES[kernel]
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.ptr) == 0
    y.fill_(float("nan"))
    assert sp.spmv_ell(E, x, y, None, n).error_code == 0
    assert np.array_equal(bits(y.cpu().numpy()), bits(y_cpu))
    y.fill_(float("nan"))
    assert sp.spmv_csr(A.ptr, x, y, sp.make_config(sp.MERGE_PATH), n).error_code == 0
    y64, scale = orc.spmv_csr_f64(n, rp_n, ci_n, va_n, x_n)
    assert_within_tolerance(y.cpu().numpy(), y64, scale, "config 2 through MERGE_PATH")
    sp.ell_destroy(E)


def test_config3_full_size_with_1m_nnz_rows(sp, orc, cuda):
    """50 M rows, avg 3 nnz per row, rows 0 / 12.5 M / 25 M / 37.5 M with 1 000 000 nnz each:
    MERGE_PATH (the selector's choice under the outlier override) within tolerance of the f64 row dot;
    the LITERAL fp32 reference (sequential fp32 sum, src/spmv_cpu.cpp:6-16) is also measured against it --
    on the 1 M-nnz rows it is the less accurate of the two."""
    gen = gen_mod()
    dev = torch.device("cuda:0")
    rows = 50_000_000
    rp, ci, va = gen.short_rows_with_outliers_csr(rows, 43, dev)
    x = gen.uniform_01_open_low(5, torch.arange(rows, device=dev), 9)
    rp_n, ci_n, va_n, x_n = rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy(), x.cpu().numpy()
    lengths = np.diff(rp_n)
    outliers = [0, rows // 4, rows // 2, 3 * (rows // 4)]
    assert all(lengths[r] == 1_000_000 for r in outliers)
    A = sp.DeviceCSR(rows, rows, rp, ci, va)
    y = torch.full((rows,), float("nan"), device=dev)
    assert sp.spmv_csr(A.ptr, x, y, sp.make_config(sp.MERGE_PATH), rows).error_code == 0
    y_gpu = y.cpu().numpy()
    y64, scale = orc.spmv_csr_f64(rows, rp_n, ci_n, va_n, x_n)
    assert_within_tolerance(y_gpu, y64, scale, "config 3 through MERGE_PATH")
    y_lit = orc.spmv_csr(rows, rp_n, ci_n, va_n, x_n)
    err_gpu = np.abs(y_gpu[outliers].astype(np.float64) - y64[outliers]) / scale[outliers]
    err_lit = np.abs(y_lit[outliers].astype(np.float64) - y64[outliers]) / scale[outliers]
    print(f"1 M-nnz rows, relative error: MERGE_PATH {err_gpu.max():.2e}, literal fp32 reference {err_lit.max():.2e}")
    assert err_gpu.max() <= 1e-5 and err_gpu.max() <= err_lit.max() + 1e-7
    assert bool((y_gpu[lengths == 0] == 0).all())


def test_config4_full_size_against_f64_oracle(sp, orc, cuda):
    """R-MAT scale 24, edge factor 16 (268 435 456 non-zeros): MERGE_PATH and the hub-column plan against
    the f64 row dot computed on the CPU, row by row within 1e-5 * sum |a_ij x_j|."""
    gen = gen_mod()
    dev = torch.device("cuda:0")
    n, rp, ci, va = gen.rmat_pagerank_csr(24, 16, 44, dev)
    assert ci.numel() == 268435456
    x = gen.vector_pm1(n, 11, dev)
    rp_n, ci_n, va_n, x_n = rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy(), x.cpu().numpy()
    y64, scale = orc.spmv_csr_f64(n, rp_n, ci_n, va_n, x_n)
    A = sp.DeviceCSR(n, n, rp, ci, va)
    y = torch.full((n,), float("nan"), device=dev)
    assert sp.spmv_csr(A.ptr, x, y, sp.make_config(sp.MERGE_PATH), n).error_code == 0
    y_merge = y.cpu().numpy()
    assert_within_tolerance(y_merge, y64, scale, "config 4 through MERGE_PATH")
    plan = sp.CsrPlan(A.ptr)
    y.fill_(float("nan"))
    assert plan.spmv(x, y) == 0
    torch.cuda.synchronize()
    assert_within_tolerance(y.cpu().numpy(), y64, scale, "config 4 through a CSR plan")
    plan.close()


def test_config5_recurrence_scale22_against_f64_oracle(sp, orc, cuda):
    """PageRank on R-MAT scale 22 (67 M edges), 5 fixed iterations: the device loop (one GPU) and the
    sharded loop (4 ranks sharing the GPU, peer-store exchange) against the f64-accumulator restatement of
    the reference recurrence (F7): L1 <= 1e-6, top-100 identical up to ties."""
    import gpu_spmv_b200.dist as D
    gen = gen_mod()
    dev = torch.device("cuda:0")
    n, rp, ci, va = gen.rmat_pagerank_csr(22, 16, 45, "cpu")
    rp_n, ci_n, va_n = rp.numpy(), ci.numpy(), va.numpy()
    iters = 5
    o_ranks = orc.pagerank_f64(n, n, rp_n, ci_n, va_n, 0.85, 1e-6, 100, fixed_it=iters)[0]
    A = sp.csr_from_arrays(n, n, rp_n, ci_n, va_n)
    cfg = sp.make_pagerank_config(0.85, 0.0, iters)
    res = D.PrDistResult()
    for world in (1, 4):
        ranks = np.empty(n, np.float32)
        devices = (C.c_int * world)(*([0] * world))
        rc = sp.lib.spmv_b200_pagerank_multi(A, C.byref(cfg), world, devices, D.EXCHANGE_P2P, 2, iters,
                                             ranks.ctypes.data_as(C.c_void_p), C.byref(res))
        assert rc == 0 and res.iterations == iters
        l1 = float(np.abs(ranks.astype(np.float64) - o_ranks.astype(np.float64)).sum())
        assert l1 <= 1e-6, f"{world} rank(s): L1 {l1:.3e}"
        k = 100
        top_gpu, top_orc = np.argsort(-ranks, kind="stable")[:k], np.argsort(-o_ranks, kind="stable")[:k]
        cut = min(ranks[top_gpu[-1]], o_ranks[top_orc[-1]])
        strictly_above = set(np.nonzero(o_ranks > cut * (1 + 1e-5))[0].tolist())
        assert strictly_above <= set(top_gpu.tolist()), "top-k differs beyond ties"
    sp.csr_destroy(A)
