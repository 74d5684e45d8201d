"""The C-ABI library loads, exports every symbol include/spmv_b200.h declares
and the 39 C++ symbols of the reference API; the product
never routes through the oracle; device entry points fail loudly without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpu-spmv_b200", "lib", "libspmv_b200.so")

APPENDIX_A = """
_ZN4spmv10csr_createEiii _ZN4spmv10ell_createEiii _ZN4spmv11csr_destroyEPNS_9CSRMatrixE
_ZN4spmv11ell_destroyEPNS_9ELLMatrixE _ZN4spmv14csr_from_denseEPNS_9CSRMatrixEPKfii
_ZN4spmv14ell_from_denseEPNS_9ELLMatrixEPKfii _ZN4spmv12csr_to_denseEPKNS_9CSRMatrixEPf
_ZN4spmv12ell_to_denseEPKNS_9ELLMatrixEPf _ZN4spmv15csr_get_elementEPKNS_9CSRMatrixEii
_ZN4spmv15ell_get_elementEPKNS_9ELLMatrixEii _ZN4spmv10csr_to_gpuEPNS_9CSRMatrixE
_ZN4spmv10ell_to_gpuEPNS_9ELLMatrixE _ZN4spmv12csr_from_gpuEPNS_9CSRMatrixE
_ZN4spmv12ell_from_gpuEPNS_9ELLMatrixE _ZN4spmv12csr_free_gpuEPNS_9CSRMatrixE
_ZN4spmv12ell_free_gpuEPNS_9ELLMatrixE _ZN4spmv13csr_serializeEPKNS_9CSRMatrixEPKc
_ZN4spmv13ell_serializeEPKNS_9ELLMatrixEPKc _ZN4spmv15csr_deserializeEPNS_9CSRMatrixEPKc
_ZN4spmv15ell_deserializeEPNS_9ELLMatrixEPKc _ZN4spmv17csr_compute_statsEPKNS_9CSRMatrixE
_ZN4spmv12ell_from_csrEPNS_9ELLMatrixEPKNS_9CSRMatrixE _ZN4spmv12spmv_cpu_csrEPKNS_9CSRMatrixEPKfPf
_ZN4spmv12spmv_cpu_ellEPKNS_9ELLMatrixEPKfPf _ZN4spmv16spmv_auto_configEPKNS_9CSRMatrixE
_ZN4spmv8spmv_csrEPKNS_9CSRMatrixEPKfPfPKNS_10SpMVConfigEi
_ZN4spmv8spmv_ellEPKNS_9ELLMatrixEPKfPfPKNS_10SpMVConfigEi
_ZN4spmv21compute_bandwidth_csrEPKNS_9CSRMatrixEf _ZN4spmv21compute_bandwidth_ellEPKNS_9ELLMatrixEf
_ZN4spmv22get_gpu_peak_bandwidthEv _ZN4spmv8pagerankEPKNS_9CSRMatrixEPKNS_14PageRankConfigE
_ZN4spmv13pagerank_freeEPNS_14PageRankResultE
_ZN4spmv14pagerank_top_kEPKNS_14PageRankResultEiiPNS_8TopKNodeE
_ZN4spmv13benchmark_csrEPKNS_9CSRMatrixEPKfPKNS_10SpMVConfigEPKNS_15BenchmarkConfigE
_ZN4spmv13benchmark_ellEPKNS_9ELLMatrixEPKfPKNS_15BenchmarkConfigE
_ZN4spmv19compare_gpu_cpu_csrEPKNS_9CSRMatrixEPKfPKNS_10SpMVConfigEPKNS_15BenchmarkConfigE
_ZN4spmv17benchmark_to_jsonB5cxx11ERKNS_15BenchmarkResultE
_ZN4spmv18comparison_to_jsonB5cxx11ERKNS_16ComparisonResultE
_ZN4spmv19benchmark_from_jsonERKNSt7__cxx1112basic_stringIcSt11char_traitsIcESaIcEEE
""".split()


def exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_every_declared_symbol_is_exported(sp):
    header = open(os.path.join(ROOT, "include", "spmv_b200.h")).read()
    declared = set(re.findall(r"SPMV_B200_API[^;(]*?\b(spmv_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 55
    syms = exported()
    assert not (declared - syms), sorted(declared - syms)
    from gpu_spmv_b200 import capi
    assert declared == set(capi.PROTOTYPES), declared ^ set(capi.PROTOTYPES)


def test_reference_cxx_symbols_exported(sp):
    assert len(APPENDIX_A) == 39  # SURVEY Appendix A lists 39 mangled names
    syms = exported()
    assert not [s for s in APPENDIX_A if s not in syms]


def test_struct_sizes_match_reference_layout(sp):
    import ctypes as C
    sizes = {sp.CSRMatrix: 72, sp.ELLMatrix: 56, sp.SpMVConfig: 12, sp.SpMVResult: 24, sp.CSRStats: 16,
             sp.PageRankConfig: 12, sp.PageRankResult: 24, sp.TopKNode: 8, sp.BandwidthMetrics: 12,
             sp.BenchmarkConfig: 12}
    for t, n in sizes.items():
        assert C.sizeof(t) == n, t
    assert sp.CSRMatrix.values.offset == 16 and sp.CSRMatrix.d_values.offset == 40
    assert sp.CSRMatrix.owns_host_memory.offset == 64 and sp.CSRMatrix.owns_device_memory.offset == 65
    assert sp.ELLMatrix.d_values.offset == 32 and sp.ELLMatrix.owns_host_memory.offset == 48
    assert sp.SpMVResult.error_code.offset == 20 and sp.PageRankResult.converged.offset == 16


def test_product_never_touches_the_oracle():
    """No file of the product (package, include/) names oracle/ code."""
    bad = []
    for base in ("gpu-spmv_b200", "include"):
        for dirpath, _dirs, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dirpath or os.sep + "lib" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"liboracle|oracle_binding|spmv_oracle|libspmv_ref|orc_[a-z_]+\(|ref_shim", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    deps = subprocess.check_output(["ldd", LIB], text=True)
    assert "oracle" not in deps and "spmv_ref" not in deps


def test_argument_validation_order(sp):
    # reference src/spmv_kernels.cu:219-232: null -> -8, dimension -> -1, missing device arrays -> -5
    A = sp.csr_create(3, 4, 0)
    assert sp.spmv_csr(None, 1, 1, None, -1).error_code == -8
    assert sp.spmv_csr(A, None, 1, None, -1).error_code == -8
    assert sp.spmv_csr(A, 1, None, None, -1).error_code == -8
    assert sp.spmv_csr(A, 1, 1, None, 3).error_code == -1
    assert sp.spmv_csr(A, 1, 1, None, 4).error_code == -5   # not on the device
    assert sp.spmv_csr(A, 1, 1, None, -1).error_code == -5  # vec_size < 0 skips the dimension check
    sp.csr_destroy(A)
    E = sp.ell_create(3, 4, 2)
    assert sp.spmv_ell(None, 1, 1, None, -1).error_code == -8
    assert sp.spmv_ell(E, 1, 1, None, 5).error_code == -1
    assert sp.spmv_ell(E, 1, 1, None, 4).error_code == -5
    sp.ell_destroy(E)
    assert sp.spmv_validate_dimensions(4, 4) and not sp.spmv_validate_dimensions(4, 5)


def test_device_entry_points_fail_loudly_without_gpu(sp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, np.eye(4, dtype=np.float32), 4, 4)
    assert sp.csr_to_gpu(A) != 0  # CUDA failure is reported, nothing is emulated on the host
    assert not A.contents.d_values
    A.contents.d_row_ptrs = 16
    A.contents.d_col_indices = 16
    A.contents.d_values = 16  # pretend device arrays: the call must still fail, not compute on the CPU
    assert sp.spmv_csr(A, 16, 16, None, 4).error_code != 0
    A.contents.d_row_ptrs = None
    A.contents.d_col_indices = None
    A.contents.d_values = None
    sp.csr_destroy(A)


def test_extension_entry_points_validate_their_arguments(sp):
    """The additive entry points (plans, assembly, top-k, multicast step) reject bad arguments with the
    reference's codes before touching the device (so this runs without a GPU)."""
    import ctypes as C
    L = sp.lib
    bad_arg, bad_fmt = int(sp.SpMVError.INVALID_ARGUMENT), int(sp.SpMVError.INVALID_FORMAT)
    handle = C.c_void_p()
    assert L.spmv_b200_csr_plan_create(None, 0, 0, C.byref(handle)) == bad_arg
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, np.eye(4, dtype=np.float32), 4, 4)
    assert L.spmv_b200_csr_plan_create(A, 0, 0, None) == bad_arg
    assert L.spmv_b200_csr_plan_create(A, 0, 0, C.byref(handle)) == bad_fmt  # host arrays only: not on the device
    assert not handle.value
    assert L.spmv_b200_spmv_csr_planned(None, None, None, None) == bad_arg
    assert L.spmv_b200_csr_plan_info(None, None, None, None) == bad_arg
    L.spmv_b200_csr_plan_destroy(None)
    L.spmv_b200_csr_forget_plan(None)
    assert sp.csr_auto_plan_info(A) == (0, 0)
    assert L.spmv_b200_csr_from_coo_device(None, 1, 1, 0, None, None, None) == bad_arg
    assert L.spmv_b200_csr_from_coo_device(A, -1, 1, 0, None, None, None) == bad_arg
    assert L.spmv_b200_csr_from_coo_device(A, 1, 1, 5, None, None, None) == bad_arg
    assert (A.contents.num_rows, A.contents.nnz) == (4, 4)  # untouched
    assert L.spmv_b200_csr_normalize_columns_device(None) == bad_arg
    assert L.spmv_b200_csr_normalize_columns_device(A) == bad_fmt
    node = (sp.TopKNode * 1)()
    assert L.spmv_b200_pagerank_top_k_device(None, 4, 1, node) == bad_arg
    assert L.spmv_b200_pagerank_top_k_device(16, -1, 1, node) == bad_arg
    assert L.spmv_b200_pagerank_top_k_device(16, 4, 0, node) == 0  # nothing to do
    assert L.spmv_b200_pr_step_multicast(None, None, None, 0.85, None, None, None, None, 2, 0, None) == bad_arg
    assert L.spmv_b200_pr_plan_set_hot(None, 0, 0, None) == bad_arg
    assert L.spmv_b200_pagerank_device_history(A, None, None, None, None, None, None, -1) == bad_arg
    # host-buffer call and its probes
    plan = C.c_void_p()
    assert L.spmv_b200_ell_host_plan_create(None, 0, C.byref(plan)) == bad_arg and not plan.value
    E = sp.ell_create(0, 0, 0)
    assert L.spmv_b200_ell_host_plan_create(E, 0, C.byref(plan)) == bad_fmt  # not on the device
    sp.ell_destroy(E)
    assert L.spmv_b200_spmv_ell_host(None, None, None) == bad_arg
    assert L.spmv_b200_ell_host_plan_gated(None, None, None) == bad_arg
    assert L.spmv_b200_ell_host_plan_info(None, None, None, None) == bad_arg
    L.spmv_b200_ell_host_plan_destroy(None)
    out = (C.c_longlong * 5)()
    assert L.spmv_b200_probe_h2d_order(None, 1024, 4, out, 0, 0) == bad_arg
    assert L.spmv_b200_probe_h2d_order(out, 1024, 0, out, 0, 0) == bad_arg
    assert L.spmv_b200_probe_h2d_order(out, 2, 4, out, 0, 0) == bad_arg
    sp.csr_destroy(A)
