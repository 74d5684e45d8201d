"""The SHARDED PageRank path on a box with a single GPU: spmv_b200_pagerank_multi with an explicit
device list that names device 0 for every rank.  The ranks are real -- one host thread, one stream,
one row shard, one symmetric rank-vector pair and one control block each -- and they exchange their
slices with peer stores and meet in the in-kernel flag barrier exactly as on 2-8 GPUs; only the
"peer" mappings point into the same device.  This is what lets the round-end GPU run (one B200)
check the sharded path against the oracle; tests/test_gpu_multi.py covers real multi-GPU boxes
(multicast and NCCL transports need distinct devices).

Checks: the f64-accumulator restatement of the reference recurrence (reference src/pagerank.cu:93-150)
at equal iteration count, L1 <= 1e-6 (north_star); the single-GPU device loop; the stop rule."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _multi(sp, D, A, n, world, iters=0, tol=1e-6, max_it=100, exchange=1, weight=4):
    cfg = sp.make_pagerank_config(0.85, tol, max_it)
    ranks = np.empty(n, np.float32)
    res = D.PrDistResult()
    devices = (C.c_int * world)(*([0] * world))
    rc = sp.lib.spmv_b200_pagerank_multi(A, C.byref(cfg), world, devices, exchange, weight, iters,
                                         ranks.ctypes.data_as(C.c_void_p), C.byref(res))
    assert rc == 0, sp.spmv_error_string(rc)
    return ranks, res


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("scale", [12, 16])
def test_ranks_sharing_one_gpu_match_oracle(sp, orc, cuda, world, scale):
    import gpu_spmv_b200.dist as D
    import gpu_spmv_b200.gen as gen
    n, rp, ci, va = gen.rmat_pagerank_csr(scale, 16, 11, "cpu")
    rp_n, ci_n, va_n = rp.numpy(), ci.numpy(), va.numpy()
    A = sp.csr_from_arrays(n, n, rp_n, ci_n, va_n)
    iters = 8
    ranks, res = _multi(sp, D, A, n, world, iters=iters)
    assert res.exchange == D.EXCHANGE_P2P and res.iterations == iters
    o_ranks, _, o_l2, _, _ = orc.pagerank_f64(n, n, rp_n, ci_n, va_n, 0.85, 1e-6, 100, fixed_it=iters)
    l1 = float(np.abs(ranks.astype(np.float64) - o_ranks).sum())
    assert l1 <= 1e-6, f"world {world}: L1 distance to the f64 restatement {l1:.3e}"
    # the residual of a nearly converged vector is fp32 rounding noise (1e-9): compare above that floor only
    assert abs(res.final_residual - o_l2) <= 1e-2 * o_l2 + 1e-8
    assert abs(float(ranks.astype(np.float64).sum()) - 1.0) <= 1e-6
    # the shard count must not change a bit pattern that depends only on per-row sums ... it does change
    # the merge tiles, so across shard counts only the tolerance holds; the SAME shard count is deterministic
    again, _ = _multi(sp, D, A, n, world, iters=iters)
    assert np.array_equal(ranks.view(np.uint32), again.view(np.uint32)), "sharded PageRank is not deterministic"
    sp.csr_destroy(A)


def test_stop_rule_and_single_gpu_agreement(sp, orc, cuda):
    """Stop rule (L2 of the delta < tol, reference src/pagerank.cu:118-127) read one iteration late: the
    sharded loop must stop at the same iteration as the single-GPU device loop and as the oracle."""
    import gpu_spmv_b200.dist as D
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    n, rp, ci, va = gen.rmat_pagerank_csr(14, 16, 5, "cpu")
    rp_n, ci_n, va_n = rp.numpy(), ci.numpy(), va.numpy()
    A = sp.csr_from_arrays(n, n, rp_n, ci_n, va_n)
    assert sp.csr_to_gpu(A) == 0
    d_ranks = torch.empty(n, dtype=torch.float32, device=dev)
    rc, it1, res1, conv1, _ = sp.pagerank_device(A, d_ranks, sp.make_pagerank_config(0.85, 1e-6, 100))
    assert rc == 0 and conv1
    ranks, res = _multi(sp, D, A, n, 3)
    assert res.converged == 1 and res.iterations == it1
    o_ranks, o_it, _, _, o_conv = orc.pagerank_f64(n, n, rp_n, ci_n, va_n, 0.85, 1e-6, 100)
    assert o_conv and o_it == res.iterations
    assert float(np.abs(ranks.astype(np.float64) - o_ranks).sum()) <= 1e-6
    assert float(np.abs(ranks.astype(np.float64) - d_ranks.cpu().numpy().astype(np.float64)).sum()) <= 1e-6
    sp.csr_destroy(A)


def test_empty_shards_and_dangling_nodes(sp, orc, cuda):
    """More ranks than non-empty rows on one side of the graph, and a graph whose second half is all
    dangling nodes (no out-links): shards with zero non-zeros must still take part in every barrier."""
    import gpu_spmv_b200.dist as D
    n = 4096
    rng = np.random.default_rng(3)
    src = rng.integers(0, n // 2, size=20000)          # only the first half has out-links
    dst = rng.integers(0, n // 8, size=20000)          # only the first eighth has in-links
    outdeg = np.bincount(src, minlength=n)
    order = np.lexsort((src, dst))
    src, dst = src[order], dst[order]
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum(np.bincount(dst, minlength=n))
    va = (1.0 / outdeg[src]).astype(np.float32)
    A = sp.csr_from_arrays(n, n, rp, src.astype(np.int32), va)
    ranks, res = _multi(sp, D, A, n, 4, iters=8, weight=0)   # weight 0: pure nnz balance -> trailing shards are empty
    o_ranks = orc.pagerank_f64(n, n, rp, src.astype(np.int32), va, 0.85, 1e-6, 100, fixed_it=8)[0]
    assert float(np.abs(ranks.astype(np.float64) - o_ranks).sum()) <= 1e-6
    sp.csr_destroy(A)


def test_rejects_bad_device_lists(sp, cuda):
    import gpu_spmv_b200.dist as D
    n = 64
    rp = np.arange(n + 1, dtype=np.int32)
    A = sp.csr_from_arrays(n, n, rp, np.arange(n, dtype=np.int32), np.ones(n, np.float32))
    cfg = sp.make_pagerank_config(0.85, 1e-6, 10)
    out = np.empty(n, np.float32)
    res = D.PrDistResult()
    bad = (C.c_int * 2)(0, 99)
    rc = sp.lib.spmv_b200_pagerank_multi(A, C.byref(cfg), 2, bad, 1, 4, 0, out.ctypes.data_as(C.c_void_p), C.byref(res))
    assert rc == int(sp.SpMVError.INVALID_ARGUMENT)
    rc = sp.lib.spmv_b200_pagerank_multi(A, C.byref(cfg), 9, None, 1, 4, 0, out.ctypes.data_as(C.c_void_p), C.byref(res))
    assert rc == int(sp.SpMVError.INVALID_ARGUMENT)
    sp.csr_destroy(A)


def test_plain_c_caller_single_gpu(cuda):
    """tests/c/pagerank_multi_test.c (a C program linked against libspmv_b200.so through include/spmv_b200.h,
    no Python in the loop) with n_gpus = 1: spmv_b200_pagerank_multi against the drop-in spmv_b200_pagerank."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c", "_build", "pagerank_multi_test")
    if not os.path.exists(exe):
        pytest.skip("tests/c/_build/pagerank_multi_test not built (run __graft_entry__.build())")
    p = subprocess.run([exe, "1", "15"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and "PASS" in p.stdout, p.stdout[-2000:]
