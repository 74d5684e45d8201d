"""Device-side CSR assembly: COO triplets -> CSR sorted by (row, col), and the
column normalisation pagerank() expects (reference include/spmv/pagerank.h:28).  Index arrays and
row_ptrs must be BIT-EXACT against a stable host sort (format conversion is integer work); values
travel unchanged, and 1/outdeg values match the generator's bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def host_csr(rows, r, c, v):
    order = np.lexsort((c, r))  # stable: duplicates keep their input order
    rp = np.zeros(rows + 1, np.int32)
    rp[1:] = np.cumsum(np.bincount(r, minlength=rows))
    return rp, c[order].astype(np.int32), v[order].astype(np.float32)


@pytest.mark.parametrize("rows,cols,n,seed", [(1, 1, 1, 0), (7, 5, 0, 1), (1000, 300, 5000, 2), (100003, 65537, 400000, 3),
                                             (50, 70000, 200000, 4), (200000, 3, 100000, 5)])
def test_coo_to_csr_matches_stable_host_sort(sp, cuda, rows, cols, n, seed):
    rng = np.random.default_rng(seed)
    r = rng.integers(0, rows, n).astype(np.int32)
    c = rng.integers(0, cols, n).astype(np.int32)
    if n > 10:  # force duplicates with distinct values: their order must be the input order
        r[1::7], c[1::7] = r[0], c[0]
    v = rng.uniform(-1, 1, n).astype(np.float32)
    A = sp.csr_create(0, 0, 0)
    t = lambda a: torch.as_tensor(a).to(cuda)  # noqa: E731
    rc = sp.csr_from_coo_device(A, rows, cols, t(r), t(c), t(v))
    assert rc == 0, sp.spmv_error_string(rc)
    m = A.contents
    assert (m.num_rows, m.num_cols, m.nnz) == (rows, cols, n) and m.owns_device_memory and m.owns_host_memory
    assert sp.csr_from_gpu(A) == 0
    rp, ci, va = sp.csr_arrays(A)
    e_rp, e_ci, e_va = host_csr(rows, r, c, v)
    assert np.array_equal(rp, e_rp)
    if n:
        assert np.array_equal(ci, e_ci) and np.array_equal(va.view(np.uint32), e_va.view(np.uint32))
    # the assembled matrix is a normal citizen of the API
    x = rng.uniform(-1, 1, cols).astype(np.float32)
    d_y = torch.empty(rows, dtype=torch.float32, device=cuda)
    res = sp.spmv_csr(A, t(x), d_y, sp.make_config(sp.SCALAR_CSR), cols)
    if n:
        assert res.error_code == 0
        assert np.array_equal(d_y.cpu().numpy().view(np.uint32), sp.spmv_cpu_csr(A, x).view(np.uint32))
    sp.csr_destroy(A)


def test_rmat_assembly_and_column_normalisation_match_generator(sp, cuda):
    import gpu_spmv_b200.gen as gen
    scale, ef, seed = 13, 16, 9
    n, rp, ci, va = gen.rmat_pagerank_csr(scale, ef, seed, "cpu")
    src, dst = gen.rmat_edges(scale, ef, seed, cuda)
    A = sp.csr_create(0, 0, 0)
    ones = torch.ones(src.numel(), dtype=torch.float32, device=cuda)
    assert sp.csr_from_coo_device(A, n, n, dst.to(torch.int32), src.to(torch.int32), ones) == 0
    assert sp.csr_normalize_columns_device(A) == 0
    assert sp.csr_from_gpu(A) == 0
    g_rp, g_ci, g_va = sp.csr_arrays(A)
    assert np.array_equal(g_rp, rp.numpy()) and np.array_equal(g_ci, ci.numpy())
    assert np.array_equal(g_va.view(np.uint32), va.numpy().view(np.uint32))
    sp.csr_destroy(A)


def test_bad_index_is_rejected_and_leaves_the_matrix_alone(sp, cuda):
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, np.array([[1, 0], [0, 2]], np.float32), 2, 2)
    r = torch.tensor([0, 5], dtype=torch.int32, device=cuda)
    c = torch.tensor([0, 1], dtype=torch.int32, device=cuda)
    v = torch.ones(2, device=cuda)
    assert sp.csr_from_coo_device(A, 3, 3, r, c, v) == int(sp.SpMVError.INVALID_ARGUMENT)
    assert (A.contents.num_rows, A.contents.nnz) == (2, 2)
    assert sp.csr_from_coo_device(A, 3, 3, c, torch.tensor([0, -1], dtype=torch.int32, device=cuda), v) \
        == int(sp.SpMVError.INVALID_ARGUMENT)
    sp.csr_destroy(A)
