"""Matrix Market coordinate files (gpu-spmv_b200/csrc/matrix_market.cpp; SURVEY 8f rank 4 -- the
reference lists real-matrix input as a requirement, requirements.md:90, but has no loader).
Host-only: runs in the CPU suite."""
import numpy as np
import pytest


def write(path, text):
    with open(path, "w") as f:
        f.write(text)
    return str(path)


def arrays(sp, A):
    rp, ci, va = sp.csr_arrays(A)
    return rp.copy(), ci.copy(), va.copy()


def test_general_real_is_sorted_and_keeps_duplicates_in_file_order(sp, tmp_path):
    p = write(tmp_path / "a.mtx", "%%MatrixMarket matrix coordinate real general\n% a comment\n\n"
                                  "3 4 6\n3 4 5.5\n1 3 2\n1 1 1\n2 2 -3\n1 3 7\n3 1 0.25\n")
    A = sp.csr_create(0, 0, 0)
    assert sp.csr_load_matrix_market(A, p) == 0
    m = A.contents
    assert (m.num_rows, m.num_cols, m.nnz) == (3, 4, 6) and m.owns_host_memory
    rp, ci, va = arrays(sp, A)
    assert rp.tolist() == [0, 3, 4, 6]
    assert ci.tolist() == [0, 2, 2, 1, 0, 3]
    assert va.tolist() == [1.0, 2.0, 7.0, -3.0, 0.25, 5.5]  # the two (1,3) entries in file order
    _, dense = sp.csr_to_dense(A)  # duplicate (row, col): last wins, as csr_to_dense does (src/csr_matrix.cpp:97-114)
    assert dense[0, 2] == 7.0 and dense[2, 3] == 5.5
    sp.csr_destroy(A)


def test_symmetric_skew_pattern_and_integer(sp, tmp_path):
    A = sp.csr_create(0, 0, 0)
    p = write(tmp_path / "s.mtx", "%%MatrixMarket matrix coordinate integer symmetric\n3 3 3\n1 1 4\n3 1 2\n3 2 -1\n")
    assert sp.csr_load_matrix_market(A, p) == 0
    assert np.array_equal(sp.csr_to_dense(A)[1], np.array([[4, 0, 2], [0, 0, -1], [2, -1, 0]], np.float32))
    p = write(tmp_path / "k.mtx", "%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 3.5\n")
    assert sp.csr_load_matrix_market(A, p) == 0
    assert np.array_equal(sp.csr_to_dense(A)[1], np.array([[0, -3.5], [3.5, 0]], np.float32))
    p = write(tmp_path / "p.mtx", "%%MatrixMarket matrix coordinate pattern general\n2 3 2\n1 3\n2 1\n")
    assert sp.csr_load_matrix_market(A, p) == 0
    assert np.array_equal(sp.csr_to_dense(A)[1], np.array([[0, 0, 1], [1, 0, 0]], np.float32))
    p = write(tmp_path / "e.mtx", "%%MatrixMarket matrix coordinate real general\n4 2 0\n")
    assert sp.csr_load_matrix_market(A, p) == 0
    assert (A.contents.num_rows, A.contents.nnz) == (4, 0) and sp.csr_arrays(A)[0].tolist() == [0] * 5
    sp.csr_destroy(A)


def test_round_trip_is_bit_exact_and_matches_from_dense(sp, orc, tmp_path):
    rng = np.random.default_rng(3)
    dense = np.where(rng.random((37, 53)) < 0.2, rng.uniform(-10, 10, (37, 53)), 0).astype(np.float32)
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, dense, 37, 53)
    p = str(tmp_path / "r.mtx")
    assert sp.csr_save_matrix_market(A, p) == 0
    B = sp.csr_create(0, 0, 0)
    assert sp.csr_load_matrix_market(B, p) == 0
    for a, b in zip(arrays(sp, A), arrays(sp, B)):
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    x = rng.uniform(-1, 1, 53).astype(np.float32)
    assert np.array_equal(sp.spmv_cpu_csr(A, x).view(np.uint32), sp.spmv_cpu_csr(B, x).view(np.uint32))
    sp.csr_destroy(A)
    sp.csr_destroy(B)


@pytest.mark.parametrize("text", [
    "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
    "%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n",
    "%%MatrixMarket matrix coordinate real hermitian\n1 1 1\n1 1 1\n",
    "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n",   # row out of range
    "%%MatrixMarket matrix coordinate real general\n2 2 1\n1 0 1.0\n",   # 0-based column
    "%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n",   # fewer entries than announced
    "%%MatrixMarket matrix coordinate real general\nnot a size line\n",
    "MatrixMarket matrix coordinate real general\n1 1 0\n",
    "",
])
def test_rejected_files_leave_the_matrix_unchanged(sp, tmp_path, text):
    A = sp.csr_create(0, 0, 0)
    sp.csr_from_dense(A, np.array([[1, 0], [0, 2]], np.float32), 2, 2)
    before = [a.tolist() for a in arrays(sp, A)]
    assert sp.csr_load_matrix_market(A, write(tmp_path / "bad.mtx", text)) == int(sp.SpMVError.INVALID_FORMAT)
    assert [a.tolist() for a in arrays(sp, A)] == before
    assert sp.csr_load_matrix_market(A, str(tmp_path / "missing.mtx")) == int(sp.SpMVError.FILE_IO)
    assert sp.csr_load_matrix_market(None, b"x") == int(sp.SpMVError.INVALID_ARGUMENT)
    sp.csr_destroy(A)
