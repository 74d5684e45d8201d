"""The executable model of the segmented-stream kernel's index logic (scripts/seg_model.py): runs of 4
non-zeros per lane, two steps per warp, per-span head counts, lead / tail / closed segments, the
cross-warp fold and the cross-tile fix-up, checked against direct row sums on random CSR structures
(empty rows, rows longer than several tiles, ragged ends).  CPU only: it guards the algorithm the CUDA
kernel (gpu-spmv_b200/csrc/csr_seg_kernels.cu) transcribes; the kernel itself is checked by
tests/test_gpu_hot.py."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_model():
    spec = importlib.util.spec_from_file_location("seg_model", os.path.join(ROOT, "scripts", "seg_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_model_reproduces_row_sums():
    m = load_model()
    rng = np.random.default_rng(11)
    shapes = [np.array([3, 0, 0, 5000, 1, 0, 2]), np.zeros(40, int), np.full(700, 4), rng.integers(0, 9, 3000),
              np.where(rng.random(2500) < 0.6, 0, rng.integers(1, 50, 2500)), np.array([1] + [0] * 10 + [9000])]
    for lens in shapes:
        rows = len(lens)
        rp = np.zeros(rows + 1, np.int64)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        prod = rng.integers(-8, 9, nnz).astype(np.float64)  # exact in any summation order
        expect = np.array([prod[rp[i]:rp[i + 1]].sum() for i in range(rows)])
        got = m.spmv_model(rows, rp, prod) if nnz else np.zeros(rows)
        assert np.array_equal(got, expect), (rows, nnz)


def test_plan_tables_of_the_model():
    m = load_model()
    rp = np.array([0, 0, 3, 3, 2051, 2051, 4100])
    head, rows_nz, thb, num_tiles = m.build_plan(6, rp)
    assert rows_nz.tolist() == [1, 3, 5] and num_tiles == 3
    assert np.nonzero(head)[0].tolist() == [0, 3, 2051]
    assert thb.tolist() == [0, 2, 3, 3]  # heads before non-zero 0, 2048, 4096 and 6144 (the closing entry)
