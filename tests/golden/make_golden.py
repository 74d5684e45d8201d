"""Generates tests/golden/*.npz from the UNMODIFIED reference library
(oracle/_ref/libspmv_ref.so, built from /root/reference by oracle/Makefile).

Run in the authoring container only (the reference tree does not exist on the
GPU box):   python tests/golden/make_golden.py

What is stored (inputs exactly as the reference's own tests generate them,
outputs exactly as the reference's own code computes them on the CPU):

  spmv_property_cases.npz  the first 12 inputs of SpMVPropertyTest.CSRCorrectness
      (tests/test_spmv.cu:40-48: RandomGenerator rng{42}; rows, cols in
      randInt(1,200); density randFloat(.01,.3); generateRandomDenseMatrix;
      generateRandomVector) with csr_from_dense, csr_compute_stats,
      spmv_auto_config, spmv_cpu_csr, ell_from_dense, ell_from_csr and
      spmv_cpu_ell outputs.
  config1_random10k.npz    BASELINE config 1: generateRandomDenseMatrix(10000,
      10000, 0.001f, RandomGenerator{42}) then generateRandomVector(10000),
      as CSR + x + spmv_cpu_csr output + stats/selector.
  pagerank_cases.npz       the first 6 inputs of PageRankPropertyTest.ScoreInvariants
      (tests/test_pagerank.cu:18-40) as CSR (column-normalised), for the
      PageRank parity tests (reference pagerank() itself needs a GPU; its CPU
      restatement is pinned by construction against spmv_cpu_csr here).
  known_answers.npz        README / design.md / unit-test literals.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_binding import Oracle, Ref  # noqa: E402


def main():
    assert Ref.available(), "build oracle/_ref first: make -C oracle"
    ref, orc = Ref(), Oracle()

    # ---- SpMVPropertyTest.CSRCorrectness replay --------------------------------
    ref.rng_seed(42)
    out = {}
    n_cases = 12
    for it in range(n_cases):
        rows = ref.rng_int(1, 200)
        cols = ref.rng_int(1, 200)
        density = ref.rng_float(0.01, 0.3)
        dense = ref.rng_dense(rows, cols, density)
        x = ref.rng_vector(cols)
        h, st = ref.csr_from_dense(dense)
        assert st == 0
        r, c, nnz, rp, ci, va = ref.csr_fields(h)
        y = ref.spmv_cpu_csr(h, x, rows)
        avg, mx, mn, skew = ref.stats(h)
        kt, bs, tex = ref.auto_config(h)
        he, st = ref.ell_from_dense(dense)
        _, _, w, ec, ev = ref.ell_fields(he)
        he2, st2 = ref.ell_from_csr(h)
        _, _, w2, ec2, ev2 = ref.ell_fields(he2)
        assert w == w2 and np.array_equal(ec, ec2) and np.array_equal(ev.view(np.uint32), ev2.view(np.uint32))
        y_ell = ref.spmv_cpu_ell(he, x, rows)
        p = f"c{it}_"
        out.update({p + "dense": dense, p + "x": x, p + "row_ptrs": rp, p + "col_indices": ci, p + "values": va,
                    p + "y_csr": y, p + "stats": np.array([avg, mx, mn, skew], np.float64),
                    p + "density": np.array([density], np.float32),
                    p + "selector": np.array([kt, bs, int(tex)], np.int32), p + "ell_width": np.array([w], np.int32),
                    p + "ell_cols": ec, p + "ell_values": ev, p + "y_ell": y_ell})
        ref.L.ref_csr_destroy(h)
        ref.L.ref_ell_destroy(he)
        ref.L.ref_ell_destroy(he2)
        if it < 3:
            print(f"iter {it}: {rows}x{cols} density {density:.9g} nnz {nnz} avg {avg:.9g} max {mx} min {mn} "
                  f"skew {skew:.9g} kernel {kt} fnv(rp) {orc.fnv(rp):016x} fnv(ci) {orc.fnv(ci):016x} "
                  f"fnv(va) {orc.fnv(va):016x} fnv(y) {orc.fnv(y):016x} y[0] {y[0]:.9g}")
    out["n_cases"] = np.array([n_cases], np.int32)
    np.savez_compressed(os.path.join(HERE, "spmv_property_cases.npz"), **out)

    # ---- BASELINE config 1 ------------------------------------------------------
    ref.rng_seed(42)
    dense = ref.rng_dense(10000, 10000, np.float32(0.001))
    x = ref.rng_vector(10000)
    h, st = ref.csr_from_dense(dense)
    r, c, nnz, rp, ci, va = ref.csr_fields(h)
    y = ref.spmv_cpu_csr(h, x, r)
    avg, mx, mn, skew = ref.stats(h)
    kt, bs, tex = ref.auto_config(h)
    print(f"config1: nnz {nnz} avg {avg:.9g} max {mx} min {mn} skew {skew:.9g} kernel {kt} tex {tex}")
    np.savez_compressed(os.path.join(HERE, "config1_random10k.npz"), row_ptrs=rp, col_indices=ci, values=va, x=x,
                        y_csr=y, stats=np.array([avg, mx, mn, skew], np.float64),
                        selector=np.array([kt, bs, int(tex)], np.int32))
    ref.L.ref_csr_destroy(h)
    del dense

    # ---- PageRankPropertyTest.ScoreInvariants replay ---------------------------
    ref.rng_seed(42)
    out = {}
    n_pr = 6
    for it in range(n_pr):
        n = ref.rng_int(5, 50)
        density = ref.rng_float(0.1, 0.5)
        adj = ref.rng_dense(n, n, density, 0.0, 1.0)
        for j in range(n):  # column normalisation exactly as tests/test_pagerank.cu:27-37 (fp32, sequential)
            s = np.float32(0.0)
            for i in range(n):
                s = np.float32(s + adj[i, j])
            if s > 0.0:
                for i in range(n):
                    adj[i, j] = np.float32(adj[i, j] / s)
        h, st = ref.csr_from_dense(adj)
        r, c, nnz, rp, ci, va = ref.csr_fields(h)
        p = f"p{it}_"
        out.update({p + "n": np.array([n], np.int32), p + "row_ptrs": rp, p + "col_indices": ci, p + "values": va})
        ref.L.ref_csr_destroy(h)
    out["n_cases"] = np.array([n_pr], np.int32)
    np.savez_compressed(os.path.join(HERE, "pagerank_cases.npz"), **out)

    # ---- literals from the reference's README / design.md / unit tests ----------
    known = {
        # README 3x3 / tests/test_ell.cpp:153-172
        "readme_dense": np.array([[1, 0, 2], [0, 3, 4], [0, 0, 5]], np.float32),
        "readme_y_ones": np.array([3, 7, 5], np.float32),
        "readme_ell_cols": np.array([0, 1, 2, 2, 2, -1], np.int32),
        # design.md:372-385 3x4 layout example
        "design_values": np.array([1, 2, 3, 4, 5], np.float32),
        "design_cols": np.array([0, 2, 1, 2, 3], np.int32),
        "design_row_ptrs": np.array([0, 2, 4, 5], np.int32),
        "design_ell_values": np.array([1, 3, 5, 2, 4, 0], np.float32),
        "design_ell_cols": np.array([0, 1, 3, 2, 2, -1], np.int32),
        # tests/test_spmv.cu:188-218
        "unit_dense": np.array([[1, 2, 0], [0, 0, 0], [3, 0, 4]], np.float32),
        "unit_y_ones": np.array([3, 0, 7], np.float32),
    }
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **known)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
