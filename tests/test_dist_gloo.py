"""Host logic of the N > 1 path on CPU (gloo, world_size 2): nnz-balanced
partitioning, shard extraction (bit-exact concatenation), the variable-length
slice all-gather, the all-reduce of the partial sums and the stop rule of
gpu_spmv_b200.dist.pagerank_loop.  The local step is supplied by the checker
(oracle arithmetic on the rank's shard): the product's CUDA step cannot run
without a GPU and nothing here is a product fallback."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleStep:
    """Checker-side stand-in for CudaShard: same contract, numpy/oracle arithmetic."""

    def __init__(self, orc, n, lo, rp, ci, va, dangling, damping):
        self.orc, self.n, self.lo = orc, n, lo
        self.rows = len(rp) - 1
        self.rp, self.ci, self.va = rp, ci, va
        self.dangling = dangling.astype(bool)
        self.d = np.float32(damping)
        self.dsum = np.float32(0)

    def __call__(self, r_old, r_new, partial):
        x = r_old.numpy()
        y64, _ = self.orc.spmv_csr_f64(self.rows, self.rp, self.ci, self.va, x)
        y = y64.astype(np.float32)
        dc = np.float32(np.float32(self.d * self.dsum) / np.float32(self.n))
        tp = np.float32((np.float32(1.0) - self.d) / np.float32(self.n))
        v = ((self.d * y).astype(np.float32) + dc).astype(np.float32) + tp
        v = v.astype(np.float32)
        sl = slice(self.lo, self.lo + self.rows)
        diff = v.astype(np.float64) - x[sl].astype(np.float64)
        r_new.numpy()[sl] = v
        partial[0] = float((diff * diff).sum())
        partial[1] = float(np.abs(diff).sum())
        partial[2] = float(v[self.dangling[sl]].astype(np.float64).sum())

    def set_dangling_mass(self, partial):
        self.dsum = np.float32(partial[2].item())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, scale, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from _load_pkg import load_pkg
    load_pkg()
    import gpu_spmv_b200.dist as D
    import gpu_spmv_b200.gen as gen
    from oracle_binding import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        n, rp, ci, va = gen.rmat_pagerank_csr(scale, 8, 11, "cpu")
        bounds = D.partition_rows(rp, world)
        assert bounds == orc.partition_rows(n, rp.numpy(), world).tolist()
        srp, sci, sva = D.extract_shard(rp, ci, va, bounds[rank], bounds[rank + 1])
        # shards concatenate back to the global arrays bit for bit
        pieces = [None] * world
        dist.all_gather_object(pieces, (srp.numpy(), sci.numpy(), sva.numpy()))
        assert np.array_equal(np.concatenate([p[1] for p in pieces]), ci.numpy())
        assert np.array_equal(np.concatenate([p[2] for p in pieces]).view(np.uint32), va.numpy().view(np.uint32))
        rebuilt = np.concatenate([[0]] + [p[0][1:] + rp.numpy()[bounds[i]] for i, p in enumerate(pieces)])
        assert np.array_equal(rebuilt, rp.numpy())

        _, dangling = orc.find_dangling(n, n, rp.numpy(), ci.numpy(), va.numpy())
        step = OracleStep(orc, n, bounds[rank], srp.numpy(), sci.numpy(), sva.numpy(), dangling, 0.85)
        r_a = torch.full((n,), 1.0 / n, dtype=torch.float32)
        r_a[:] = torch.tensor(np.float32(1.0) / np.float32(n))
        r_b = torch.full((n,), float("nan"), dtype=torch.float32)
        step.dsum = np.float32(r_a.numpy()[dangling.astype(bool)].astype(np.float64).sum())
        partial = torch.zeros(3, dtype=torch.float64)
        fin, iters, residual, conv, l1 = D.pagerank_loop(step, r_a, r_b, partial, bounds, 0.85, 1e-6, 100)
        assert not torch.isnan(fin).any()  # every slice arrived
        total = np.float32(fin.numpy().astype(np.float64).sum())
        ranks = fin.numpy() / total
        np.save(os.path.join(out_dir, f"ranks_{rank}.npy"), ranks)
        np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([iters, residual, float(conv), l1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scale", [9, 12])
def test_sharded_pagerank_two_ranks_gloo(orc, tmp_path, scale):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, scale, str(tmp_path)), nprocs=2, join=True)
    from _load_pkg import load_pkg
    load_pkg()
    import gpu_spmv_b200.gen as gen
    n, rp, ci, va = gen.rmat_pagerank_csr(scale, 8, 11, "cpu")
    r0, r1 = np.load(tmp_path / "ranks_0.npy"), np.load(tmp_path / "ranks_1.npy")
    m0, m1 = np.load(tmp_path / "meta_0.npy"), np.load(tmp_path / "meta_1.npy")
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32)) and np.array_equal(m0, m1)
    o_ranks, o_it, o_l2, o_l1, o_conv = orc.pagerank_f64(n, n, rp.numpy(), ci.numpy(), va.numpy(), 0.85, 1e-6, 100)
    assert int(m0[0]) == o_it and bool(m0[2]) == o_conv
    assert np.abs(r0.astype(np.float64) - o_ranks).sum() <= 1e-6
    assert abs(m0[1] - o_l2) <= 1e-3 * o_l2 + 1e-9


def test_partition_properties(sp, orc):
    """Bounds are monotone, cover all rows, balance nnz, and match the oracle."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        rows = int(rng.integers(1, 500))
        lens = rng.integers(0, 30, rows)
        if rng.random() < 0.3:
            lens[int(rng.integers(0, rows))] = 5000
        rp = np.zeros(rows + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        for parts in (1, 2, 3, 8):
            for w in (1, 3):  # (rows + nnz)-balanced variants match the oracle and stay monotone
                bw = sp.partition_rows(rp, rows, parts, w)
                assert bw[0] == 0 and bw[-1] == rows and np.all(np.diff(bw) >= 0)
                assert np.array_equal(bw, orc.partition_rows(rows, rp, parts, w))
            b = sp.partition_rows(rp, rows, parts)
            assert b[0] == 0 and b[-1] == rows and np.all(np.diff(b) >= 0)
            assert np.array_equal(b, orc.partition_rows(rows, rp, parts))
            nnz = int(rp[-1])
            for p in range(1, parts):
                assert rp[b[p]] >= (nnz * p) // parts and (b[p] == 0 or rp[b[p] - 1] < (nnz * p) // parts or b[p] == b[p - 1])


def test_merge_path_search_matches_oracle(sp, orc):
    rng = np.random.default_rng(4)
    for _ in range(20):
        rows = int(rng.integers(1, 60))
        lens = rng.integers(0, 9, rows)
        rp = np.zeros(rows + 1, np.int32)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        for d in range(rows + nnz + 1):
            assert sp.merge_path_search(d, rp, rows, nnz) == orc.merge_path_search(d, rp, rows, nnz)
    with pytest.raises(ValueError):
        sp.merge_path_search(rows + nnz + 1, rp, rows, nnz)
