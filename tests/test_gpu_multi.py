"""Sharded PageRank on >= 2 GPUs of one box (one process per GPU, NCCL): the NCCL all-gather
path, the fused peer-store exchange and the fused NVSwitch-multicast exchange must all reproduce
the single-GPU result, and must be bitwise identical to each other (same kernels, same summation
order, only the transport differs).
Skipped on boxes with a single GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


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
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n, rp, ci, va = gen.rmat_pagerank_csr(scale, 16, 7, dev)
        bounds = D.partition_rows(rp, world)
        srp, sci, sva = D.extract_shard(rp, ci, va, bounds[rank], bounds[rank + 1])
        torch.cuda.synchronize()
        results = {}
        used_multicast = False
        for mode in ("nccl", "p2p", "multicast"):
            shard = D.CudaShard(n, bounds[rank], srp, sci, sva)
            out = D.pagerank_sharded(shard, bounds, 0.85, 1e-6, 100,
                                     fused_exchange={"nccl": False, "p2p": True, "multicast": "multicast"}[mode])
            torch.cuda.synchronize()
            results[mode] = (out.ranks.clone(), out.iterations, out.final_residual, out.converged)
            if mode == "multicast":
                used_multicast = bool(shard._multicast)  # False: no NVSwitch multicast here, peer stores were used
            dist.barrier()
            shard.disable_peer_exchange()
            shard.disable_multicast_exchange()
            shard.close()
        a, b, c = results["nccl"], results["p2p"], results["multicast"]
        assert a[1:] == b[1:] == c[1:], (a[1:], b[1:], c[1:])
        assert torch.equal(a[0], b[0]), "fused peer-store exchange differs from the NCCL all-gather path"
        assert torch.equal(a[0], c[0]), "multicast exchange differs from the NCCL all-gather path"
        if rank == 0:
            print(f"multicast exchange exercised: {used_multicast}")
        np.save(os.path.join(out_dir, f"ranks_{rank}.npy"), a[0].cpu().numpy())
        np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([a[1], a[2], float(a[3])]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scale", [14, 18])
def test_sharded_pagerank_nccl_and_fused_exchange(sp, orc, cuda, tmp_path, scale):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), scale, str(tmp_path)), nprocs=world, join=True)
    import gpu_spmv_b200.gen as gen
    n, rp, ci, va = gen.rmat_pagerank_csr(scale, 16, 7, "cpu")
    ranks = [np.load(tmp_path / f"ranks_{r}.npy") for r in range(world)]
    meta = [np.load(tmp_path / f"meta_{r}.npy") for r in range(world)]
    for r in range(1, world):  # every rank ends with the same full vector
        assert np.array_equal(ranks[r].view(np.uint32), ranks[0].view(np.uint32)) and np.array_equal(meta[r], meta[0])
    iters = int(meta[0][0])
    o_same, _, l2, l1, conv = orc.pagerank_f64(n, n, rp.numpy(), ci.numpy(), va.numpy(), 0.85, 1e-6, 100, fixed_it=iters)
    assert bool(meta[0][2]) and np.abs(ranks[0].astype(np.float64) - o_same).sum() <= 1e-6
