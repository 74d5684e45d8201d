"""Sharded PageRank through the native C++ path (spmv_b200_pr_dist_*), one process per GPU.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dist_pagerank_check.py --scale 20 --iters 10 [--check] [--exchanges multicast,p2p,nccl]

Uses only RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT from the launcher: the rendezvous is the library's
own (abstract unix socket), torch.distributed is not initialised.  --check: rank 0 builds the whole graph
on the CPU and compares the result of every transport with the f64-accumulator restatement of the
reference recurrence (oracle/, checker use only) at equal iteration count: L1 <= 1e-6; the transports
must agree bit for bit."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from _load_pkg import load_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--seed", type=int, default=45)
ap.add_argument("--row-weight", type=int, default=4)
ap.add_argument("--check", action="store_true")
ap.add_argument("--converge", action="store_true", help="also run with the stop rule (tol 1e-6, <= 100 iterations)")
ap.add_argument("--exchanges", default="multicast,p2p,nccl")
ap.add_argument("--repeat", type=int, default=2)
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sp = load_pkg()
import gpu_spmv_b200.dist as D  # noqa: E402
import gpu_spmv_b200.gen as gen  # noqa: E402

comm = D.NativeComm(rank, world, f"check-{os.environ.get('MASTER_PORT', '0')}")
t0 = time.perf_counter()
n, bounds, rp, ci, va, n_edges = gen.rmat_pagerank_shard(args.scale, 16, args.seed, rank, world, dev, row_weight=args.row_weight)
torch.cuda.synchronize()
csr = sp.DeviceCSR(bounds[rank + 1] - bounds[rank], n, rp, ci, va)
if rank == 0:
    print(f"[check] R-MAT {args.scale}: n={n} nnz={n_edges} world={world} bounds={bounds} build {time.perf_counter() - t0:.1f}s", flush=True)

results, vectors = {}, {}
names = {"multicast": D.EXCHANGE_MULTICAST, "p2p": D.EXCHANGE_P2P, "nccl": D.EXCHANGE_NCCL}
for name in args.exchanges.split(","):
    if world == 1 and name != "p2p":
        continue
    pr = D.NativeShardedPageRank(comm, csr, bounds[rank], n, names[name])
    used = D.EXCHANGE_NAMES[pr.exchange]
    best = None
    for rep in range(args.repeat):
        res = pr.run(0.85, 0.0, 0, fixed_iterations=args.iters)
        sec = max(comm.allgather_doubles(res.device_seconds))
        best = sec if best is None else min(best, sec)
    vec = pr.ranks(dev).clone()
    entry = {"requested": name, "used": used, "iters": res.iterations, "launched": res.iterations_launched,
             "ms_per_iter": best / args.iters * 1e3, "iters_per_s": args.iters / best, "l2_residual": res.final_residual,
             "graph": res.graph_replay, "kernels_per_iteration": res.kernels_per_iteration, "hub_columns": pr.hub_columns}
    if args.converge:
        res2 = pr.run(0.85, 1e-6, 100)
        entry.update({"converged": res2.converged, "conv_iterations": res2.iterations, "conv_residual": res2.final_residual})
    results[name] = entry
    vectors[name] = vec
    pr.close()
    if rank == 0:
        print("[check]", json.dumps(entry), flush=True)

ok = True
keys = list(vectors)
for k in keys[1:]:
    same = bool(torch.equal(vectors[keys[0]].view(torch.int32), vectors[k].view(torch.int32)))
    if rank == 0:
        print(f"[check] {keys[0]} vs {k}: bit-identical = {same}", flush=True)
    ok = ok and same
# every rank holds the same vector
if keys:
    mine = float(vectors[keys[0]].double().sum().item())
    sums = comm.allgather_doubles(mine)
    if rank == 0:
        print(f"[check] sum of ranks per rank: {sums}", flush=True)
    ok = ok and all(abs(s - sums[0]) == 0.0 for s in sums) and abs(sums[0] - 1.0) <= 1e-6
if args.check and rank == 0 and keys:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle
    orc = Oracle()
    _, frp, fci, fva = gen.rmat_pagerank_csr(args.scale, 16, args.seed, "cpu")
    o_ranks, o_it, o_l2, o_l1, _ = orc.pagerank_f64(n, n, frp.numpy(), fci.numpy(), fva.numpy(), 0.85, 1e-6, 100, fixed_it=args.iters)
    for k in keys:
        l1 = float(np.abs(vectors[k].cpu().numpy().astype(np.float64) - o_ranks).sum())
        print(f"[check] {k}: L1 distance to the f64 restatement after {args.iters} iterations = {l1:.3e} (<= 1e-6), "
              f"residual {results[k]['l2_residual']:.6e} vs {o_l2:.6e}", flush=True)
        ok = ok and l1 <= 1e-6 and abs(results[k]["l2_residual"] - o_l2) <= 1e-3 * o_l2 + 1e-12
flags = comm.allgather_doubles(1.0 if ok else 0.0)
comm.close()
if rank == 0:
    print("[check] RESULT", "PASS" if all(f == 1.0 for f in flags) else "FAIL", flush=True)
sys.exit(0 if all(f == 1.0 for f in flags) else 1)
