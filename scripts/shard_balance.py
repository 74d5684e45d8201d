"""Cost model of a PageRank row shard, measured on ONE GPU: for every rank p of a P-way partition
(work(row) = nnz + w) build that shard of the R-MAT graph and time its LOCAL fused step
(spmv_b200_pr_step: no exchange, no waiting for peers).  The slowest shard bounds the sharded
iteration, so the table shows which row weight balances the ranks.

    python scripts/shard_balance.py --scale 26 --parts 2 --weights 1,4,8,16 [--relabelled]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from _load_pkg import load_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=26)
ap.add_argument("--seed", type=int, default=45)
ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--weights", default="1,4,8,16")
ap.add_argument("--ranks", default="", help="comma list of ranks to time (default: all)")
ap.add_argument("--relabelled", action="store_true")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

sp = load_pkg()
import gpu_spmv_b200.dist as D  # noqa: E402
import gpu_spmv_b200.gen as gen  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
ranks = [int(r) for r in args.ranks.split(",")] if args.ranks else list(range(args.parts))
for w in [int(v) for v in args.weights.split(",")]:
    rows_out = []
    for p in ranks:
        n, bounds, rp, ci, va, n_edges = gen.rmat_pagerank_shard(args.scale, 16, args.seed, p, args.parts, dev, row_weight=w,
                                                                 relabelled=args.relabelled)
        torch.cuda.synchronize()
        shard = D.CudaShard(n, bounds[p], rp, ci, va)
        r_a = torch.full((n,), 1.0 / n, device=dev)
        r_b = torch.empty(n, device=dev)
        partial = torch.zeros(3, dtype=torch.float64, device=dev)
        for _ in range(3):
            shard(r_a, r_b, partial)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            shard(r_a, r_b, partial)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.steps
        rows_p, nnz_p = bounds[p + 1] - bounds[p], ci.numel()
        rows_out.append({"rank": p, "rows": rows_p, "nnz": nnz_p, "ms": round(ms, 4),
                         "ns_per_nnz": round(ms * 1e6 / max(nnz_p, 1), 4)})
        shard.close()
        del shard, rp, ci, va, r_a, r_b
        torch.cuda.empty_cache()
    print(json.dumps({"scale": args.scale, "parts": args.parts, "row_weight": w, "relabelled": args.relabelled,
                      "max_ms": max(r["ms"] for r in rows_out), "shards": rows_out}), flush=True)
