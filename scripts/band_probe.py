"""Design probe (not product): what would COLUMN BANDS buy when x does not fit L2?

A shard of the R-MAT graph (or config 3) is split by column into B bands of equal width; every
band is an ordinary CSR matrix over the same rows, multiplied by the planned merge-path kernel
(hub-column plan).  sum of the band times vs the unsplit product = upper bound of the gain of a
band plan (the real one adds a read-modify-write of y per extra band).

    python scripts/band_probe.py --what rmat --scale 26 --parts 8 --rank 3 --bands 1,2,3,4 [--relabelled]
    python scripts/band_probe.py --what c3 --bands 1,2,3
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from _load_pkg import load_pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="rmat")
ap.add_argument("--scale", type=int, default=26)
ap.add_argument("--parts", type=int, default=8)
ap.add_argument("--rank", type=int, default=3)
ap.add_argument("--bands", default="1,2,3,4")
ap.add_argument("--relabelled", action="store_true")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()

sp = load_pkg()
import gpu_spmv_b200.gen as gen  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
if args.what == "rmat":
    n, bounds, rp, ci, va, _ = gen.rmat_pagerank_shard(args.scale, 16, 45, args.rank, args.parts, dev, row_weight=2,
                                                       relabelled=args.relabelled)
    rows = bounds[args.rank + 1] - bounds[args.rank]
    x = torch.full((n,), 1.0 / n, device=dev)
else:
    n = rows = 50_000_000
    rp, ci, va = gen.short_rows_with_outliers_csr(n, 43, dev)
    x = gen.uniform_01_open_low(5, torch.arange(n, device=dev), 9)
torch.cuda.synchronize()
counts = (rp[1:] - rp[:-1]).to(torch.int64)
row_of = torch.repeat_interleave(torch.arange(rows, dtype=torch.int32, device=dev), counts)
y = torch.empty(rows, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / args.steps


for B in [int(b) for b in args.bands.split(",")]:
    width = (n + B - 1) // B
    total, parts = 0.0, []
    for b in range(B):
        lo, hi = b * width, min(n, (b + 1) * width)
        if B == 1:
            rp_b, ci_b, va_b = rp, ci, va
        else:
            m = (ci >= lo) & (ci < hi)
            ci_b, va_b = ci[m].contiguous(), va[m].contiguous()
            cnt = torch.bincount(row_of[m].to(torch.int64), minlength=rows)
            rp_b = torch.zeros(rows + 1, dtype=torch.int32, device=dev)
            rp_b[1:] = torch.cumsum(cnt, 0).to(torch.int32)
            del m, cnt
        A = sp.DeviceCSR(rows, n, rp_b, ci_b, va_b)
        plan = sp.CsrPlan(A.ptr)
        ms = timed(lambda: plan.spmv(x, y))
        parts.append({"band": b, "nnz": int(ci_b.numel()), "ms": round(ms, 4), "plan": plan.info()})
        total += ms
        plan.close()
        del A
        if B > 1:
            del rp_b, ci_b, va_b
        torch.cuda.empty_cache()
    print(json.dumps({"what": args.what, "scale": args.scale, "parts": args.parts, "rank": args.rank, "relabelled": args.relabelled,
                      "rows": rows, "nnz": int(ci.numel()), "bands": B, "total_ms": round(total, 4), "per_band": parts}), flush=True)
