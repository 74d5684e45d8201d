"""Design probe: L2 levers for products whose x does not fit L2 (one GPU).
  * cudaLimitMaxL2FetchGranularity 64 (default) vs 32 bytes      (spmv_b200_set_l2_fetch_granularity)
  * SPMV_B200_HOT_L2 = 0 / 1 / 3: stream loads evict_first, x gathers evict_last in the hub-column kernel
on  c3 (config 3, MERGE_PATH), an R-MAT 26 shard (1/8, un-permuted and relabelled) and the whole graph.

    python scripts/l2_probe.py            # runs every combination in child processes
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    what, gran = sys.argv[2], int(sys.argv[3])
    import ctypes as C
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    torch.zeros(1, device=dev)
    assert sp.lib.spmv_b200_set_l2_fetch_granularity(gran) == 0
    got = sp.lib.spmv_b200_get_l2_fetch_granularity()
    if what == "c3":
        n = rows = 50_000_000
        rp, ci, va = gen.short_rows_with_outliers_csr(n, 43, dev)
        x = gen.uniform_01_open_low(5, torch.arange(n, device=dev), 9)
    else:
        parts, rank, relab = {"shard": (8, 3, False), "shard_relabelled": (8, 3, True), "whole": (1, 0, False)}[what]
        n, bounds, rp, ci, va, _ = gen.rmat_pagerank_shard(26, 16, 45, rank, parts, dev, row_weight=2, relabelled=relab)
        rows = bounds[rank + 1] - bounds[rank]
        x = torch.full((n,), 1.0 / n, device=dev)
    torch.cuda.synchronize()
    A = sp.DeviceCSR(rows, n, rp, ci, va)
    y = torch.empty(rows, device=dev)
    cfg = sp.make_config(sp.MERGE_PATH)
    if what == "c3":
        run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), None)  # noqa: E731
        mode = "merge"
    else:
        plan = sp.CsrPlan(A.ptr)
        run = lambda: plan.spmv(x, y)  # noqa: E731
        mode = plan.info()[2]
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        run()
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"what": what, "l2_fetch_granularity": got, "hot_l2": os.environ.get("SPMV_B200_HOT_L2", "default"),
                      "mode": mode, "ms": round(t0.elapsed_time(t1) / 10, 4)}), flush=True)
    sys.exit(0)

for what in ("c3", "shard", "shard_relabelled", "whole"):
    for gran in (64, 32):
        for hot_l2 in (("0", "1", "3") if what != "c3" else ("0",)):
            env = dict(os.environ, SPMV_B200_HOT_L2=hot_l2)
            subprocess.run([sys.executable, __file__, "child", what, str(gran)], env=env)
