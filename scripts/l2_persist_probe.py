#!/usr/bin/env python
"""Design probe: L2 PERSISTENCE for the x of products whose x does not fit L2 (one GPU) -- a persisting set-aside
(cudaLimitPersistingL2CacheSize) plus an access-policy window over x on the launching stream (hit ratio = share of
the window that persists, the rest streams).  On c3 (config 3, MERGE_PATH), an R-MAT 26 shard (1/8, un-permuted and
relabelled) and the whole graph.  Output -> profiles/r2_l2_persist.jsonl

    python scripts/l2_persist_probe.py            # every combination in child processes
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    what = sys.argv[2]
    import ctypes as C
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    torch.zeros(1, device=dev)
    lim = [C.c_ulonglong() for _ in range(3)]
    assert sp.lib.spmv_b200_l2_persistence_limits(*[C.byref(v) for v in lim]) == 0
    max_aside, max_window, l2 = [v.value for v in lim]
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
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream
    if what == "c3":
        run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), C.c_void_p(sh))  # noqa: E731
        mode = "merge"
    else:
        plan = sp.CsrPlan(A.ptr)
        run = lambda: plan.spmv(x, y, sh)  # noqa: E731
        mode = plan.info()[2]
    x_bytes = x.numel() * 4
    y_ref = None
    for aside_frac, ratio in ((0.0, 0.0), (1.0, None), (1.0, 1.0), (0.75, None), (0.5, None), (1.0, 0.2)):
        aside = int(max_aside * aside_frac)
        window = min(x_bytes, max_window)
        hit = (min(1.0, aside / window) if ratio is None else ratio) if aside else 0.0
        if aside:
            rc = sp.lib.spmv_b200_set_l2_persistence(C.c_void_p(sh), C.c_void_p(x.data_ptr()), window, hit, aside)
        else:
            rc = sp.lib.spmv_b200_set_l2_persistence(C.c_void_p(sh), None, 0, 0.0, 0)
        with torch.cuda.stream(stream):
            for _ in range(3):
                assert run() == 0
            stream.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for _ in range(10):
                run()
            t1.record(stream)
            stream.synchronize()
        same = None
        if y_ref is None:
            y_ref = y.clone()
        else:
            same = bool(torch.equal(y.view(torch.int32), y_ref.view(torch.int32)))
        print(json.dumps({"what": what, "mode": mode, "x_mb": round(x_bytes / 1e6, 1), "l2_mb": round(l2 / 1e6, 1),
                          "max_set_aside_mb": round(max_aside / 1e6, 1), "max_window_mb": round(max_window / 1e6, 1),
                          "set_aside_mb": round(aside / 1e6, 1), "window_mb": round(window / 1e6, 1) if aside else 0,
                          "hit_ratio": round(hit, 3), "rc": rc, "ms": round(t0.elapsed_time(t1) / 10, 4),
                          "bit_identical_to_no_window": same}), flush=True)
    sys.exit(0)

for what in sys.argv[1:] or ("c3", "shard", "shard_relabelled", "whole"):
    subprocess.run([sys.executable, __file__, "child", what])
