#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpMV hot path (see BASELINE.json / BASELINE.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--quick]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: one
ELL SpMV over BASELINE config 2 (5-point Laplacian on a 4096^2 grid, 16.7 M
rows, 83.9 M non-zeros, 805 MB of matrix -- larger than L2, so no flush is
needed between iterations).  `value` is whole-job effective HBM GB/s =
algorithmic bytes (reference src/bandwidth.cpp:66-75) x steps / device time,
inputs resident in HBM.  `e2e` is the same metric through the blocking
reference-facing C-ABI call with HOST x / y buffers (H2D of x and D2H of y in
the timed region).  `extra` carries the other BASELINE configurations
(CSR kernels on config 2, config 3, config 4 R-MAT SpMV -- plain merge-path and
through a hub-column plan --, PageRank iter/s).

N > 1 (one rank per GPU, NCCL): weak scaling for the headline -- every rank
owns a 16.7 M-row band (row shard) of a 4096 x (4096 N) Laplacian with the
vector replicated, no data-path collective; `extra.pagerank` is the
row-sharded PageRank with the rank-vector all-gather and the 3-double
all-reduce over NCCL (strong scaling: one R-MAT graph split over N ranks).

--impl reference times the reference's own CPU implementation of the step
(spmv_cpu_ell from oracle/_ref, the unmodified reference sources; the oracle
port if that library is absent) on the host, rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 4096            # BASELINE config 2
SEED_X = 42
FALLBACK_PEAK_GBS = 6650.0


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm

def build_config2_host(rows_lo, rows_hi, grid_y):
    """Config-2 band as host ELL arrays (ell_from_csr semantics) + x, generated on the CPU."""
    import numpy as np
    import torch
    from _load_pkg import load_pkg
    load_pkg()
    import gpu_spmv_b200.gen as gen
    rp, ci, va = gen.laplacian_band_csr(GRID, rows_lo, rows_hi, grid_y, "cpu")
    x = gen.vector_pm1(GRID * grid_y, SEED_X, "cpu").numpy()
    return rp.numpy(), ci.numpy(), va.numpy(), x


def cpu_reference_time(rp, ci, va, x, rows, cols, reps):
    """Seconds per spmv_cpu_ell pass through the unmodified reference (oracle/_ref), else the
    oracle port.  Returns (seconds, kind, threads)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle, Ref
    if Ref.available():
        ref = Ref()
        h, keep = ref.csr_wrap(rows, cols, rp, ci, va)
        he, st = ref.ell_from_csr(h)
        assert st == 0
        y = np.empty(rows, np.float32)
        fp = C.POINTER(C.c_float)
        ref.L.ref_time_spmv_cpu_ell(he, x.ctypes.data_as(fp), y.ctypes.data_as(fp), 1)  # warm-up
        sec = ref.L.ref_time_spmv_cpu_ell(he, x.ctypes.data_as(fp), y.ctypes.data_as(fp), reps)
        ref.L.ref_ell_destroy(he)
        ref.L.ref_csr_destroy(h)
        return sec, "reference", 1
    orc = Oracle()
    w, ec, ev = orc.ell_from_csr(rows, rp, ci, va)
    orc.spmv_ell(rows, w, ec, ev, x)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.spmv_ell(rows, w, ec, ev, x)
    return (time.perf_counter() - t0) / reps, "port", 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_full = GRID * GRID
    # bounded sample: a probe band sizes the per-step sample so that steps + warm-up stay near 2 minutes
    probe_rows = 1 << 20
    rp, ci, va, x = build_config2_host(0, probe_rows, GRID)
    probe_sec, _, _ = cpu_reference_time(rp, ci, va, x, probe_rows, n_full, 2)
    budget_rows = int(120.0 / max(args.steps + args.warmup, 1) / (probe_sec / probe_rows))
    n = max(probe_rows, min(n_full, budget_rows // GRID * GRID))
    rp, ci, va, x = build_config2_host(0, n, GRID)
    bytes_step = 8 * n * 5 + 4 * (n_full if n == n_full else n + GRID) + 4 * n
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_time(rp, ci, va, x, n, n_full, 1)
    sec, kind, threads = cpu_reference_time(rp, ci, va, x, n, n_full, args.steps)
    gbs = bytes_step / sec / 1e9
    sample = ("the full config-2 matrix" if n == n_full else f"the first {n} rows (a band) of the config-2 matrix")
    line = {
        "impl": "reference", "metric": "spmv_effective_hbm_gbs", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 5-point Laplacian 4096x4096 grid, ELL width 5, spmv_cpu_ell on the host",
                   "rows": n, "nnz": 5 * n - 4 * GRID, "bytes_per_step": bytes_step},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": kind,
                         "sample": f"{sample}, {args.steps} passes of spmv_cpu_ell per run, one pass per step "
                                   "(single-threaded: the reference has no threading)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- product arm

def timed_region(torch, stream, fn, steps, dist_on):
    """K calls of fn on `stream` between two CUDA events, barrier + sync on both sides."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        start.record(stream)
        for _ in range(steps):
            fn()
        stop.record(stream)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = start.elapsed_time(stop)
    if dist_on:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_kernel(torch, sp, stream, launch, bytes_per_launch, steps, warmup, dist_on=False, world=1):
    torch.cuda.synchronize()  # inputs were produced on torch's default stream; `stream` does not wait for it
    for _ in range(max(warmup, 3)):
        launch()
    torch.cuda.synchronize()
    ms = timed_region(torch, stream, launch, steps, dist_on)
    per = ms / steps
    return {"ms_per_step": per, "gbs": world * bytes_per_launch / (per * 1e-3) / 1e9}


def relabel(v, scale):
    """Bijective pseudo-random relabelling of vertex ids in [0, 2^scale) (xorshift, odd multiply,
    xorshift), the role of Graph500's vertex permutation: it removes R-MAT's id/degree correlation."""
    mask = (1 << scale) - 1
    v = v ^ (v >> (scale // 2))
    v = (v * 0x9E3779B1 + 0x7F4A7C15) & mask
    return v ^ (v >> (scale // 2 + 1))


def build_rmat_shard(torch, gen, scale, edge_factor, seed, rank, world, device, chunk=1 << 24, row_weight=1,
                     relabelled=False):
    """This rank's row shard of the column-normalised R-MAT matrix without materialising the
    whole graph: pass 1 counts in/out degrees of every edge (bincount), pass 2 keeps the edges
    whose destination falls in this rank's nnz-balanced row range and sorts only those."""
    n = 1 << scale
    n_edges = edge_factor << scale
    indeg = torch.zeros(n, dtype=torch.int64, device=device)
    outdeg = torch.zeros(n, dtype=torch.int64, device=device)
    ta = int(gen.RMAT_A * 4294967296.0)
    tb = int((gen.RMAT_A + gen.RMAT_B) * 4294967296.0)
    tc = int((gen.RMAT_A + gen.RMAT_B + gen.RMAT_C) * 4294967296.0)

    def edges(lo, hi):
        e = torch.arange(lo, hi, dtype=torch.int64, device=device)
        s = torch.zeros_like(e)
        d = torch.zeros_like(e)
        for level in range(scale):
            u = gen.hash32(seed, e, stream=16 + level)
            s = (s << 1) | (u >= tb).to(torch.int64)
            d = (d << 1) | (((u >= ta) & (u < tb)) | (u >= tc)).to(torch.int64)
        if relabelled:
            s, d = relabel(s, scale), relabel(d, scale)
        return s, d

    for lo in range(0, n_edges, chunk):
        s, d = edges(lo, min(lo + chunk, n_edges))
        indeg += torch.bincount(d, minlength=n)
        outdeg += torch.bincount(s, minlength=n)
    row_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(indeg, dim=0)
    import gpu_spmv_b200.dist as D
    bounds = D.partition_rows(row_ptrs, world, row_weight=row_weight)  # work(row) = nnz + row_weight
    r_lo, r_hi = bounds[rank], bounds[rank + 1]
    keys = []
    for lo in range(0, n_edges, chunk):
        s, d = edges(lo, min(lo + chunk, n_edges))
        m = (d >= r_lo) & (d < r_hi)
        keys.append(((d[m] << 32) | s[m]))
    key = torch.cat(keys) if keys else torch.zeros(0, dtype=torch.int64, device=device)
    del keys
    key, _ = torch.sort(key)
    cols = (key & 0xFFFFFFFF)
    vals = torch.ones((), dtype=torch.float32, device=device) / outdeg.to(torch.float32)[cols]
    rp_local = (row_ptrs[r_lo:r_hi + 1] - row_ptrs[r_lo]).to(torch.int32)
    return n, bounds, rp_local, cols.to(torch.int32), vals, n_edges


def run_product_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from _load_pkg import load_pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sp = load_pkg()
    import gpu_spmv_b200.dist as D
    import gpu_spmv_b200.gen as gen

    peak, peak_src = measured_peak()
    stream = torch.cuda.Stream()
    s_ptr = stream.cuda_stream
    extra = {}

    def log(msg):
        if rank == 0:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    # ---- headline: config 2, ELL, one 16.7 M-row band per rank (weak scaling) -----------------
    log("config 2: build + ELL headline")
    n_loc = GRID * GRID
    grid_y = GRID * world
    n_cols = GRID * grid_y
    rp, ci, va = gen.laplacian_band_csr(GRID, rank * n_loc, (rank + 1) * n_loc, grid_y, dev)
    A = sp.DeviceCSR(n_loc, n_cols, rp, ci, va)
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.ptr) == 0 and E.contents.max_nnz_per_row == 5
    x = gen.vector_pm1(n_cols, SEED_X, dev)
    y = torch.empty(n_loc, dtype=torch.float32, device=dev)
    x_touched = n_loc + 2 * GRID if world > 1 else n_cols   # a band only reads its rows +- one grid line of x
    ell_bytes = 8 * n_loc * 5 + 4 * x_touched + 4 * n_loc
    csr_bytes = 8 * ci.numel() + 4 * (n_loc + 1) + 4 * x_touched + 4 * n_loc

    def ell_step():
        rc = sp.lib.spmv_b200_spmv_ell_async(E, sp.dptr(x), sp.dptr(y), C.c_void_p(s_ptr))
        assert rc == 0

    torch.cuda.synchronize()  # x / matrix were produced on the default stream; `stream` does not wait for it
    for _ in range(max(args.warmup, 3)):
        ell_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches_before = sp.launch_count()
    ms = timed_region(torch, stream, ell_step, args.steps, dist_on)
    gpu_launches = sp.launch_count() - launches_before
    # keep the GPU under the same load a little longer so the clock sampler sees it
    if args.steps * (ms / args.steps) < 400:
        timed_region(torch, stream, ell_step, int(400 / max(ms / args.steps, 1e-3)) + 1, dist_on)
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    value = world * ell_bytes / (ms_per_step * 1e-3) / 1e9
    launch_gbs = ell_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: blocking C-ABI call with host x / y (pinned), copies inside the timed region -----
    log("e2e")
    x_host = x.cpu().pin_memory()
    y_host = torch.empty(n_loc, dtype=torch.float32).pin_memory()
    res = sp.SpMVResult()

    def e2e_step():
        x.copy_(x_host, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        rc = sp.lib.spmv_b200_spmv_ell(E, sp.dptr(x), sp.dptr(y), None, n_cols, C.byref(res))
        assert rc == 0
        y_host.copy_(y, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    if dist_on:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e = {"value": world * ell_bytes / e2e_sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n_cols,
           "d2h_bytes_per_step": 4 * n_loc, "ms_per_step": e2e_sec * 1e3,
           "api": "spmv_b200_spmv_ell (blocking C ABI) + pinned H2D of x + D2H of y"}

    # ---- CPU baseline (rank 0, N = 1): the reference's spmv_cpu_ell on the same matrix ---------
    log("cpu baseline")
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        reps = 3
        sec, kind, threads = cpu_reference_time(rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy(),
                                                x_host.numpy(), n_loc, n_cols, reps)
        cpu_baseline = {"value": ell_bytes / sec / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                        "ms_per_step": sec * 1e3, "host_cpus": os.cpu_count(),
                        "sample": f"the full config-2 matrix, {reps} passes of spmv_cpu_ell after 1 warm-up "
                                  "(single-threaded: the reference has no threading)"}

    # ---- extra: the CSR kernels on config 2 -----------------------------------------------------
    log("config 2: CSR kernels")
    if not args.quick and not args.only_pagerank:
        for kernel, name in ((sp.VECTOR_CSR, "csr_vector"), (sp.SCALAR_CSR, "csr_scalar"), (sp.MERGE_PATH, "csr_merge")):
            cfg = sp.make_config(kernel)
            r = bench_kernel(torch, sp, stream, lambda: sp.lib.spmv_b200_spmv_csr_async(
                A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), C.c_void_p(s_ptr)), csr_bytes, args.steps, args.warmup,
                dist_on, world)
            extra[f"config2_{name}"] = {"gbs": r["gbs"], "ms": r["ms_per_step"], "frac_of_measured_peak": r["gbs"] / world / peak,
                                        "frac_of_8000": r["gbs"] / world / 8000.0, "bytes": csr_bytes}
        # the same matrix through a CSR plan: no hub worth a table, >= 4 non-zeros per row -> the
        # segmented-stream kernel (csr_seg_kernels.cu)
        plan2 = sp.CsrPlan(A.ptr)
        r = bench_kernel(torch, sp, stream, lambda: plan2.spmv(x, y, s_ptr), csr_bytes, args.steps, args.warmup, dist_on, world)
        extra["config2_csr_planned"] = {"gbs": r["gbs"], "ms": r["ms_per_step"], "frac_of_measured_peak": r["gbs"] / world / peak,
                                        "frac_of_8000": r["gbs"] / world / 8000.0, "bytes": csr_bytes, "plan_mode": plan2.info()[2]}
        plan2.close()
        # and with the caller's permission to snapshot the values: the plan re-lays the matrix out as ELL
        plan3 = sp.CsrPlan(A.ptr, snapshot_values=True)
        r = bench_kernel(torch, sp, stream, lambda: plan3.spmv(x, y, s_ptr), csr_bytes, args.steps, args.warmup, dist_on, world)
        extra["config2_csr_planned_value_snapshot"] = {
            "gbs": r["gbs"], "ms": r["ms_per_step"], "frac_of_measured_peak": r["gbs"] / world / peak,
            "frac_of_8000": r["gbs"] / world / 8000.0, "bytes": csr_bytes, "plan_mode": plan3.info()[2],
            "note": "CSR algorithmic bytes over the time of the ELL kernel the plan routes to (it moves 805 MB)"}
        plan3.close()
    sp.ell_destroy(E)
    del A, rp, ci, va, x, y, x_host, y_host
    torch.cuda.empty_cache()

    if not args.quick and not args.only_pagerank and world == 1:
        # ---- config 3: short rows + 4 outlier rows of 1 M nnz: scalar (reference policy) vs merge ---
        log("config 3")
        rows3 = args.c3_rows
        rp3, ci3, va3 = gen.short_rows_with_outliers_csr(rows3, 43, dev)
        A3 = sp.DeviceCSR(rows3, rows3, rp3, ci3, va3)
        x3 = gen.uniform_01_open_low(5, torch.arange(rows3, device=dev), 9)
        y3 = torch.empty(rows3, dtype=torch.float32, device=dev)
        b3 = sp.csr_bytes(rows3, rows3, ci3.numel())
        for kernel, name in ((sp.SCALAR_CSR, "csr_scalar"), (sp.MERGE_PATH, "csr_merge")):
            cfg = sp.make_config(kernel)
            r = bench_kernel(torch, sp, stream, lambda: sp.lib.spmv_b200_spmv_csr_async(
                A3.ptr, sp.dptr(x3), sp.dptr(y3), C.byref(cfg), C.c_void_p(s_ptr)), b3, max(5, args.steps // 2), 3)
            extra[f"config3_{name}"] = {"gbs": r["gbs"], "ms": r["ms_per_step"], "frac_of_measured_peak": r["gbs"] / peak,
                                        "frac_of_8000": r["gbs"] / 8000.0, "bytes": b3, "rows": rows3, "nnz": ci3.numel()}
        extra["config3_selector"] = "MERGE_PATH (outlier override; reference policy: SCALAR_CSR)"
        del A3, rp3, ci3, va3, x3, y3
        torch.cuda.empty_cache()

    # ---- config 4 / 5: R-MAT SpMV (merge-path) and PageRank, row-sharded over the ranks -----------
    def rmat_section(scale, seed, do_vector, do_pagerank, relabelled=False):
        log(f"R-MAT scale {scale}: build" + (" (relabelled vertices)" if relabelled else ""))
        tag = "_relabelled" if relabelled else ""
        # SpMV shards balance the merge items (rows + nnz); a PageRank shard also updates and sends
        # every owned row, so rows weigh more there (8 GPUs, multicast exchange: weight 1 / 4 / 8 ->
        # 995 / 1063 / 920 iter/s; with unicast peer stores 8 was best, profiles/)
        weight = args.row_weight if args.row_weight >= 0 else (1 if not do_pagerank or world == 1 or relabelled else 4)
        n, bounds, srp, sci, sva, n_edges = build_rmat_shard(torch, gen, scale, 16, seed, rank, world, dev,
                                                            row_weight=weight, relabelled=relabelled)
        torch.cuda.synchronize()
        shard = D.CudaShard(n, bounds[rank], srp, sci, sva, stream=s_ptr)
        xg = torch.full((n,), 1.0 / n, dtype=torch.float32, device=dev)
        yg = torch.empty(n, dtype=torch.float32, device=dev)
        rows_p = bounds[rank + 1] - bounds[rank]
        b4 = 8 * sci.numel() + 4 * (rows_p + 1) + 4 * n + 4 * rows_p
        tot4 = torch.tensor([float(b4)], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(tot4)
        kernels = [(sp.MERGE_PATH, "csr_merge")] + ([(sp.VECTOR_CSR, "csr_vector")] if do_vector else [])
        for kernel, name in kernels:
            log(f"R-MAT scale {scale}: {name}")
            r = bench_kernel(torch, sp, stream, lambda: shard.spmv(xg, yg, kernel), 0, max(5, args.steps // 2), 3, dist_on)
            gbs = float(tot4.item()) / (r["ms_per_step"] * 1e-3) / 1e9
            extra[f"rmat{scale}{tag}_{name}"] = {"gbs": gbs, "ms": r["ms_per_step"], "frac_of_measured_peak": gbs / world / peak,
                                            "frac_of_8000": gbs / world / 8000.0, "bytes_all_ranks": float(tot4.item()),
                                            "nnz": n_edges, "rows": n, "scaling": "strong (one graph, row shards)"}
        # the same product through a CSR plan: merge coordinates computed once + hub-column table
        # (csr_hot_kernels.cu) for scale-free shards; bit-identical results, checked here on the full-size shard
        log(f"R-MAT scale {scale}: csr_merge through a CSR plan")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan = sp.CsrPlan(shard.csr.ptr)
        torch.cuda.synchronize()
        plan_ms = (time.perf_counter() - t0) * 1e3
        hot_columns, hot_nnz, hot_mode = plan.info()
        y_slice = yg[bounds[rank]:bounds[rank] + rows_p]
        y_ref = y_slice.clone()  # left by the plain merge-path run above (or the vector kernel)
        shard.spmv(xg, yg, sp.MERGE_PATH)
        torch.cuda.synchronize()
        y_ref.copy_(y_slice)
        r = bench_kernel(torch, sp, stream, lambda: plan.spmv(xg, y_slice, s_ptr), 0, max(5, args.steps // 2), 3, dist_on)
        torch.cuda.synchronize()
        same = bool(torch.equal(y_ref.view(torch.int32), y_slice.view(torch.int32)))
        gbs = float(tot4.item()) / (r["ms_per_step"] * 1e-3) / 1e9
        extra[f"rmat{scale}{tag}_csr_merge_hub_plan"] = {
            "gbs": gbs, "ms": r["ms_per_step"], "frac_of_measured_peak": gbs / world / peak, "frac_of_8000": gbs / world / 8000.0,
            "bytes_all_ranks": float(tot4.item()), "hub_columns_rank0": hot_columns,
            "hub_nnz_fraction_rank0": hot_nnz / max(sci.numel(), 1), "plan_mode": hot_mode, "plan_build_ms": plan_ms,
            "bit_identical_to_csr_merge": same, "scaling": "strong (one graph, row shards)"}
        plan.close()
        del y_ref
        if do_pagerank:
            hub_cols = shard.set_hot(-1)
            # fixed number of iterations of the sharded loop (stop rule evaluated every iteration,
            # one iteration late), wall clock around the loop after a device sync, max over ranks
            shard.damping = 0.85
            modes = (["multicast", "p2p"] if relabelled else ["multicast", "p2p", "nccl"]) if dist_on else ["single"]
            with torch.cuda.stream(stream):
                shard.setup_dangling()
            have_multicast = False
            for mode in modes:
                log(f"R-MAT scale {scale}: PageRank ({mode})")
                with torch.cuda.stream(stream):
                    if mode == "multicast":
                        pair = shard.enable_multicast_exchange()
                        if pair is None:
                            log("  no NVSwitch multicast on this box / torch build: " + getattr(shard, "_multicast_error", "multicast_ptr == 0"))
                            continue
                        have_multicast = True
                        r_a, r_b = pair
                    elif mode == "p2p":
                        r_a, r_b = shard.enable_peer_exchange()
                    else:
                        r_a = torch.empty(n, dtype=torch.float32, device=dev)
                        r_b = torch.empty_like(r_a)
                    partial = torch.zeros(3, dtype=torch.float64, device=dev)
                    shard.init_vector(r_a)
                    D.pagerank_loop(shard, r_a, r_b, partial, bounds, 0.85, 0.0, 0, fixed_iterations=3)  # warm-up
                    shard.init_vector(r_a)
                    if dist_on:
                        dist.barrier()
                    torch.cuda.synchronize()
                    iters = args.pr_iters
                    t0 = time.perf_counter()
                    fin, done, residual, conv, l1 = D.pagerank_loop(shard, r_a, r_b, partial, bounds, 0.85, 0.0, 0,
                                                                    fixed_iterations=iters)
                    torch.cuda.synchronize()
                    sec = time.perf_counter() - t0
                if dist_on:
                    t = torch.tensor([sec], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    sec = float(t.item())
                    dist.barrier()
                if mode == "p2p":
                    shard.disable_peer_exchange()
                if mode == "multicast":
                    shard.disable_multicast_exchange()
                it_bytes = float(tot4.item())
                exchange = {"multicast": "fused into the step kernel: ONE NVSwitch-multicast store (multimem.st) per finished "
                                         "row value, delivered to all GPUs by the switch (torch symmetric memory) "
                                         "+ NCCL all-reduce of 3 f64",
                            "p2p": "fused into the step kernel: peer stores of finished rows over NVLink (CUDA IPC) "
                                   "+ NCCL all-reduce of 3 f64",
                            "nccl": "NCCL all-gather of the rank slices + all-reduce of 3 f64",
                            "single": "none (1 GPU)"}[mode]
                key = {"multicast": "pagerank", "single": "pagerank", "nccl": "pagerank_nccl_allgather",
                       "p2p": "pagerank_p2p_unicast" if have_multicast else "pagerank"}[mode]
                extra[key + tag] = {
                    "iters_per_s": iters / sec, "ms_per_iter": sec / iters * 1e3,
                    "graph": f"R-MAT scale {scale} x16, d=0.85" + (", vertex ids relabelled (Graph500-style)" if relabelled
                                                                   else ", no vertex permutation"), "n": n, "nnz": n_edges, "iterations_timed": iters,
                    "l2_residual_after": residual, "effective_gbs": it_bytes / (sec / iters) / 1e9, "scaling": "strong",
                    "frac_of_measured_peak": it_bytes / (sec / iters) / 1e9 / world / peak, "exchange": exchange,
                    "hub_columns_rank0": hub_cols, "partition": f"work(row) = nnz + {weight}", "rows_per_rank_max": int(max(bounds[i + 1] - bounds[i] for i in range(world))),
                    "includes": "fused step + exchange + lagged host read of the residual every iteration; "
                                "excluded: setup + final normalisation"}
                del r_a, r_b
        shard.close()
        del shard, srp, sci, sva, xg, yg
        torch.cuda.empty_cache()

    if not args.quick and not args.only_pagerank:
        rmat_section(args.rmat_scale, 44, do_vector=(world == 1), do_pagerank=(args.pr_scale == args.rmat_scale))
        if args.pr_scale != args.rmat_scale:
            rmat_section(args.pr_scale, 45, do_vector=False, do_pagerank=True)
    if args.only_pagerank:
        rmat_section(args.pr_scale, 45, do_vector=False, do_pagerank=True)
    if args.relabelled and (not args.quick or args.only_pagerank):
        rmat_section(args.pr_scale, 45, do_vector=False, do_pagerank=True, relabelled=True)

    if rank == 0:
        roofline = {"bound": "hbm", "achieved": launch_gbs, "peak": peak, "unit": "GB/s", "frac": launch_gbs / peak,
                    "peak_source": peak_src, "frac_of_8000_nominal": launch_gbs / 8000.0,
                    "kernel": "ell_tma_pipe_kernel<1>", "algorithmic_bytes_per_launch": ell_bytes,
                    # dram__bytes_read.sum + dram__bytes_write.sum of ell_tma_pipe_kernel<1> on this workload, one
                    # `ncu --set full` capture (profiles/r1_ell_tma_pipe_c2_raw.csv: 738.8 + 61.9 MB; y is
                    # partly still in L2 at kernel end, hence slightly below the 805.3 MB algorithmic)
                    "traffic": args.ncu_traffic if args.ncu_traffic is not None else (800728576.0 if world == 1 else None)}
        line = {
            "metric": "spmv_effective_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: 5-point Laplacian 4096x4096 grid per GPU (16.7M rows, 83.9M nnz), ELL width 5, "
                                   "spmv_ell; x = hash U[-1,1)", "rows_per_gpu": n_loc, "bytes_per_step_per_gpu": ell_bytes,
                       "l2": "inputs (805 MB/GPU) larger than L2, no flush", "parallelism": f"row shards x{world}, x replicated"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline + e2e + cpu baseline only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rmat-scale", type=int, default=24, help="BASELINE config 4: R-MAT SpMV")
    ap.add_argument("--pr-scale", type=int, default=26, help="BASELINE config 5: PageRank graph")
    ap.add_argument("--c3-rows", type=int, default=50_000_000)
    ap.add_argument("--pr-iters", type=int, default=20)
    ap.add_argument("--row-weight", type=int, default=-1, help="partition work(row) = nnz + row_weight (-1: default policy)")
    ap.add_argument("--only-pagerank", action="store_true", help="tuning aid: headline + PageRank section only")
    ap.add_argument("--relabelled", type=int, default=1,
                    help="also run PageRank on the same R-MAT graph with relabelled vertex ids (extra.pagerank_relabelled)")
    ap.add_argument("--ncu-traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from an ncu capture (default: the committed one, "
                         "N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_product_arm(args)


if __name__ == "__main__":
    main()
