#!/usr/bin/env python
"""Host-buffer ELL call (spmv_b200_spmv_ell_host) on BASELINE config 2: the gated single-kernel form
against the chunked form, over the x chunk count / copy streams / lead.  One process per setting of
the process-wide knobs (they are read once).  Output -> profiles/r2_host_gated.txt.

    python scripts/host_gated_sweep.py            # the sweep (spawns children)
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    grid = int(os.environ.get("GRID", "4096"))
    n = grid * grid
    rp, ci, va = gen.laplacian_2d_csr(grid, dev)
    x = gen.vector_pm1(n, 42, dev)
    A = sp.DeviceCSR(n, n, rp, ci, va)
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.ptr) == 0
    y = torch.empty(n, device=dev)
    xh = x.cpu().pin_memory()
    yh = torch.empty(n).pin_memory()
    res = sp.SpMVResult()
    assert sp.lib.spmv_b200_spmv_ell(E, sp.dptr(x), sp.dptr(y), None, n, C.byref(res)) == 0
    y_ref = y.cpu()
    tag = os.environ.get("TAG", "")
    nocheck = os.environ.get("SWEEP_NOCHECK") == "1"
    for chunks in [int(c) for c in os.environ.get("SWEEP_CHUNKS", "32").split(",")]:
        if chunks:
            os.environ["SPMV_B200_HOST_GATED_CHUNKS"] = str(chunks)
        plan = C.c_void_p()
        assert sp.lib.spmv_b200_ell_host_plan_create(E, chunks if os.environ.get("SPMV_B200_HOST_GATED") == "0" else 0, C.byref(plan)) == 0
        gated, xc = C.c_int(), C.c_int()
        sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(gated), C.byref(xc))
        for _ in range(3):
            yh.fill_(float("nan"))
            assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
            assert nocheck or torch.equal(yh.view(torch.int32), y_ref.view(torch.int32)), "y differs"
        still = C.c_int()
        sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(still), None)
        best = 1e9
        tot = 0.0
        reps = 5
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr())
            ms = (time.perf_counter() - t0) / 10 * 1e3
            best = min(best, ms)
            tot += ms
        assert nocheck or torch.equal(yh.view(torch.int32), y_ref.view(torch.int32)), "y differs"
        print(f"{tag:34s} gated={gated.value}->{still.value} x_chunks={xc.value if gated.value else chunks:3d}  "
              f"mean {tot / reps:.3f} ms  best {best:.3f} ms  {(48 * n + 4) / (tot / reps) / 1e6:.0f} GB/s (alg.)", flush=True)
        sp.lib.spmv_b200_ell_host_plan_destroy(plan)
    sys.exit(0)

G = "SPMV_B200_HOST_GATED"
sets = [
    ("chunked (GATED=0), y by the kernel", {G: "0", "SWEEP_CHUNKS": "12"}),
    ("gated, y by the kernel (default)", {"SWEEP_CHUNKS": "0"}),
    ("gated, y in 8 D2H copies", {"SPMV_B200_HOST_ZERO_COPY_Y": "0", "SWEEP_CHUNKS": "8"}),
    ("gated, y by the kernel, poll 0", {G + "_POLL_NS": "0", "SWEEP_CHUNKS": "0"}),
    ("gated, launch-blocking", {"CUDA_LAUNCH_BLOCKING": "1", "SWEEP_CHUNKS": "0"}),
]
for tag, env in sets:
    e = dict(os.environ, TAG=tag, **env)
    subprocess.run([sys.executable, __file__, "child"], env=e, timeout=600)
