"""Timing aid: SCALAR / VECTOR / MERGE on mid-size uniform random matrices (rows 4 M, avg 3 .. 24 nnz per row,
columns uniform over 4 M): checks that the short-row ring (avg <= 16) is not a regression against the general ring.
    python scripts/time_mid.py            (runs itself with SPMV_B200_CSR_SHORT=0 as well)"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    rows = 1 << 22
    for avg in (3, 5, 8, 12, 16, 24):
        rp, ci, va = gen.random_csr(rows, rows, avg, seed=avg, device=dev)
        x = gen.vector_pm1(rows, 3, dev)
        A = sp.DeviceCSR(rows, rows, rp, ci, va)
        y = torch.empty(rows, device=dev)
        nbytes = sp.csr_bytes(rows, rows, ci.numel())
        longest = int((rp[1:] - rp[:-1]).max().item())
        out = []
        for name, k in (("scalar", sp.SCALAR_CSR), ("vector", sp.VECTOR_CSR), ("merge", sp.MERGE_PATH)):
            cfg = sp.make_config(k)
            run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), None)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                run()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 20
            out.append(f"{name} {ms:.4f} ms ({nbytes / ms / 1e6 / 6548.5:.2f})")
        print(f"  {os.environ.get('TAG', '')} avg {avg:2d} (nnz {ci.numel() / 1e6:.0f} M, longest {longest}): " + "  ".join(out), flush=True)
    # uniform rows: periodic stencils (every row has exactly len(offsets) entries, columns sorted)
    g = 160
    n = g * g * g
    stencils = {"3-D 7-point": [0, 1, -1, g, -g, g * g, -g * g],
                "2.5-D 9-point": [0, 1, -1, g, -g, g + 1, g - 1, -g + 1, -g - 1],
                "3-D 13-point": [0, 1, -1, 2, -2, g, -g, 2 * g, -2 * g, g * g, -g * g, 2 * g * g, -2 * g * g]}
    for name_s, offs in stencils.items():
        i = torch.arange(n, dtype=torch.int64, device=dev)
        cols = torch.stack([(i + o) % n for o in offs], dim=1)
        cols, _ = torch.sort(cols, dim=1)
        w = len(offs)
        rp = (torch.arange(n + 1, dtype=torch.int64, device=dev) * w).to(torch.int32)
        ci = cols.reshape(-1).to(torch.int32).contiguous()
        va = gen.uniform_pm1(9, torch.arange(n * w, dtype=torch.int64, device=dev), 3)
        x = gen.vector_pm1(n, 3, dev)
        A = sp.DeviceCSR(n, n, rp, ci, va)
        y = torch.empty(n, device=dev)
        nbytes = sp.csr_bytes(n, n, ci.numel())
        out = []
        for name, k in (("scalar", sp.SCALAR_CSR), ("vector", sp.VECTOR_CSR), ("merge", sp.MERGE_PATH)):
            cfg = sp.make_config(k)
            run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), None)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                run()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 20
            out.append(f"{name} {ms:.4f} ms ({nbytes / ms / 1e6 / 6548.5:.2f})")
        print(f"  {os.environ.get('TAG', '')} {name_s} (rows {n / 1e6:.1f} M, {w} per row): " + "  ".join(out), flush=True)
    sys.exit(0)
for tag, env in (("short ring ", {}), ("general    ", {"SPMV_B200_CSR_SHORT": "0"})):
    subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, TAG=tag, **env))
