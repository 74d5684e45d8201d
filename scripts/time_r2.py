"""Round-2 timing aid (one GPU): config-2 CSR through SCALAR/VECTOR with and without the short-row
kernel, and the pipelined host-buffer ELL call against the serial form.  CUDA-event / wall times; never run under ncu."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    n = 4096 * 4096
    rp, ci, va = gen.laplacian_2d_csr(4096, dev)
    x = gen.vector_pm1(n, 42, dev)
    A = sp.DeviceCSR(n, n, rp, ci, va)
    y = torch.empty(n, device=dev)
    nbytes = sp.csr_bytes(n, n, ci.numel())
    for name, k in (("scalar", sp.SCALAR_CSR), ("vector", sp.VECTOR_CSR)):
        cfg = sp.make_config(k)
        run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), None)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(50):
            run()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 50
        print(f"  {os.environ.get('TAG','')} c2 csr {name}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s  frac {nbytes / ms / 1e6 / 6548.5:.3f}", flush=True)
    sys.exit(0)

env_sets = [("short-row ring (default)", {}), ("general ring (SHORT=0)", {"SPMV_B200_CSR_SHORT": "0"}),
            ("short 6 CTAs/SM", {"SPMV_B200_CSR_CTAS_PER_SM": "6"})]
if len(sys.argv) > 1 and sys.argv[1] == "csr":
    for tag, env in env_sets:
        e = dict(os.environ, TAG=tag, **env)
        subprocess.run([sys.executable, __file__, "child"], env=e)
    sys.exit(0)

# ---- host-buffer ELL: serial vs pipelined -------------------------------------------------------
import torch
from _load_pkg import load_pkg
sp = load_pkg()
import gpu_spmv_b200.gen as gen
dev = torch.device("cuda:0")
n = 4096 * 4096
rp, ci, va = gen.laplacian_2d_csr(4096, dev)
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
for chunks in (6, 8, 12, 16, 24, 32):
    plan = C.c_void_p()
    assert sp.lib.spmv_b200_ell_host_plan_create(E, chunks, C.byref(plan)) == 0
    for _ in range(3):
        assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
    assert torch.equal(yh, y_ref)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr())
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"ell host pipelined, {chunks} chunks: {ms:.3f} ms/step  {805306368 / ms / 1e6:.0f} GB/s", flush=True)
    sp.lib.spmv_b200_ell_host_plan_destroy(plan)


def serial():
    x.copy_(xh, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    sp.lib.spmv_b200_spmv_ell(E, sp.dptr(x), sp.dptr(y), None, n, C.byref(res))
    yh.copy_(y, non_blocking=True)
    torch.cuda.current_stream().synchronize()


for _ in range(3):
    serial()
t0 = time.perf_counter()
for _ in range(20):
    serial()
ms = (time.perf_counter() - t0) / 20 * 1e3
print(f"ell host serial (H2D, spmv_ell, D2H): {ms:.3f} ms/step")

# ---- floor of the host-buffer call: both PCIe directions busy at once, no compute --------------------------
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
for parts in (1, 8):
    step = n // parts

    def duplex():
        for p in range(parts):
            with torch.cuda.stream(s_up):
                x[p * step:(p + 1) * step].copy_(xh[p * step:(p + 1) * step], non_blocking=True)
            with torch.cuda.stream(s_down):
                yh[p * step:(p + 1) * step].copy_(y[p * step:(p + 1) * step], non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()

    for _ in range(3):
        duplex()
    t0 = time.perf_counter()
    for _ in range(20):
        duplex()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"duplex copies only (67 MB up + 67 MB down at once, {parts} part(s) each): {ms:.3f} ms  "
          f"({2 * 4 * n / ms / 1e6:.1f} GB/s both directions)")
for name, fn in (("H2D only", lambda: x.copy_(xh, non_blocking=True)), ("D2H only", lambda: yh.copy_(y, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{name}: {ms:.3f} ms  ({4 * n / ms / 1e6:.1f} GB/s)")
