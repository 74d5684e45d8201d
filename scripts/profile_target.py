"""One kernel family on one BASELINE dataset, a few launches -- the command profiled by ncu.

    python scripts/profile_target.py <target> [launches]
    targets: ell_c2 vector_c2 scalar_c2 merge_c2 scalar_c3 merge_c3 merge_rmat<scale> vector_rmat<scale>
             pagerank_rmat<scale> hot<cap>_rmat<scale> (hub-column plan with <cap> table entries, 0 = max)
Prints the CUDA-event time per launch (never a bench value when run under a profiler).
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from _load_pkg import load_pkg  # noqa: E402

sp = load_pkg()
import gpu_spmv_b200.dist as D  # noqa: E402
import gpu_spmv_b200.gen as gen  # noqa: E402

target = sys.argv[1]
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
kernels = {"scalar": sp.SCALAR_CSR, "vector": sp.VECTOR_CSR, "merge": sp.MERGE_PATH}
what, data = target.split("_", 1)

if data == "c2":
    n = 4096 * 4096
    rp, ci, va = gen.laplacian_2d_csr(4096, dev)
    x = gen.vector_pm1(n, 42, dev)
elif data == "c3":
    n = 50_000_000
    rp, ci, va = gen.short_rows_with_outliers_csr(n, 43, dev)
    x = gen.uniform_01_open_low(5, torch.arange(n, device=dev), 9)
elif data.startswith("rmat"):
    n, rp, ci, va = gen.rmat_pagerank_csr(int(data[4:]), 16, 44, dev)
    x = torch.full((n,), 1.0 / n, device=dev)
else:
    raise SystemExit(f"unknown dataset {data}")
A = sp.DeviceCSR(n, n, rp, ci, va)
y = torch.empty(n, device=dev)
nbytes = sp.csr_bytes(n, n, ci.numel())
torch.cuda.synchronize()

if what == "ell":
    E = sp.ell_create(0, 0, 0)
    assert sp.ell_from_csr_device(E, A.ptr) == 0
    nbytes = sp.ell_bytes(n, n, E.contents.max_nnz_per_row)
    run = lambda: sp.lib.spmv_b200_spmv_ell_async(E, sp.dptr(x), sp.dptr(y), None)  # noqa: E731
elif what.startswith("hot"):
    plan = sp.CsrPlan(A.ptr, int(what[3:] or 0))
    print("plan (hot columns, hot nnz, mode):", plan.info())
    run = lambda: plan.spmv(x, y)  # noqa: E731
elif what == "pagerank":
    shard = D.CudaShard(n, 0, rp, ci, va)
    shard.setup_dangling()
    r_a, r_b = torch.empty(n, device=dev), torch.empty(n, device=dev)
    partial = torch.zeros(3, dtype=torch.float64, device=dev)
    shard.init_vector(r_a)
    run = lambda: shard(r_a, r_b, partial)  # noqa: E731
else:
    cfg = sp.make_config(kernels[what])
    run = lambda: sp.lib.spmv_b200_spmv_csr_async(A.ptr, sp.dptr(x), sp.dptr(y), C.byref(cfg), None)  # noqa: E731

for _ in range(3):
    run()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(launches):
    run()
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / launches
print(f"{target}: {ms:.4f} ms/launch, {nbytes / ms / 1e6:.1f} GB/s algorithmic ({nbytes} B)")
