"""Target of the compute-sanitizer runs (scripts/sanitizer.sh -> profiles/r2_sanitizer.txt): one small pass of every
kernel family of the library -- CSR scalar / vector / merge-path, ELL (TMA ring), the planned kernels (hub-column,
segmented stream, ELL layout), the fused PageRank iteration (device loop, graph replay off and on), the sharded path
with two ranks on one GPU (peer stores), device assembly, top-k, and the gated host-buffer call."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from _load_pkg import load_pkg
sp = load_pkg()
import gpu_spmv_b200.gen as gen
dev = torch.device("cuda:0")
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = grid * grid
for name, (rp, ci, va) in (("laplacian", gen.laplacian_2d_csr(grid, dev)), ("random", gen.random_csr(n, n, 9, 3, dev, 0.02))):
    A = sp.DeviceCSR(n, n, rp, ci, va)
    x = gen.vector_pm1(n, 1, dev)
    y = torch.empty(n, device=dev)
    for k in (0, 1, 2):
        r = sp.spmv_csr(A.ptr, x, y, sp.make_config(k), n)
        torch.cuda.synchronize()
        print(name, sp.KERNEL_NAMES[k], "rc", r.error_code, flush=True)
    E = sp.ell_create(0, 0, 0)
    print("ell_from_csr_device", sp.ell_from_csr_device(E, A.ptr), E.contents.max_nnz_per_row, flush=True)
    r = sp.spmv_ell(E, x, y, None, n)
    print(name, "ELL rc", r.error_code, flush=True)
    # the gated host-buffer call (pinned y: stored by the kernel; pageable y: D2H chunks)
    plan = C.c_void_p()
    assert sp.lib.spmv_b200_ell_host_plan_create(E, 0, C.byref(plan)) == 0
    xh, yh = x.cpu().pin_memory(), torch.empty(n).pin_memory()
    yp = np.empty(n, np.float32)
    for _ in range(2):
        assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yh.data_ptr()) == 0
        assert sp.lib.spmv_b200_spmv_ell_host(plan, xh.data_ptr(), yp.ctypes.data) == 0
    gated = C.c_int()
    sp.lib.spmv_b200_ell_host_plan_gated(plan, C.byref(gated), None)
    print(name, "host call: gated", gated.value, "same", bool(torch.equal(yh, y.cpu())), bool(np.array_equal(yp, y.cpu().numpy())), flush=True)
    sp.lib.spmv_b200_ell_host_plan_destroy(plan)
    sp.ell_destroy(E)
    for force, snap in ((True, False), (True, True)):
        p = sp.CsrPlan(A.ptr, 64, force=force, snapshot_values=snap)
        print(name, "plan mode", p.info()[2], "rc", p.spmv(x, y), flush=True)
        torch.cuda.synchronize()
        p.close()

# PageRank: device loop on a small R-MAT graph, then two ranks sharing this GPU (peer stores)
m, prp, pci, pva = gen.rmat_pagerank_csr(11, 8, seed=9, device="cpu")
G = sp.csr_from_arrays(m, m, prp.numpy(), pci.numpy(), pva.numpy())
assert sp.csr_to_gpu(G) == 0
ranks = torch.empty(m, dtype=torch.float32, device=dev)
print("pagerank_device", sp.pagerank_device(G, ranks, sp.make_pagerank_config(0.85, 1e-6, 50))[:3], flush=True)
ids, vals = sp.pagerank_top_k_device(ranks, 8)
print("top-k", ids[:3], flush=True)
import gpu_spmv_b200.dist as D
cfg = sp.make_pagerank_config(0.85, 1e-6, 50)
out = np.empty(m, np.float32)
res = D.PrDistResult()
devices = (C.c_int * 2)(0, 0)
rc = sp.lib.spmv_b200_pagerank_multi(G, C.byref(cfg), 2, devices, 1, 4, 6, out.ctypes.data_as(C.c_void_p), C.byref(res))
print("pagerank_multi (2 ranks sharing the GPU, peer stores) rc", rc, "sum", float(out.sum()), flush=True)
print("done")
