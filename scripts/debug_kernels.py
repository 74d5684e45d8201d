"""Small driver for compute-sanitizer runs: every CSR kernel + ELL on a small Laplacian and a random matrix."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _load_pkg import load_pkg
sp = load_pkg()
import gpu_spmv_b200.gen as gen
dev = torch.device("cuda:0")
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = grid * grid
for name, (rp, ci, va) in (("laplacian", gen.laplacian_2d_csr(grid, dev)), ("random", gen.random_csr(n, n, 9, 3, dev, 0.02))):
    A = sp.DeviceCSR(n, n, rp, ci, va)
    x = gen.vector_pm1(n, 1, dev)
    y = torch.empty(n, device=dev)
    for k in (0, 1, 2):
        r = sp.spmv_csr(A.ptr, x, y, sp.make_config(k), n)
        torch.cuda.synchronize()
        print(name, sp.KERNEL_NAMES[k], "rc", r.error_code, "ms", r.elapsed_ms, flush=True)
    E = sp.ell_create(0, 0, 0)
    print("ell_from_csr_device", sp.ell_from_csr_device(E, A.ptr), E.contents.max_nnz_per_row, flush=True)
    r = sp.spmv_ell(E, x, y, None, n)
    print(name, "ELL rc", r.error_code, "ms", r.elapsed_ms, flush=True)
    sp.ell_destroy(E)
print("done")
