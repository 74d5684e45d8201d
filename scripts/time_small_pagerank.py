"""Wall time of spmv_b200_pagerank_device on a small graph (launch-bound): SPMV_B200_PR_GRAPH=0|1."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from _load_pkg import load_pkg  # noqa: E402

sp = load_pkg()
import gpu_spmv_b200.gen as gen  # noqa: E402

dev = torch.device("cuda:0")
n, rp, ci, va = gen.rmat_pagerank_csr(12, 8, 9, dev)
G = sp.DeviceCSR(n, n, rp, ci, va)
d_ranks = torch.empty(n, device=dev)
cfg = sp.make_pagerank_config(0.85, 0.0, 200)  # tolerance 0: always 200 iterations
for _ in range(3):
    sp.pagerank_device(G.ptr, d_ranks, cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    rc, iters, res, conv, l1 = sp.pagerank_device(G.ptr, d_ranks, cfg)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"SPMV_B200_PR_GRAPH={os.environ.get('SPMV_B200_PR_GRAPH', '1')}: {dt * 1e3:.3f} ms per call, {iters} iterations, "
      f"{dt / iters * 1e6:.2f} us per iteration (incl. set-up), checksum {float(d_ranks.double().sum()):.12f}")
