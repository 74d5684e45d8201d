"""Wall time of spmv_b200_pagerank_device on small graphs (launch-bound sizes): the one-kernel persistent loop
(pagerank_small.cu, SPMV_B200_PR_SMALL=2 forces it) against the multi-kernel loop with CUDA-graph replay
(SPMV_B200_PR_SMALL=0).  -> profiles/r2_small_pagerank.txt

    python scripts/time_small_pagerank.py            # both modes, R-MAT scales 8..18 (child processes)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    from _load_pkg import load_pkg
    sp = load_pkg()
    import gpu_spmv_b200.gen as gen
    dev = torch.device("cuda:0")
    for scale in [int(s) for s in sys.argv[2].split(",")]:
        n, rp, ci, va = gen.rmat_pagerank_csr(scale, 16, 9, dev)
        G = sp.DeviceCSR(n, n, rp, ci, va)
        d_ranks = torch.empty(n, device=dev)
        cfg = sp.make_pagerank_config(0.85, 0.0, 200)  # tolerance 0: always 200 iterations
        for _ in range(3):
            sp.pagerank_device(G.ptr, d_ranks, cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            rc, iters, res, conv, l1 = sp.pagerank_device(G.ptr, d_ranks, cfg)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        np.save(os.path.join(ROOT, "gpurun_out", f"small_pr_{os.environ.get('SPMV_B200_PR_SMALL', '1')}_{scale}.npy"), d_ranks.cpu().numpy())
        print(f"SPMV_B200_PR_SMALL={os.environ.get('SPMV_B200_PR_SMALL', '1')} scale {scale:2d} (n {n:7d}, nnz {int(ci.numel()):8d}): "
              f"{dt * 1e3:8.3f} ms per call, {iters} iterations, {dt / iters * 1e6:7.2f} us per iteration (incl. set-up), "
              f"residual {res:.3e}", flush=True)
    sys.exit(0)

import numpy as np
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
scales = "8,10,12,14,16,17,18"
for mode in ("0", "2"):
    subprocess.run([sys.executable, __file__, "child", scales], env=dict(os.environ, SPMV_B200_PR_SMALL=mode))
for scale in scales.split(","):
    a = np.load(os.path.join(ROOT, "gpurun_out", f"small_pr_0_{scale}.npy")).astype(np.float64)
    b = np.load(os.path.join(ROOT, "gpurun_out", f"small_pr_2_{scale}.npy")).astype(np.float64)
    print(f"scale {scale}: L1 distance between the two loops after 200 iterations {np.abs(a - b).sum():.3e}, sum {b.sum():.9f}")
