#!/bin/bash
# The ncu recipe behind profiles/ (run on the GPU box through gpurun; ONE GPU, never a multi-rank command):
#
#   gpurun --timeout 900 -- 'scripts/ncu_capture.sh vector_c2 csr_short'        # config 2 through the drop-in spmv_csr
#   gpurun --timeout 900 -- 'scripts/ncu_capture.sh ell_c2 ell_tma_pipe'        # the headline ELL kernel
#   gpurun --timeout 900 -- 'scripts/ncu_capture.sh hot0_rmat24 merge_hot'      # the hub-column kernel on R-MAT 24
#   gpurun --timeout 900 -- 'scripts/ncu_capture.sh pagerank_rmat24 merge_hot'  # the fused PageRank step
#
# 1. the target runs once WITHOUT ncu and must exit 0 (its CUDA-event time is printed; a time measured under
#    ncu is never a bench value);
# 2. launch list of the same command (per-launch durations, cold caches, serialised: compare SHARES);
# 3. one `--set full` capture of the named kernel after 3 warm-up launches, source-correlated (-lineinfo);
# 4. the raw page as CSV next to it -- copy what you want judged into profiles/ and rebuild
#    profiles/ncu_traffic.json with scripts/ncu_traffic.py.
# NVTX: every entry point of the library pushes a range (spmv_b200:spmv_csr, :spmv_ell, :spmv_csr_planned,
# :pagerank_device, :pagerank_multi.run, :spmv_ell_host); add `--nvtx --nvtx-include "spmv_b200:pagerank_device/"`
# to restrict a capture to the kernels of one call.
set -e
TARGET="${1:?target of scripts/profile_target.py}"
KERNEL="${2:?kernel name regex}"
OUT="gpurun_out/ncu_${TARGET}"
mkdir -p gpurun_out
python scripts/profile_target.py "$TARGET" 5 | tee "${OUT}_plain.log"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "${OUT}_launches.csv" \
    python scripts/profile_target.py "$TARGET" 5 > "${OUT}_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k "regex:${KERNEL}" -s 3 -c 1 -f -o "${OUT}" \
    python scripts/profile_target.py "$TARGET" 5 > "${OUT}_full.log" 2>&1
ncu -i "${OUT}.ncu-rep" --page raw --csv > "${OUT}_raw.csv"
grep -c . "${OUT}_raw.csv"
