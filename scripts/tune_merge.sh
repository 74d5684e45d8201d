#!/bin/bash
# A/B timing of the merge-path tile kernel options (run on the GPU box)
for cfg in "0 -1" "0 50" "0 40" "1 -1" "1 66" "1 50" "1 44"; do set -- $cfg
  for t in merge_c2 merge_rmat24 pagerank_rmat24 merge_c3; do
    echo "tma=$1 carveout=$2 $(SPMV_B200_MERGE_TMA=$1 SPMV_B200_MERGE_CARVEOUT=$2 python scripts/profile_target.py $t 20 | tail -1)"
  done
done
