#!/bin/bash
# A/B timing of the CSR row-owner kernels on config 2 (run on the GPU box)
run() { echo "$1 u=$2 rpg=$3 stages=$4 ctas=$5: $(SPMV_B200_CSR_U=$2 SPMV_B200_CSR_RPG=$3 SPMV_B200_CSR_STAGES=$4 SPMV_B200_CSR_CTAS_PER_SM=$5 python scripts/profile_target.py $1 30 | tail -1)"; }
for rpg in 2 4 6 8; do run vector_c2 4 $rpg 0 0; done
run vector_c2 4 4 3 0; run vector_c2 4 8 3 0; run vector_c2 6 4 0 0
