#!/bin/bash
# A/B timing of the ELL kernel variants on config 2 (run on the GPU box)
run() { echo "variant=$1 stages=$2 ctas=$3: $(SPMV_B200_ELL_VARIANT=$1 SPMV_B200_ELL_STAGES=$2 SPMV_B200_ELL_CTAS_PER_SM=$3 python scripts/profile_target.py ell_c2 50 | tail -1)"; }
run 5 2 4; run 5 2 5; run 5 3 3; run 6 2 8; run 6 3 6; run 6 2 6; run 6 4 5
