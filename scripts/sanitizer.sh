#!/bin/bash
# compute-sanitizer over one small pass of every kernel family (scripts/sanitizer_target.py); SURVEY section 5 asks for it,
# the reference has none.  NOT RUN in this round: compute-sanitizer is closed on this GPU pool (the wrapper refuses to start
# it); kept runnable for any other box.  There:   gpurun --timeout 900 -- 'bash scripts/sanitizer.sh > gpurun_out/sanitizer.txt 2>&1'
# The gated host-buffer kernel POLLS device memory the copy engine is writing; under the sanitizer's serialisation the
# upload may complete before the kernel starts, which is fine (it then never waits).
set +e
export SPMV_B200_PR_GRAPH=${SPMV_B200_PR_GRAPH:-0}   # the sanitizer instruments launches, not graph replays
for tool in memcheck racecheck synccheck initcheck; do
    echo "=== compute-sanitizer --tool $tool"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python scripts/sanitizer_target.py 96 2>&1 \
        | grep -v "^$" | grep -E "rc|done|ERROR SUMMARY|=========.*(Error|error|Invalid|Race|Hazard|Uninit|Barrier)|same|pagerank|top-k" | head -60
    echo "exit code: ${PIPESTATUS[0]}"
done
