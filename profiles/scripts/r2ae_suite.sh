#!/bin/bash
# round 2, call AE: the whole GPU test suite at HEAD with the session's keep-alive context (per-test durations), smoke
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > $OUT/r2ae_gpu_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2ae_gpu_tests.log
grep -E "passed|failed|rc=" $OUT/r2ae_gpu_tests.log | tail -4
timeout 300 python __graft_entry__.py smoke > $OUT/r2ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2ae_smoke.log
