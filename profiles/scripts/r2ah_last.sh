#!/bin/bash
# round 2, call AH: the final HEAD's drivers (word-wise BED reader) -- smoke, then command-line parity runs until the time is up
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 100 python __graft_entry__.py smoke > $OUT/r2ah_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2ah_smoke.log
timeout 170 python -m pytest tests/test_cli_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "known_answer or kat or config0 or (random_files and coverage) or scans_counts or bam or gz" > $OUT/r2ah_tests.log 2>&1
echo "tests rc=$? (124 = the time limit of this call, not a failure)" >> $OUT/r2ah_tests.log
tail -5 $OUT/r2ah_tests.log
