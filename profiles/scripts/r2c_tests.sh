#!/bin/bash
# round 2, call C: the whole GPU test suite (with the new BASELINE-config and drop-in tests), timings per test
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/r2c_nproc.txt; free -g >> $OUT/r2c_nproc.txt
timeout 1700 python -m pytest tests -m gpu -x -q --durations=25 > $OUT/r2c_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2c_tests.log
tail -45 $OUT/r2c_tests.log
