#!/bin/bash
# round 2, call S (2 GPUs): gtb_mgpu tests (one process, several devices) and the end-to-end rate from host memory
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2s_box.txt
timeout 900 python -m pytest tests/test_mgpu.py -m gpu -x -q > $OUT/r2s_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2s_tests.log
tail -5 $OUT/r2s_tests.log
timeout 600 python profiles/scripts/time_mgpu.py 400000000 > $OUT/r2s_mgpu.json 2> $OUT/r2s_mgpu.err; echo "rc=$?"; cat $OUT/r2s_mgpu.json; tail -n 3 $OUT/r2s_mgpu.err
