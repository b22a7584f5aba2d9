#!/bin/bash
# round 2, call T (8 GPUs): gtb_mgpu end to end from host memory at 1 / 2 / 4 / 8 devices in one process; tests with the real devices
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2t_box.txt
timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q > $OUT/r2t_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2t_tests.log
tail -3 $OUT/r2t_tests.log
timeout 600 python profiles/scripts/time_mgpu.py 800000000 > $OUT/r2t_mgpu.json 2> $OUT/r2t_mgpu.err; echo "rc=$?"; cat $OUT/r2t_mgpu.json; tail -n 3 $OUT/r2t_mgpu.err
