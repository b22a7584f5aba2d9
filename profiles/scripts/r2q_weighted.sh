#!/bin/bash
# round 2, call Q: weighted queries through the DIRECT engine
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "direct or weight or engine or config1 or label" > $OUT/r2q_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2q_tests.log
tail -5 $OUT/r2q_tests.log
timeout 300 python profiles/scripts/time_weighted.py > $OUT/r2q_weighted.json 2> $OUT/r2q_weighted.err; echo "rc=$?"; cat $OUT/r2q_weighted.json; tail -n 3 $OUT/r2q_weighted.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2q_bench1.json 2> $OUT/r2q_bench1.err; echo "rc=$?"; cut -c1-400 $OUT/r2q_bench1.json
