#!/bin/bash
# round 2, call U: the pre-lookup sorted-input path of the DIRECT engine -- tests, bench line (random input must not move), sorted input at 100 M and 1 B reads
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "direct or golden or random or config0 or config1 or full_size or sharded" > $OUT/r2u_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2u_tests.log
tail -4 $OUT/r2u_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2u_bench1.json 2> $OUT/r2u_bench1.err; echo "rc=$?"; cut -c1-330 $OUT/r2u_bench1.json
timeout 600 python bench_configs.py --only count,coverage,sorted > $OUT/r2u_configs_100m.jsonl 2> $OUT/r2u_configs_100m.err; echo "rc=$?"; cut -c1-260 $OUT/r2u_configs_100m.jsonl
timeout 600 python bench_configs.py --only count,sorted --reads 1000000000 > $OUT/r2u_configs_1b.jsonl 2> $OUT/r2u_configs_1b.err; echo "rc=$?"; cut -c1-260 $OUT/r2u_configs_1b.jsonl
tail -n 3 $OUT/r2u_configs_1b.err
