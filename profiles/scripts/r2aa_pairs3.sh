#!/bin/bash
# round 2, call AA: read pairs without offsets under count without -gaps (one pass + enumeration of the exceptions), skewed pairs
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_baseline_configs.py tests/test_gpu_parity.py tests/test_mgpu.py -m gpu -x -q -k "pairs or config3 or direct or golden or random or multi" > $OUT/r2aa_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2aa_tests.log
tail -8 $OUT/r2aa_tests.log
timeout 400 python profiles/scripts/time_paired_count.py > $OUT/r2aa_paired_count.json 2> $OUT/r2aa_paired_count.err; cat $OUT/r2aa_paired_count.json; tail -n 3 $OUT/r2aa_paired_count.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2aa_bench1.json 2> $OUT/r2aa_bench1.err; cut -c1-330 $OUT/r2aa_bench1.json
