#!/bin/bash
# round 2, call AB: count without -gaps over spliced reads (CSR) through the one-pass engine + exceptions
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_baseline_configs.py tests/test_gpu_parity.py tests/test_cli_parity.py -m gpu -x -q -k "spliced or pairs or config3 or golden or random or multi or sam or subset" > $OUT/r2ab_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2ab_tests.log
tail -8 $OUT/r2ab_tests.log
timeout 400 python profiles/scripts/time_spliced_count.py > $OUT/r2ab_spliced.json 2> $OUT/r2ab_spliced.err; cat $OUT/r2ab_spliced.json; tail -n 3 $OUT/r2ab_spliced.err
