#!/bin/bash
# round 2, call AF: the tests whose parameters were thinned, then the bench lines of configs 2-4 at the final HEAD
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_baseline_configs.py tests/test_regions_ops.py tests/test_cli_parity.py -m gpu -x -q -k "first_10M or test_link or (random_files and density) or (subset_overlap and not sam)" --durations=8 > $OUT/r2af_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2af_tests.log
tail -14 $OUT/r2af_tests.log
timeout 400 python bench.py --config 2 --steps 5 > $OUT/r2af_bench2.json 2> $OUT/r2af_bench2.err; echo "bench2 rc=$?"
timeout 400 python bench.py --config 3 --steps 5 > $OUT/r2af_bench3.json 2> $OUT/r2af_bench3.err; echo "bench3 rc=$?"
timeout 400 python bench.py --config 4 --steps 3 > $OUT/r2af_bench4_n1.json 2> $OUT/r2af_bench4.err; echo "bench4 rc=$?"
for f in $OUT/r2af_bench*.json; do echo "== $f"; cut -c1-330 $f; done
tail -n 3 $OUT/r2af_*.err
