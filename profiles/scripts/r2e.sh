#!/bin/bash
# round 2, call E: fixes after call D (lazy index checks under -S, paired-read cluster trigger, packed host form), bench line, configs[3]
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cli_parity.py "tests/test_gpu_parity.py" "tests/test_baseline_configs.py::test_config3_paired_coverage_abi" "tests/test_baseline_configs.py::test_scale_limits_counts_past_2_32" -m gpu -x -q > $OUT/r2e_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2e_tests.log
tail -4 $OUT/r2e_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2e_bench1.json 2> $OUT/r2e_bench1.err; echo "bench1 rc=$?"
timeout 900 python bench.py --config 3 --steps 5 > $OUT/r2e_bench3.json 2> $OUT/r2e_bench3.err; echo "bench3 rc=$?"
for f in $OUT/r2e_bench*.json; do echo "== $f"; cut -c1-300 $f; done
tail -5 $OUT/r2e_bench*.err
