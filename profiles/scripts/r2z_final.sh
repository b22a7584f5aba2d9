#!/bin/bash
# round 2, call Z: the whole GPU test suite at the final HEAD, smoke(), the default bench line, bench_cli
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/r2z_gpu_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2z_gpu_tests.log
tail -6 $OUT/r2z_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > $OUT/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/r2z_smoke.log
timeout 600 python bench.py > $OUT/r2z_bench1.json 2> $OUT/r2z_bench1.err; echo "bench1 rc=$?"; cut -c1-260 $OUT/r2z_bench1.json
timeout 900 python bench_cli.py --out $OUT/r2z_cli_end_to_end.jsonl > $OUT/r2z_cli.log 2>&1; echo "cli rc=$?"; tail -4 $OUT/r2z_cli.log | cut -c1-300
