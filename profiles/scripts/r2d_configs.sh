#!/bin/bash
# round 2, call D: multi-interval fast path + -S admission parity, the bench line, the other BASELINE configs, ncu of HEAD direct_count
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_cli_parity.py tests/test_sharded_gloo.py "tests/test_gpu_parity.py" tests/test_dropin_shim.py "tests/test_baseline_configs.py::test_config3_paired_coverage_abi" "tests/test_baseline_configs.py::test_config3_paired_density_cli" -m gpu -x -q > $OUT/r2d_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2d_tests.log
tail -4 $OUT/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2d_bench1.json 2> $OUT/r2d_bench1.err; echo "bench1 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2d_bench1_ref.json 2>> $OUT/r2d_bench1.err; echo "bench1 ref rc=$?"
timeout 600 python bench.py --config 3 --reads 200000000 --steps 5 > $OUT/r2d_bench3_200m.json 2> $OUT/r2d_bench3.err; echo "bench3 200M rc=$?"
GTB_NO_MULTI_FAST=1 timeout 600 python bench.py --config 3 --reads 200000000 --steps 3 > $OUT/r2d_bench3_200m_general.json 2>> $OUT/r2d_bench3.err; echo "bench3 general rc=$?"
timeout 900 python bench.py --config 3 --steps 5 > $OUT/r2d_bench3.json 2>> $OUT/r2d_bench3.err; echo "bench3 rc=$?"
timeout 900 python bench.py --config 2 --steps 5 > $OUT/r2d_bench2.json 2> $OUT/r2d_bench2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --config 4 --steps 3 > $OUT/r2d_bench4_n1.json 2> $OUT/r2d_bench4.err; echo "bench4 rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $OUT/r2d_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"direct_count|direct_commit" -s 6 -c 2 -o $OUT/prof_direct_r2d $B > $OUT/r2d_ncu1.log 2>&1
echo "ncu rc=$?"
for f in $OUT/r2d_bench*.json; do echo "== $f"; cut -c1-600 $f; done
tail -5 $OUT/r2d_bench*.err
