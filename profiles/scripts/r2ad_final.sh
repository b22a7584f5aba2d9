#!/bin/bash
# round 2, call AD: the whole GPU test suite at HEAD with per-test durations, smoke, the bench line + reference arm, ncu launch list and a
# full capture of direct_count at HEAD
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=60 > $OUT/r2ad_gpu_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2ad_gpu_tests.log
grep -E "passed|failed|rc=" $OUT/r2ad_gpu_tests.log | tail -4
timeout 300 python __graft_entry__.py smoke > $OUT/r2ad_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2ad_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2ad_bench1_ref.json 2> $OUT/r2ad_bench1_ref.err; echo "bench1 ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2ad_bench1.json 2> $OUT/r2ad_bench1.err; echo "bench1 rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $OUT/r2ad_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2ad_bench_launches_ncu.csv $B > $OUT/r2ad_ncu0.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"direct_count|direct_commit" -s 6 -c 2 -o $OUT/prof_direct_r2ad $B > $OUT/r2ad_ncu1.log 2>&1
echo "ncu direct rc=$?"
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_checked.so timeout 600 python -m pytest tests/test_query_counts.py tests/test_baseline_configs.py -m gpu -x -q -k "matches or spliced or pairs_without" > $OUT/r2ad_checked_tests.log 2>&1
echo "checked rc=$?" >> $OUT/r2ad_checked_tests.log
tail -3 $OUT/r2ad_checked_tests.log
cut -c1-600 $OUT/r2ad_bench1.json; cut -c1-300 $OUT/r2ad_bench1_ref.json
tail -n 3 $OUT/r2ad_*.err
ls -la $OUT/*r2ad*.ncu-rep
