#!/bin/bash
# round 2, call V: the whole GPU test suite at HEAD, the bench lines of all configs, the reference arm, ncu launch list + full captures
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2v_gpu_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2v_gpu_tests.log
tail -4 $OUT/r2v_gpu_tests.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2v_bench1_ref.json 2> $OUT/r2v_bench1_ref.err; echo "bench1 ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2v_bench1.json 2> $OUT/r2v_bench1.err; echo "bench1 rc=$?"
timeout 900 python bench.py --config 2 --steps 5 > $OUT/r2v_bench2.json 2> $OUT/r2v_bench2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --config 3 --steps 5 > $OUT/r2v_bench3.json 2> $OUT/r2v_bench3.err; echo "bench3 rc=$?"
timeout 900 python bench.py --config 4 --steps 3 > $OUT/r2v_bench4_n1.json 2> $OUT/r2v_bench4.err; echo "bench4 rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > $OUT/r2v_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2v_bench_launches_ncu.csv $B > $OUT/r2v_ncu0.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"direct_count|direct_commit" -s 6 -c 2 -o $OUT/prof_direct_r2v $B > $OUT/r2v_ncu1.log 2>&1
echo "ncu direct rc=$?"
S="python bench.py --config 2 --reads 200000000 --steps 1 --warmup 3"
$S > $OUT/r2v_scan200m.json 2> $OUT/r2v_scan200m.err && ncu --set full --clock-control none --import-source on -k regex:"wc_partition|scan_bucket_hist|scan_windows" -s 12 -c 4 -o $OUT/prof_scan_r2v $S > $OUT/r2v_ncu2.log 2>&1
echo "ncu scan rc=$?"
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_checked.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_query_counts.py tests/test_mgpu.py -m gpu -x -q -k "not cli" > $OUT/r2v_checked_tests.log 2>&1
echo "checked rc=$?" >> $OUT/r2v_checked_tests.log
tail -3 $OUT/r2v_checked_tests.log
for f in $OUT/r2v_bench*.json; do echo "== $f"; cut -c1-400 $f; done
tail -n 4 $OUT/r2v_*.err
ls -la $OUT/*.ncu-rep
