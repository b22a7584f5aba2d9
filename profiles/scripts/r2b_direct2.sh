#!/bin/bash
# round 2, call B: the DIRECT engine's second form after the instruction diet -- parity, timing variants, one ncu capture
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "direct or golden or random or sorted or hg19 or full_size or dense or streaming or fatal or negative" > $OUT/r2b_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2b_tests.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run() { echo "== $*" >> $OUT/r2b_variants.log; env "$@" GTB_DEBUG_DIRECT=1 timeout 300 $B >> $OUT/r2b_variants.log 2>&1; echo "rc=$?" >> $OUT/r2b_variants.log; }
run GTB_DIRECT_FORM=1
run GTB_DIRECT_FORM=2
run GTB_DIRECT2_SMEM_KB=227
run GTB_DIRECT2_SMEM_KB=211
run GTB_DIRECT2_QCAP=1024
tail -3 $OUT/r2b_tests.log
grep -E "^==|value|second form" $OUT/r2b_variants.log | cut -c1-330
$B > $OUT/r2b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:direct2_count -s 3 -c 1 -o $OUT/prof_direct2_r2b $B > $OUT/r2b_ncu.log 2>&1
echo "ncu rc=$?"
