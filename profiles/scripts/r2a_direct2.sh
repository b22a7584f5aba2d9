#!/bin/bash
# round 2, call A: the DIRECT engine's second form -- parity, then timing variants (device-resident, CUDA events)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "direct or golden or random or sorted or hg19 or full_size or dense or streaming" > $OUT/r2a_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2a_tests.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
run() { echo "== $*" >> $OUT/r2a_variants.log; env "$@" GTB_DEBUG_DIRECT=1 timeout 300 $B >> $OUT/r2a_variants.log 2>&1; echo "rc=$?" >> $OUT/r2a_variants.log; }
run GTB_DIRECT_FORM=1
run GTB_DIRECT_FORM=2
run GTB_DIRECT2_SMEM_KB=163
run GTB_DIRECT2_SMEM_KB=227
run GTB_DIRECT2_SMEM_KB=211
run GTB_DIRECT2_QCAP=1024
run GTB_DIRECT2_QCAP=640
run GTB_DIRECT2_CELL_BP=16384
run GTB_DIRECT2_CELL_BP=32768
tail -5 $OUT/r2a_tests.log
grep -E "^==|value|second form" $OUT/r2a_variants.log | cut -c1-400
