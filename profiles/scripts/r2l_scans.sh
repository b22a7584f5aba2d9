#!/bin/bash
# round 2, call L: window counts on a uint32 table (+ carry plane), compact window output, buckets walked by 3 CTAs instead of 4;
# the partition with 1 024 threads / up to 1 024 buckets (libgtb200_wc1024.so) against the 512-thread one; DSMEM atomic rates
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py tests/test_cli_parity.py -m gpu -x -q -k "scan" > $OUT/r2l_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2l_tests.log
tail -4 $OUT/r2l_tests.log
timeout 600 python bench.py --config 2 --steps 5 > $OUT/r2l_bench2.json 2> $OUT/r2l_bench2.err; echo "bench2 rc=$?"
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_wc1024.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan" > $OUT/r2l_tests_wc1024.log 2>&1
echo "tests rc=$?" >> $OUT/r2l_tests_wc1024.log
tail -3 $OUT/r2l_tests_wc1024.log
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_wc1024.so timeout 600 python bench.py --config 2 --steps 5 > $OUT/r2l_bench2_wc1024.json 2> $OUT/r2l_bench2_wc1024.err; echo "bench2 wc1024 rc=$?"
for f in $OUT/r2l_bench2*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["ms_per_step"], d["roofline"]["step_frac"], {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
tail -3 $OUT/r2l_bench2*.err
timeout 120 profiles/microbench/dsmem_rate > $OUT/r2l_dsmem_rate.txt 2>&1; cat $OUT/r2l_dsmem_rate.txt
