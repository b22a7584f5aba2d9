#!/bin/bash
# round 2, call M: the partition's front table in lane-private copies (no bank conflicts); 512- and 1 024-thread partition; 1 M regions
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "scan or bucket or 1M_regions" > $OUT/r2m_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2m_tests.log
tail -4 $OUT/r2m_tests.log
timeout 600 python bench.py --config 2 --steps 5 > $OUT/r2m_bench2.json 2> $OUT/r2m_bench2.err; echo "bench2 rc=$?"
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_wc1024.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan" > $OUT/r2m_tests_wc1024.log 2>&1
echo "tests rc=$?" >> $OUT/r2m_tests_wc1024.log
tail -3 $OUT/r2m_tests_wc1024.log
GTB200_LIB=$PWD/ibm-cbc-genomic-tools_b200/lib/libgtb200_wc1024.so timeout 600 python bench.py --config 2 --steps 5 > $OUT/r2m_bench2_wc1024.json 2> $OUT/r2m_bench2_wc1024.err; echo "bench2 wc1024 rc=$?"
timeout 600 python bench.py --config 4 --reads 1000000000 --steps 5 --no-e2e --no-cpu-baseline > $OUT/r2m_bench4_1b.json 2> $OUT/r2m_bench4_1b.err; echo "bench4 rc=$?"
for f in $OUT/r2m_bench*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["ms_per_step"], d["roofline"].get("step_frac"), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
tail -n 3 $OUT/r2m_bench*.err
