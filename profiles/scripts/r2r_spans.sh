#!/bin/bash
# round 2, call R: coverage of spans (-gaps) through the DIRECT engine
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_baseline_configs.py -m gpu -x -q -k "config3 or direct or golden or random" > $OUT/r2r_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2r_tests.log
tail -5 $OUT/r2r_tests.log
timeout 900 python bench.py --config 3 --steps 5 > $OUT/r2r_bench3.json 2> $OUT/r2r_bench3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2r_bench3.json"):
    d=json.loads(l); print(d["metric"], d["ms_per_step"], {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
tail -n 3 $OUT/r2r_bench3.err
