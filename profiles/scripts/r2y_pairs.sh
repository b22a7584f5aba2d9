#!/bin/bash
# round 2, call Y: read pairs as regions of two intervals without offsets (pair check inside the DIRECT engine)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_baseline_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "config3 or golden or random or direct or multi" > $OUT/r2y_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2y_tests.log
tail -8 $OUT/r2y_tests.log
timeout 900 python bench.py --config 3 --steps 5 > $OUT/r2y_bench3.json 2> $OUT/r2y_bench3.err; echo "bench3 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2y_bench3.json"):
    d=json.loads(l); print(d["metric"], d["config"]["workload"][-40:], round(d["ms_per_step"],3), round(d["roofline"]["step_frac"],3), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
PY
tail -n 3 $OUT/r2y_bench3.err
