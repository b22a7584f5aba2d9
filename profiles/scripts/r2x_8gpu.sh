#!/bin/bash
# round 2, call X (N GPUs): gtb_mgpu end to end from host memory at 1 .. N devices in one process; bench.py under torchrun at HEAD
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $OUT/r2x_box.txt
timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q > $OUT/r2x_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2x_tests.log
tail -3 $OUT/r2x_tests.log
timeout 600 python profiles/scripts/time_mgpu.py 800000000 > $OUT/r2x_mgpu.json 2> $OUT/r2x_mgpu.err; echo "rc=$?"; cat $OUT/r2x_mgpu.json; tail -n 3 $OUT/r2x_mgpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2x_bench1_n$N.json 2> $OUT/r2x_bench1_n$N.err; echo "bench rc=$?"; cut -c1-300 $OUT/r2x_bench1_n$N.json
