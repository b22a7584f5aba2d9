#!/bin/bash
# round 2, call F (2 GPUs): radix sort + packed entry tests, the sharded bench at N=2 (weak configs[1], strong configs[4] on a quarter of the reads)
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_regions or packed or sharded" > $OUT/r2f_tests.log 2>&1
echo "tests rc=$?" >> $OUT/r2f_tests.log
tail -4 $OUT/r2f_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2f_bench1_n2.json 2> $OUT/r2f_bench1_n2.err; echo "bench1 n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 4 --reads 1000000000 --steps 3 > $OUT/r2f_bench4_n2_1b.json 2> $OUT/r2f_bench4_n2.err; echo "bench4 n2 rc=$?"
for f in $OUT/r2f_bench*.json; do echo "== $f"; cut -c1-400 $f; done
tail -n 6 $OUT/r2f_bench1_n2.err; tail -n 6 $OUT/r2f_bench4_n2.err
