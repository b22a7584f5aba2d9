#!/bin/bash
# round 2, call I (8 GPUs): configs[4] -- 4 B reads x 1 M regions, count, STRONG scaling at 1 / 2 / 4 / 8 GPUs -- and configs[1] weak at 8
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/r2i_smi.txt 2>&1
nproc >> $OUT/r2i_smi.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --config 4 --steps 5 > $OUT/r2i_bench4_n1.json 2> $OUT/r2i_bench4_n1.err; echo "bench4 n1 rc=$?"
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --config 4 --steps 5 > $OUT/r2i_bench4_n$n.json 2> $OUT/r2i_bench4_n$n.err; echo "bench4 n$n rc=$?"
done
timeout 600 $TR --nproc-per-node 8 --master-port 29650 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/r2i_bench1_n8.json 2> $OUT/r2i_bench1_n8.err; echo "bench1 n8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29651 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/r2i_bench1_n4.json 2> $OUT/r2i_bench1_n4.err; echo "bench1 n4 rc=$?"
for f in $OUT/r2i_bench*.json; do echo "== $f"; grep '^{' $f | cut -c1-330; done
for f in $OUT/r2i_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" $f | tail -n 4; done
