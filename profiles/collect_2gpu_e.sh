#!/bin/bash
# 2-GPU: overlapped multimem exchange, grid-shape variants (threads per block x blocks)
tag=${1:-r01_v19}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29560
i=0
for cfg in "BAGS_AR_THREADS=128 BAGS_AR_MAX_BLOCKS=32" "BAGS_AR_THREADS=256 BAGS_AR_MAX_BLOCKS=32"; do
  port=$((port+1)); i=$((i+1))
  f=$out/${tag}_bench_${N}gpu_cfg$i
  env $cfg timeout 100 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 --allreduce peer --exchange overlap --profile > $f.json 2> $f.err; echo "bench [$cfg] rc=$?"
  tail -1 $f.json | cut -c1-200
done
