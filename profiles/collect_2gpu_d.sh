#!/bin/bash
# 2-GPU: overlapped exchange with few blocks / peer-ldst transport
tag=${1:-r01_v18}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29540
i=0
for cfg in "BAGS_AR_MAX_BLOCKS=16" "BAGS_AR_MAX_BLOCKS=16 BAGS_AR_NO_MULTIMEM=1" "BAGS_AR_NO_MULTIMEM=1" "BAGS_AR_MAX_BLOCKS=48 BAGS_AR_NO_MULTIMEM=1"; do
  port=$((port+1)); i=$((i+1))
  f=$out/${tag}_bench_${N}gpu_cfg$i
  env $cfg timeout 200 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 --allreduce peer --exchange overlap --profile > $f.json 2> $f.err; echo "bench [$cfg] rc=$?"
  tail -1 $f.json | cut -c1-200
done
