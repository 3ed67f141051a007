#!/bin/bash
# 2-GPU evidence (run under gpurun --gpus 2): peer-memory exchange vs NCCL, bench at N=2 with both
tag=${1:-r01_v10}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tests/multi_gpu_allreduce.py --json $out/${tag}_allreduce_2gpu.json > $out/${tag}_allreduce_2gpu.log 2>&1; echo "allreduce probe rc=$?"
grep -v "^\[rank1\]" $out/${tag}_allreduce_2gpu.log | grep -v "^$" | head -60
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 240 --warmup 12 --allreduce peer > $out/${tag}_bench_2gpu_peer.json 2> $out/${tag}_bench_2gpu_peer.err; echo "bench peer rc=$?"
cat $out/${tag}_bench_2gpu_peer.json; grep -i "warn\|error\|peer" $out/${tag}_bench_2gpu_peer.err | head
if [ "$2" = "nccl" ]; then
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 240 --warmup 12 --allreduce nccl > $out/${tag}_bench_2gpu_nccl.json 2> $out/${tag}_bench_2gpu_nccl.err; echo "bench nccl rc=$?"
cat $out/${tag}_bench_2gpu_nccl.json
fi
