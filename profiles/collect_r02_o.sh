#!/bin/bash
tag=${1:-r02_o}
out=gpurun_out
mkdir -p $out
for n in 8 4; do echo -n "N=$n default "; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29970+n)) bench.py --gpus $n --steps 240 --warmup 12 --profile 2>/dev/null | grep -E "^\{" | tail -1; done | tee $out/${tag}_bench_default_8_4.log
