#!/bin/bash
tag=${1:-r02_n}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graphed.py tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
for i in 1 2; do echo -n "N=1 split (default prep) "; timeout 200 python bench.py --exchange instep-overlap-dx --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2; do echo -n "N=2 default "; timeout 200 $TR --master-port $((29950+i)) bench.py --gpus 2 --steps 240 --warmup 12 --profile 2>/dev/null | grep -E "^\{" | tail -1; done | tee $out/${tag}_bench_2gpu.log
echo -n "N=2 prep=fwd-zero (previous default) "; timeout 200 $TR --master-port 29960 bench.py --gpus 2 --steps 240 --warmup 12 --profile --prep fwd-zero 2>/dev/null | grep -E "^\{" | tail -1 | tee -a $out/${tag}_bench_2gpu.log
timeout 400 $TR --master-port 29961 bench.py --gpus 2 --steps 240 --warmup 12 > $out/${tag}_bench_2gpu_default_full.json 2> $out/${tag}_bench_2gpu_default_full.err; echo "full rc=$?"; cut -c1-200 $out/${tag}_bench_2gpu_default_full.json
