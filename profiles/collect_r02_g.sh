#!/bin/bash
# Round 2, 2-GPU call G (shakedown before the 8-GPU call): data-parallel public API test, full default bench line at N=2
tag=${1:-r02_g}
out=gpurun_out
mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tests/multi_gpu_head_step.py > $out/${tag}_head_step_2gpu.log 2>&1; echo "head step rc=$?"; grep -E " ok | FAIL |rror|Traceback" $out/${tag}_head_step_2gpu.log | grep -v "^\[rank[1-9]" | head -20; tail -5 $out/${tag}_head_step_2gpu.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 240 --warmup 12 --exchange instep-overlap-dx --ar-blocks 48 > $out/${tag}_bench_2gpu_full.json 2> $out/${tag}_bench_2gpu_full.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/${tag}_bench_2gpu_full.json') if l.startswith('{')][-1])
    print('  ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['eager_ms_per_step']))
    print('  collective:', d['config']['collective'][:300]); print('  check:', d['config']['exchange_check'])
except Exception as e:
    print('  no result', e)
PY
tail -3 $out/${tag}_bench_2gpu_full.err | cut -c1-300
for b in 96 120; do echo -n "ldst overlapdx b$b "; BAGS_AR_NO_MULTIMEM=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29710+b)) bench.py --gpus 2 --steps 240 --warmup 12 --profile --exchange instep-overlap-dx --ar-blocks $b 2>/dev/null | grep -E "^\{" | tail -1; done
echo -n "mm overlapdx b32 "; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 2 --steps 240 --warmup 12 --profile --exchange instep-overlap-dx --ar-blocks 32 2>/dev/null | grep -E "^\{" | tail -1
