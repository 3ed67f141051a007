#!/bin/bash
# Round 2, 2-GPU call I: multi-GPU pytest, the default N=2 bench line, stream arrangement of the overlapped exchange,
# NMS stage timing (one GPU).
tag=${1:-r02_i}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest_multi.log 2>&1; echo "pytest(multi) rc=$?"; tail -4 $out/${tag}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29801 bench.py --gpus 2 --steps 240 --warmup 12 > $out/${tag}_bench_2gpu_default.json 2> $out/${tag}_bench_2gpu_default.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/${tag}_bench_2gpu_default.json') if l.startswith('{')][-1])
    print('  ms_per_step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['eager_ms_per_step']))
    print('  collective:', d['config']['collective'][:300]); print('  check:', d['config']['exchange_check'])
except Exception as e:
    print('  no result', e)
PY
for a in "" "--dx-side" "" "--dx-side"; do echo -n "arr=[$a] "; timeout 200 $TR --master-port $((29810+RANDOM%50)) bench.py --gpus 2 --steps 240 --warmup 12 --profile $a 2>/dev/null | grep -E "^\{" | tail -1; done | tee $out/${tag}_split_streams_2gpu.log
timeout 200 python tests/gpu_probe_nms.py 2>&1 | tee $out/${tag}_nms_stages.log | tail -14
