#!/bin/bash
# Round 2, 8-GPU call L: the default bench line at 8 GPUs (value + e2e), detector harness configs 4-5 shapes on 8 GPUs,
# NMS after the 32-warp rework (one GPU).
tag=${1:-r02_l}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29901 bench.py --gpus 8 --steps 240 --warmup 12 > $out/${tag}_bench_8gpu_default.json 2> $out/${tag}_bench_8gpu_default.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/${tag}_bench_8gpu_default.json') if l.startswith('{')][-1])
    print('  ms_per_step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['eager_ms_per_step']))
    print('  collective:', d['config']['collective'][:300]); print('  check:', d['config']['exchange_check']); print('  clocks:', d['clocks'])
except Exception as e:
    print('  no result', e)
PY
tail -2 $out/${tag}_bench_8gpu_default.err | cut -c1-300
for cfg in cascade htc; do timeout 300 $TR --master-port $((29910+RANDOM%40)) tools/bench_detector.py --config $cfg --steps 8 --warmup 3 > $out/${tag}_detector_${cfg}_8gpu.json 2> $out/${tag}_detector_${cfg}_8gpu.err; echo "detector $cfg rc=$?"; grep -E "^\{" $out/${tag}_detector_${cfg}_8gpu.json | cut -c1-330; done
timeout 200 python -m pytest tests/test_gpu_nms.py -m gpu -q 2>&1 | tail -2
timeout 200 python tests/gpu_probe_nms.py 2>&1 | tee $out/${tag}_nms_stages.log | grep -E "nms_dense|sort|topk"
