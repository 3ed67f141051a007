#!/bin/bash
# multi-GPU: exchange probe + bench variants (peer overlap with pool 20 / 5, peer inline)
tag=${1:-r01_v16}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tests/multi_gpu_allreduce.py --json $out/${tag}_allreduce_${N}gpu.json > $out/${tag}_allreduce_${N}gpu.log 2>&1; echo "allreduce probe rc=$?"
grep -E " ok | FAIL|us per|Error|error" $out/${tag}_allreduce_${N}gpu.log | grep -v "^\[rank[1-9]" | head -40
port=29512
run() {
  port=$((port+1))
  f=$out/${tag}_bench_${N}gpu_$1; shift
  timeout 300 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 "$@" > $f.json 2> $f.err; echo "bench $* rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$f.json') if l.startswith('{')][-1])
    print('  ms_per_step %.4f  value %.3e  e2e %s  collective: %s' % (d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d['config']['collective'][:60]))
except Exception as e:
    print('  no result', e)
PY
  grep -i "warn\|error\|failed" $f.err | grep -v "OMP_NUM\|destroy_process" | head -5
}
run peer_overlap --allreduce peer --exchange overlap
run peer_overlap_pool5 --allreduce peer --exchange overlap --pool 5
run peer_inline --allreduce peer --exchange inline
if [ "$3" = "nccl" ]; then run nccl_overlap --allreduce nccl --exchange overlap; fi
