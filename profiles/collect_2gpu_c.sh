#!/bin/bash
# 2-GPU: exchange protocol / block-size variants (probe timings), then bench with the two best candidates
tag=${1:-r01_v17}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29520
for cfg in "0 256" "1 256" "1 128" "0 128"; do
  set -- $cfg
  port=$((port+1))
  echo "== BAGS_AR_EPOCH=$1 BAGS_AR_THREADS=$2"
  BAGS_AR_EPOCH=$1 BAGS_AR_THREADS=$2 timeout 120 $TR --master-port $port tests/multi_gpu_allreduce.py --quick > $out/${tag}_probe_e$1_t$2.log 2>&1; echo "rc=$?"
  grep -E "FAIL|us per|rror" $out/${tag}_probe_e$1_t$2.log | grep -v "^\[rank[1-9]" | head -8
done
tag=${tag}
for cfg in "0 256" "1 256"; do
  set -- $cfg
  port=$((port+1))
  f=$out/${tag}_bench_${N}gpu_e$1_t$2_overlap
  BAGS_AR_EPOCH=$1 BAGS_AR_THREADS=$2 timeout 300 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 --allreduce peer --exchange overlap > $f.json 2> $f.err; echo "bench epoch=$1 threads=$2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$f.json') if l.startswith('{')][-1])
    print('  ms_per_step %.4f  value %.3e  e2e %s  collective: %s' % (d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d['config']['collective'][:60]))
except Exception as e:
    print('  no result', e)
PY
done
