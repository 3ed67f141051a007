#!/bin/bash
# Round 2, 2-GPU call C: exchange protocol switches (BAGS_AR_MODE bits: 1 relaxed first-barrier store, 2 relaxed polling,
# 4 late dependent launch) x grid x transport, in the step; the split backward (dW+db -> exchange || dX) with 128x256 dW units.
tag=${1:-r02_c}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nms.py tests/test_gpu_harness.py -m gpu -x -q -k "split_backward or nms or harness or unit_size" > $out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"; tail -4 $out/${tag}_pytest_new.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench(1gpu) rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/${tag}_bench_1gpu.json') if l.startswith('{')][-1])
    print('  ms_per_step', d['ms_per_step'], 'kernel_us', d['kernel_us'], 'e2e', d['e2e']['value'], 'eager', d['e2e']['eager_ms_per_step'])
except Exception as e:
    print('  no result', e)
PY
for ex in instep instep-overlap-dx; do echo -n "N=1 $ex "; timeout 200 python bench.py --exchange $ex --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_split_1gpu.log
port=29560
for mode in 0 3 7; do
  port=$((port+1))
  BAGS_AR_MODE=$mode DIAG_GRIDS=48,148,320 timeout 300 $TR --master-port $port tests/multi_gpu_exchange_diag.py --json $out/${tag}_exchange_diag_${N}gpu_mode$mode.json > $out/${tag}_exchange_diag_${N}gpu_mode$mode.log 2>&1; echo "diag mode=$mode rc=$?"
  grep -E "^\{" $out/${tag}_exchange_diag_${N}gpu_mode$mode.log | cut -c1-420
done
run_bench() {  # name, env assignments (string), extra args
  port=$((port+1))
  f=$out/${tag}_bench_${N}gpu_$1
  envs=$2
  shift; shift
  env $envs timeout 300 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 --profile "$@" > $f.json 2> $f.err; rc=$?
  echo -n "$(basename $f) rc=$rc "; grep -E "^\{" $f.json | tail -1 | cut -c1-120
}
for mode in 0 3 7; do
  run_bench mm_instep_b148_mode$mode "BAGS_AR_MODE=$mode" --exchange instep --ar-blocks 148
  run_bench ldst_instep_b148_mode$mode "BAGS_AR_MODE=$mode BAGS_AR_NO_MULTIMEM=1" --exchange instep --ar-blocks 148
done
run_bench ldst_instep_b320_mode3 "BAGS_AR_MODE=3 BAGS_AR_NO_MULTIMEM=1" --exchange instep --ar-blocks 320
for b in 48 148; do
  run_bench mm_overlapdx_b${b}_mode3 "BAGS_AR_MODE=3" --exchange instep-overlap-dx --ar-blocks $b
  run_bench ldst_overlapdx_b${b}_mode3 "BAGS_AR_MODE=3 BAGS_AR_NO_MULTIMEM=1" --exchange instep-overlap-dx --ar-blocks $b
done
run_bench ldst_overlapdx_b148_mode7 "BAGS_AR_MODE=7 BAGS_AR_NO_MULTIMEM=1" --exchange instep-overlap-dx --ar-blocks 148
run_bench nccl_overlapdx "BAGS_AR_MODE=0" --exchange instep-overlap-dx --allreduce nccl
