#!/bin/bash
# 1-GPU scheduling probe: does a latency-bound side-stream kernel hide under the next step's GEMM kernels?
tag=${1:-r01_v15}
out=gpurun_out
mkdir -p $out
log=$out/${tag}_sched_probe.log
: > $log
for fe in "" "155,256,30" "16,256,30" "155,32,30" "1,32,30" "155,256,5" "74,256,30" ""; do
  echo -n "fake_exchange=[$fe] " >> $log
  timeout 200 python bench.py --profile --steps 480 --warmup 20 ${fe:+--fake-exchange $fe} 2>/dev/null | tail -1 >> $log
done
cat $log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json
