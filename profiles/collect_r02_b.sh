#!/bin/bash
# Round 2, 2-GPU call B: the gradient exchange -- correctness on random gradients, in-kernel phase timeline and isolated
# timing per transport / grid, then the step with the exchange IN the step (three schedules) next to NCCL.
tag=${1:-r02_b}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# new single-GPU tests first (split backward halves on the merged kernel, dense NMS)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nms.py -m gpu -x -q -k "split_backward or nms" > $out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"; tail -4 $out/${tag}_pytest_new.log
timeout 300 $TR --master-port 29541 tests/multi_gpu_allreduce.py --quick > $out/${tag}_allreduce_${N}gpu.log 2>&1; echo "allreduce probe rc=$?"; grep -E "ok|FAIL|us per" $out/${tag}_allreduce_${N}gpu.log | grep -v "^\[rank[1-9]" | head -20
timeout 400 $TR --master-port 29542 tests/multi_gpu_exchange_diag.py --json $out/${tag}_exchange_diag_${N}gpu.json > $out/${tag}_exchange_diag_${N}gpu.log 2>&1; echo "diag rc=$?"; grep -E "^\{" $out/${tag}_exchange_diag_${N}gpu.log | cut -c1-700
port=29550
run_bench() {  # name, extra args
  port=$((port+1))
  f=$out/${tag}_bench_${N}gpu_$1
  shift
  timeout 300 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 --profile "$@" > $f.json 2> $f.err; rc=$?
  echo -n "$(basename $f) rc=$rc "; grep -E "^\{" $f.json | tail -1 | cut -c1-200
}
run_bench peer_instep --exchange instep
run_bench peer_instep_b148 --exchange instep --ar-blocks 148
run_bench peer_instep_b320 --exchange instep --ar-blocks 320
run_bench peer_overlapdx_b148 --exchange instep-overlap-dx --ar-blocks 148
run_bench peer_overlapdx_b48 --exchange instep-overlap-dx --ar-blocks 48
run_bench peer_overlapdx_b16 --exchange instep-overlap-dx --ar-blocks 16
run_bench nccl_instep --exchange instep --allreduce nccl
run_bench nccl_overlapdx --exchange instep-overlap-dx --allreduce nccl
run_bench peer_nextstep --exchange overlap-next-step
# the split launches alone on one GPU (what the in-step overlap schedule costs before any exchange)
for ex in instep instep-overlap-dx; do echo -n "N=1 $ex "; timeout 200 python bench.py --exchange $ex --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_split_1gpu.log
# full line (e2e etc.) of the default N-GPU configuration
port=$((port+1)); timeout 400 $TR --master-port $port bench.py --gpus $N --steps 240 --warmup 12 > $out/${tag}_bench_${N}gpu_default_full.json 2> $out/${tag}_bench_${N}gpu_default_full.err; echo "full rc=$?"; cut -c1-400 $out/${tag}_bench_${N}gpu_default_full.json
