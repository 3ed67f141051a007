#!/bin/bash
# Round 2, 8-GPU call E: the exchange at 8 (and 4) ranks -- random-gradient exactness + phase timeline, the data-parallel
# public API (grad_bucket) against local backward + NCCL mean, then the step with the exchange inside it (serial, and
# overlapped by the dX GEMM), NCCL beside it; detector harness at 8 GPUs.
tag=${1:-r02_e}
out=gpurun_out
mkdir -p $out
port=29600
TRN() { n=$1; shift; port=$((port+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
DIAG_GRIDS=16,40 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 tests/multi_gpu_exchange_diag.py --json $out/${tag}_exchange_diag_8gpu.json > $out/${tag}_exchange_diag_8gpu.log 2>&1; echo "diag rc=$?"
grep -E "^\{" $out/${tag}_exchange_diag_8gpu.log | cut -c1-1100
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 tests/multi_gpu_head_step.py > $out/${tag}_head_step_8gpu.log 2>&1; echo "head step rc=$?"; grep -E " ok | FAIL |rror" $out/${tag}_head_step_8gpu.log | grep -v "^\[rank[1-9]" | head -16
port=29610
run_bench() {  # N, name, env assignments, extra args
  port=$((port+1))
  n=$1; f=$out/${tag}_bench_${n}gpu_$2; envs=$3
  shift; shift; shift
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 160 --warmup 10 --profile "$@" > $f.json 2> $f.err; rc=$?
  echo -n "$(basename $f) rc=$rc "; grep -E "^\{" $f.json | tail -1 | cut -c1-120
}
run_bench 8 mm_instep "A=0" --exchange instep
run_bench 8 ldst_instep "BAGS_AR_NO_MULTIMEM=1" --exchange instep
run_bench 8 mm_overlapdx "A=0" --exchange instep-overlap-dx
run_bench 8 mm_overlapdx_b16 "A=0" --exchange instep-overlap-dx --ar-blocks 16
run_bench 8 nccl_instep "A=0" --exchange instep --allreduce nccl
run_bench 4 mm_instep "A=0" --exchange instep
run_bench 4 mm_overlapdx "A=0" --exchange instep-overlap-dx --ar-blocks 48
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 tools/bench_detector.py --config faster --steps 8 --warmup 3 > $out/${tag}_detector_faster_8gpu.json 2> $out/${tag}_detector_faster_8gpu.err; echo "detector rc=$?"; cat $out/${tag}_detector_faster_8gpu.json
