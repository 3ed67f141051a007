#!/bin/bash
# Round 2, 1-GPU call D: full parity suite on the current tree (trunk GEMMs, preparation in the forward, dense NMS,
# harness), bench line, preparation variants A/B, the split launches alone, ncu launch list + full capture.
tag=${1:-r02_d}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/${tag}_pytest.log
for rep in 1 2; do for prep in bwd fwd-zero fwd-zero-colsum; do echo -n "prep=$prep "; timeout 200 python bench.py --prep $prep --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done; done | tee $out/${tag}_prep_ab.log
for prep in bwd fwd-zero fwd-zero-colsum; do echo -n "split prep=$prep "; timeout 200 python bench.py --prep $prep --exchange instep-overlap-dx --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_split_prep.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/${tag}_bench.json') if l.startswith('{')][-1])
    print('  ms_per_step', d['ms_per_step'], 'kernel_us', d['kernel_us'], 'e2e', d['e2e']['value'], 'eager', d['e2e']['eager_ms_per_step'])
    print('  lib', d['gpu_library_baseline']['best_us'], 'cpu', d['cpu_baseline'])
except Exception as e:
    print('  no result', e)
PY
timeout 300 python tools/bench_detector.py --config faster --steps 10 --warmup 3 > $out/${tag}_detector_faster.json 2> $out/${tag}_detector_faster.err; echo "detector rc=$?"; cat $out/${tag}_detector_faster.json; tail -2 $out/${tag}_detector_faster.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --profile --no-graph --steps 12 --warmup 3 > $out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bags_(fwd|bwd)_fused|sample_others' -s 9 -c 6 \
    -o $out/${tag}_full python bench.py --profile --no-graph --steps 8 --warmup 3 > $out/${tag}_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
