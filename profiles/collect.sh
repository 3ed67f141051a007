#!/bin/bash
# Evidence collection on one B200 (run under gpurun from the repo root):  bash profiles/collect.sh <tag>
# 1. GPU parity suite  2. bench line  3. ncu launch list (gpu__time_duration)  4. ncu --set full of the two GEMM kernels
tag=${1:-r01_v9}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --profile --no-graph --steps 12 --warmup 3 > $out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bags_(fwd|bwd)_fused|sample_others' -s 9 -c 6 \
    -o $out/${tag}_full python bench.py --profile --no-graph --steps 8 --warmup 3 > $out/${tag}_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
