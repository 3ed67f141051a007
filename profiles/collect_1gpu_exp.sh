#!/bin/bash
# 1-GPU experiments: parity suite, bench (graphed e2e), backward timeline by unit kind, K-major operand timing
tag=${1:-r01_v12}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/${tag}_pytest.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
timeout 200 python tests/gpu_probe.py --case timeline > $out/${tag}_timeline.log 2>&1; echo "timeline rc=$?"
grep RESULT $out/${tag}_timeline.log | cut -c8- > $out/${tag}_timeline.json
for i in 1 2; do
for lib in libbags_b200.so libbags_b200_xk.so; do
  echo -n "$lib " >> $out/${tag}_kmajor_ab.log
  BAGS_LIB=$PWD/balancedgroupsoftmax_b200/$lib timeout 200 python bench.py --profile --steps 480 --warmup 20 2>/dev/null | tail -1 >> $out/${tag}_kmajor_ab.log
done; done
cat $out/${tag}_kmajor_ab.log
timeout 120 python tests/gpu_probe_h2d.py > $out/${tag}_h2d.json 2>&1; cat $out/${tag}_h2d.json | cut -c1-50; python -c "
import json;d=json.load(open('$out/${tag}_h2d.json'));print({k:v for k,v in d.items() if k.startswith('gbs') or k.startswith('thr')})"
