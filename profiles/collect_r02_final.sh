#!/bin/bash
# Round 2, final 1-GPU validation of the committed tree: full GPU suite, default bench line (+ reference arm), smoke.
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads([l for l in open('$out/${tag}_bench.json') if l.startswith('{')][-1])
print('  ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'eager pipeline', d['e2e']['eager_ms_per_step'], 'eager var-N us', d['e2e'].get('eager_api_us_per_step_variable_n_le_1024'))
print('  roofline', d['roofline']['frac'], 'worst', d['roofline_worst']['kernel'], d['roofline_worst']['frac'], 'step', d['roofline_step']['frac'], d['roofline_step']['frac_of_burst_peak'])
print('  lib', d['gpu_library_baseline']['best_us'], 'cpu', d['cpu_baseline']['ms_per_step'], d['cpu_baseline']['cores'], 'kernel_us', d['kernel_us'])
PY
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $out/${tag}_bench_reference.json 2> /dev/null; echo "ref rc=$?"; cut -c1-160 $out/${tag}_bench_reference.json
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
