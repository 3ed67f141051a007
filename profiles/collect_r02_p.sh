#!/bin/bash
tag=${1:-r02_p}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_graphed.py -m gpu -q 2>&1 | tail -15 | tee $out/${tag}_pytest_graphed.log
timeout 400 python bench.py --no-library-baseline --steps 300 2> $out/${tag}_bench.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('ms_per_step', d['ms_per_step'], 'e2e', e['value'], 'eager var-N us', e.get('eager_api_us_per_step_variable_n_le_1024'), 'graph-cached var-N us', e.get('graph_cached_api_us_per_step_variable_n_le_1024'))"
tail -3 $out/${tag}_bench.err | cut -c1-300
