#!/bin/bash
tag=${1:-r02_m}
out=gpurun_out
mkdir -p $out
for k in 1 2 4 1 2; do echo -n "copy streams $k: "; timeout 300 python bench.py --e2e-copy-streams $k --no-library-baseline --steps 300 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('e2e %.2f M RoIs/s  %.1f us/step  h2d_in_pipe %.1f us  h2d_only %.1f us  eager %.1f us' % (e['value']/1e6, e['ms_per_step']*1e3, (e['h2d_ms_inside_pipeline'] or 0)*1e3, e['h2d_only_ms_per_step']*1e3, e['eager_ms_per_step']*1e3))"; done | tee $out/${tag}_e2e_copy_streams.log
timeout 200 python tests/gpu_probe_nms.py 2>&1 | grep -E "nms_dense" | tee $out/${tag}_nms.log
timeout 300 python -m pytest tests/test_gpu_nms.py -m gpu -q 2>&1 | tail -2
echo "== dW-only launch of the split schedule: merged kernel (in-kernel preparation) vs plain GEMM + preparation kernel (N=1, fwd + dW-only + dX-only)"
for rep in 1 2; do for v in 1 0; do echo -n "BAGS_BWD_DW_MERGED=$v "; BAGS_BWD_DW_MERGED=$v timeout 200 python bench.py --exchange instep-overlap-dx --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done; done | tee $out/${tag}_dwonly_path.log
echo -n "BAGS_BWD_DW_MERGED=0 --prep fwd-zero-colsum "; BAGS_BWD_DW_MERGED=0 timeout 200 python bench.py --prep fwd-zero-colsum --exchange instep-overlap-dx --profile --steps 480 --warmup 20 2>/dev/null | tail -1 | tee -a $out/${tag}_dwonly_path.log
