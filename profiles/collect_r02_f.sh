#!/bin/bash
# Round 2, 1-GPU call F: remaining tests (trunk, reweight), cluster capacity of this part, dW-only unit shapes for the
# split schedule, stream arrangement of the split schedule with a stand-in exchange.
tag=${1:-r02_f}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_trunk.py tests/test_gpu_reweight.py tests/test_gpu_nms.py -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/${tag}_pytest.log
python - <<PY | tee $out/${tag}_clusters.log
from balancedgroupsoftmax_b200 import _native
import torch
torch.zeros(1, device='cuda')
lib = _native.lib()
for smem in (200 * 1024, 100 * 1024, 0):
    print('dynamic smem %d KB:' % (smem // 1024), {c: lib.bags_debug_max_clusters(c, 320, smem) for c in (1, 2, 3, 4, 5, 6, 7, 8, 16)})
PY
echo "== dW-only unit shape (split schedule alone, N=1): ms_per_step = fwd + dW-only + dX-only"
for cfg in "1 0" "1 2" "1 3" "2 4" "2 5" "2 6" "2 7"; do
  set -- $cfg
  e="BAGS_DW_ONLY_MT=$1"; if [ "$2" != "0" ]; then e="$e BAGS_DW_SPLITS=$2"; fi
  echo -n "$e "; env $e timeout 200 python bench.py --exchange instep-overlap-dx --profile --steps 480 --warmup 20 2>/dev/null | tail -1
done | tee $out/${tag}_dwonly_shapes.log
echo "== stream arrangement with a stand-in exchange of 48 blocks x 20 us"
for a in "" "--dx-side"; do echo -n "arr=[$a] "; timeout 200 python bench.py --exchange instep-overlap-dx $a --fake-exchange 48,256,20 --profile --steps 480 --warmup 20 2>/dev/null | tail -1; done | tee $out/${tag}_split_streams.log
