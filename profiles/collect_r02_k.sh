#!/bin/bash
tag=${1:-r02_k}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
timeout 200 python tests/gpu_probe_nms.py 2>&1 | tee $out/${tag}_nms_stages.log | tail -12
timeout 200 python - <<PY | tee $out/${tag}_nms_timing.log
import torch, time
from balancedgroupsoftmax_b200 import ops
from balancedgroupsoftmax_b200.tables import synthetic_tables
t = synthetic_tables(1231, seed=0)
dt = ops.DeviceTables.from_tables(t, 'cuda')
g = torch.Generator().manual_seed(0)
z = (torch.randn(1000, t.num_logits, generator=g) * 3).cuda()
xy = torch.rand(1000, 2, generator=g) * 600
boxes = torch.cat([xy, xy + torch.rand(1000, 2, generator=g) * 200 + 4], 1).cuda()
def run(thr):
    s = ops.merge_scores(z, dt)
    return ops.multiclass_nms(boxes, s, thr, 0.5, 300)
for thr in (0.0, 1e-3):
    for _ in range(3): run(thr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): d, l = run(thr)
    torch.cuda.synchronize()
    print('merge + NMS (1000 x 1231, score_thr %g): %.3f ms per image (wall, incl. the one sync), %d detections' % (thr, (time.perf_counter() - t0) / 20 * 1e3, d.shape[0]))
PY
for cfg in faster cascade htc; do timeout 300 python tools/bench_detector.py --config $cfg --steps 10 --warmup 3 > $out/${tag}_detector_$cfg.json 2> /dev/null; echo "detector $cfg rc=$?"; cat $out/${tag}_detector_$cfg.json | cut -c1-330; done
