#!/bin/bash
# Round 2, 1-GPU call H: the whole GPU suite on the final tree, the bench line, smoke, detector harness configs 3-5 shapes.
tag=${1:-r02_h}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-260 $out/${tag}_bench.json; tail -2 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $out/${tag}_bench_reference.json 2> /dev/null; echo "ref rc=$?"; cut -c1-200 $out/${tag}_bench_reference.json
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
for cfg in faster cascade htc; do timeout 300 python tools/bench_detector.py --config $cfg --steps 10 --warmup 3 > $out/${tag}_detector_$cfg.json 2> $out/${tag}_detector_$cfg.err; echo "detector $cfg rc=$?"; cat $out/${tag}_detector_$cfg.json | cut -c1-420; done
timeout 200 python - <<PY | tee $out/${tag}_nms_timing.log
# merge + class-aware NMS at the LVIS test shape (1000 proposals x 1231 classes, score_thr 0, iou 0.5, 300 detections)
import torch, time
from balancedgroupsoftmax_b200 import ops
from balancedgroupsoftmax_b200.tables import synthetic_tables
t = synthetic_tables(1231, seed=0)
dt = ops.DeviceTables.from_tables(t, 'cuda')
g = torch.Generator().manual_seed(0)
z = (torch.randn(1000, t.num_logits, generator=g) * 3).cuda()
xy = torch.rand(1000, 2, generator=g) * 600
boxes = torch.cat([xy, xy + torch.rand(1000, 2, generator=g) * 200 + 4], 1).cuda()
def run():
    s = ops.merge_scores(z, dt)
    return ops.multiclass_nms(boxes, s, 0.0, 0.5, 300)
for _ in range(3): run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): d, l = run()
torch.cuda.synchronize()
print('merge + NMS (1000 x 1231, thr 0): %.3f ms per image (wall, incl. the one sync), %d detections' % ((time.perf_counter() - t0) / 20 * 1e3, d.shape[0]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s = ops.merge_scores(z, dt)
a.record()
for _ in range(20): ops.merge_scores(z, dt)
b.record(); torch.cuda.synchronize()
print('merge_scores alone: %.1f us' % (a.elapsed_time(b) / 20 * 1e3))
PY
