"""Stage timing of ops.multiclass_nms at the LVIS test shape (1000 proposals x 1231 classes): python tests/gpu_probe_nms.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balancedgroupsoftmax_b200 import ops, _native as nat
from balancedgroupsoftmax_b200.tables import synthetic_tables
t = synthetic_tables(1231, seed=0)
dt = ops.DeviceTables.from_tables(t, 'cuda')
g = torch.Generator().manual_seed(0)
z = (torch.randn(1000, t.num_logits, generator=g) * 3).cuda()
xy = torch.rand(1000, 2, generator=g) * 600
boxes = torch.cat([xy, xy + torch.rand(1000, 2, generator=g) * 200 + 4], 1).cuda()
scores = ops.merge_scores(z, dt)
n, S = 1000, 1230
def timed(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    print('%-34s %8.1f us' % (name, a.elapsed_time(b) / reps * 1e3), flush=True)
    return out
fg = timed('slice+transpose+contiguous', lambda: scores[:, 1:].float().t().contiguous())
vals, order = timed('sort stable desc [1230,1000]', lambda: fg.sort(dim=1, descending=True, stable=True))
timed('sort (not stable)', lambda: fg.sort(dim=1, descending=True))
counts = timed('counts', lambda: (vals > 0.0).sum(dim=1, dtype=torch.int32))
order32 = timed('order->int32', lambda: order.to(torch.int32))
keep = torch.empty((S, n), dtype=torch.uint8, device='cuda')
ovf = torch.zeros(1, dtype=torch.int32, device='cuda')
def kern():
    nat.check(nat.lib().bags_class_nms_dense(boxes.data_ptr(), 4, order32.data_ptr(), counts.data_ptr(), S, n, 0.5,
                                             keep.data_ptr(), ovf.data_ptr(), torch.cuda.current_stream().cuda_stream), 'nms')
timed('bags_class_nms_dense (thr 0: 1000/class)', kern)
c2 = (vals > 0.05).sum(dim=1, dtype=torch.int32)
def kern2():
    nat.check(nat.lib().bags_class_nms_dense(boxes.data_ptr(), 4, order32.data_ptr(), c2.data_ptr(), S, n, 0.5,
                                             keep.data_ptr(), ovf.data_ptr(), torch.cuda.current_stream().cuda_stream), 'nms')
timed('bags_class_nms_dense (thr 0.05: mean %.0f/class)' % c2.float().mean().item(), kern2)
kept = keep.bool()
flat = timed('where/-inf', lambda: torch.where(kept, vals, torch.full_like(vals, float('-inf'))).reshape(-1))
timed('topk 300 of 1.23M', lambda: flat.topk(300))
