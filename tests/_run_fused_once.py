import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from balancedgroupsoftmax_b200 import ops
from balancedgroupsoftmax_b200.tables import synthetic_tables
t = synthetic_tables()
dt = ops.DeviceTables.from_tables(t, 'cuda')
N = 4096
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(N, 1024, generator=g)).cuda().bfloat16()
W = (torch.randn(t.num_logits, 1024, generator=g) * 0.05).cuda().bfloat16()
b = torch.zeros(t.num_logits).cuda()
labels = torch.zeros(N, dtype=torch.long)
for s in range(0, N, 512):
    labels[s:s + 128] = torch.randint(1, 1231, (128,), generator=g)
labels = labels.cuda()
wmask, avg = ops.sample_others(labels, dt, 8.0, 1)
for _ in range(4):
    out = ops.fused_fwd(x, W, b, labels, dt, wmask, avg)
torch.cuda.synchronize()
print('done', out[0].tolist())
