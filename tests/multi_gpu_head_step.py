"""Multi-GPU check of the data-parallel public API (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 \
        tests/multi_gpu_head_step.py

Every rank runs the head on its own RoIs (bags_head_loss / GraphedHeadStep with ``grad_bucket=``): dW / db are written into
the NVLink exchange bucket, exchanged as soon as they are complete and while dX is computed.  Checked against the
reference's order of operations -- local backward, then all_reduce + divide of the flattened gradients
(mmdet/core/utils/dist_utils.py:9-41, :51-58) -- on the same inputs and masks: weight.grad / bias.grad equal the NCCL
mean to fp32 summation-order rounding (the split-K red.add order of dW is not fixed), dX equals the local one bit for bit.
"""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=120))
    from balancedgroupsoftmax_b200 import ops
    from balancedgroupsoftmax_b200.api import GraphedHeadStep, bags_head_loss
    from balancedgroupsoftmax_b200.dist import NcclGradBucket, PeerGradBucket
    from balancedgroupsoftmax_b200.tables import synthetic_tables
    t = synthetic_tables(1231, seed=0)
    dt = ops.DeviceTables.from_tables(t, dev)
    C, K, N = t.num_logits, 1024, 1024
    out = {'world': world, 'ok': True, 'checks': []}

    def check(name, cond, detail=''):
        v = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MIN)
        good = bool(v.item())
        out['checks'].append({'name': name, 'ok': good, 'detail': detail})
        out['ok'] = out['ok'] and good
        if rank == 0:
            print('%-70s %s %s' % (name, 'ok' if good else 'FAIL', detail), flush=True)

    g = torch.Generator().manual_seed(11)
    W0 = (torch.randn(C, K, generator=g) * 0.05).to(dev)         # same weights on every rank
    b0 = (torch.randn(C, generator=g) * 0.1).to(dev)
    gr = torch.Generator().manual_seed(100 + rank)               # own RoIs per rank
    x = torch.relu(torch.randn(N, K, generator=gr)).to(dev).bfloat16()
    labels = torch.zeros(N, dtype=torch.int64)
    labels[:N // 4] = torch.randint(1, t.num_classes, (N // 4,), generator=gr)
    labels = labels.to(dev)
    wmask, avg = ops.sample_others(labels, dt, 8.0, 555 + rank)
    # the exchange itself is exact at 2 ranks (bench.py / multi_gpu_exchange_diag.py check that on random data); here the
    # two sides also differ in the ORDER of the split-K red.add of dW (fp32 atomics: last-bit differences)
    tol = 1e-6

    def reference():
        W = torch.nn.Parameter(W0.clone())
        b = torch.nn.Parameter(b0.clone())
        xr = x.clone().requires_grad_(True)
        bags_head_loss(xr, W, b, labels, dt, wmask=wmask, avg=avg).sum().backward()
        flat = torch.cat([W.grad.reshape(-1), b.grad.reshape(-1)])
        dist.all_reduce(flat)
        flat /= world                                                  # dist_utils.py:20-24
        return flat[:C * K].view(C, K), flat[C * K:], xr.grad

    gW_ref, gb_ref, gx_ref = reference()
    for kind in ('peer', 'nccl'):
        bucket = PeerGradBucket([(C, K), (C,)], dev) if kind == 'peer' else NcclGradBucket([(C, K), (C,)], dev)
        if kind == 'peer':
            assert bucket.self_test()
        for it in range(2):                                        # twice: flags / epochs / zeroing are re-armed
            W = torch.nn.Parameter(W0.clone())
            b = torch.nn.Parameter(b0.clone())
            xr = x.clone().requires_grad_(True)
            bags_head_loss(xr, W, b, labels, dt, wmask=wmask, avg=avg, grad_bucket=bucket).sum().backward()
            torch.cuda.synchronize()
            eW = (W.grad - gW_ref).abs().max().item()
            eb = (b.grad - gb_ref).abs().max().item()
            check('%s bucket, eager pass %d: weight.grad / bias.grad == NCCL mean' % (kind, it), eW <= tol and eb <= tol,
                  'max abs err %.2e / %.2e' % (eW, eb))
            check('%s bucket, eager pass %d: dX == local dX (bit-exact)' % (kind, it), bool(torch.equal(xr.grad, gx_ref)))
        if kind == 'peer':
            check('peer bucket status word clean', bucket.status() == 0)
            # the CUDA-graph step of the public API with the same bucket (device sampler: statistical check only)
            W = torch.nn.Parameter(W0.clone())
            b = torch.nn.Parameter(b0.clone())
            step = GraphedHeadStep(W, b, dt, N, 8.0, x_dtype=torch.bfloat16, grad_bucket=bucket, seed=4242 + rank)
            dist.barrier()
            for it in range(3):
                losses = step(x, labels).clone()
            torch.cuda.synchronize()
            gw = step.grad_weight.clone()
            other = gw.clone()
            dist.broadcast(other, src=0)
            check('graphed step: every rank holds the same averaged weight gradient', bool(torch.equal(gw, other)))
            check('graphed step: losses finite, gradient non-zero',
                  bool(torch.isfinite(losses).all()) and gw.abs().sum().item() > 0)
            del step
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0 if out['ok'] else 1)


if __name__ == '__main__':
    main()
