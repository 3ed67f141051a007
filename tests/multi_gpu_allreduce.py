"""Multi-GPU probe of the NVLink peer-memory gradient exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_allreduce.py [--json out.json]

Checks bags_grad_allreduce (PeerGradBucket) against NCCL all_reduce(AVG) on the fc_cls bucket shape
(1236 x 1024 + 1236 fp32, mmdet/core/utils/dist_utils.py:9-41 semantics), eagerly, back to back (flag re-use) and
from a CUDA graph, with both transports (NVLS multimem / plain peer loads+stores), then times both.
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
    from balancedgroupsoftmax_b200.dist import PeerGradBucket
    out = {'world': world, 'checks': [], 'ok': True}
    shapes = [(1236, 1024), (1236,)]
    assert PeerGradBucket.available(), 'peer path unavailable: ' + PeerGradBucket.why_not

    def check(name, cond, detail=''):
        t = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        good = bool(t.item())
        out['checks'].append({'name': name, 'ok': good, 'detail': detail})
        out['ok'] = out['ok'] and good
        if rank == 0:
            print('%-58s %s %s' % (name, 'ok' if good else 'FAIL', detail), flush=True)

    for mode in ('multimem', 'peer'):
        if mode == 'peer':
            os.environ['BAGS_AR_NO_MULTIMEM'] = '1'
        else:
            os.environ.pop('BAGS_AR_NO_MULTIMEM', None)
        bucket = PeerGradBucket(shapes, dev)
        if mode == 'multimem' and not bucket.mc_ptr:
            check('multimem: multicast mapping present', False, 'multicast_ptr == 0 (no NVLS on this box)')
            continue
        tag = '%s[%s]' % (mode, bucket.transport)
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            worst = 0.0
            for it in range(8):          # back to back: the flag words must return to zero every time
                src = torch.randn(bucket.numel, device=dev, generator=g)
                bucket.flat.copy_(src)
                ref = src.clone()
                dist.all_reduce(ref, op=dist.ReduceOp.AVG)
                bucket.allreduce_()
                stream.synchronize()
                err = (bucket.flat - ref).abs().max().item()
                worst = max(worst, err)
            check('%s eager x8 == NCCL AVG (max abs err)' % tag, worst <= 1e-6, '%.2e' % worst)
            # all ranks bit-identical
            mine = bucket.flat.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            check('%s ranks bit-identical' % tag, bool(torch.equal(mine, other)))
            check('%s status word clean' % tag, bucket.status() == 0)
            # views alias the bucket
            check('%s views alias bucket' % tag, bucket.views[0].data_ptr() == bucket.flat.data_ptr())
            # CUDA graph: fill (rank+1) -> allreduce, 4 times per graph, replay 3 times
            fill = torch.full((bucket.numel,), float(rank + 1), device=dev)
            graph = torch.cuda.CUDAGraph()
            stream.synchronize()
            with torch.cuda.graph(graph, stream=stream):
                for _ in range(4):
                    bucket.flat.copy_(fill)
                    bucket.allreduce_()
            for _ in range(3):
                graph.replay()
            stream.synchronize()
            expect = sum(range(1, world + 1)) / world
            err = (bucket.flat - expect).abs().max().item()
            check('%s graph replay (4 exchanges x 3 replays)' % tag, err <= 1e-6, '%.2e' % err)
            # timing: 50 exchanges per graph
            tg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tg, stream=stream):
                for _ in range(50):
                    bucket.allreduce_()
            tg.replay()
            stream.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(4):
                tg.replay()
            b.record(stream)
            stream.synchronize()
            us = a.elapsed_time(b) / 200 * 1e3
            t = torch.tensor([us], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out['us_per_exchange_' + mode] = float(t.item())
            if rank == 0:
                print('%s: %.2f us per exchange of %.2f MB' % (tag, float(t.item()), bucket.numel * 4 / 1e6), flush=True)
        del graph, tg

    quick = '--quick' in sys.argv
    if quick:
        if rank == 0:
            print(json.dumps(out), flush=True)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0 if out['ok'] else 1)
    # the reference-shaped entry point (dist_utils.py:31-41): grads of a parameter list, through the peer bucket
    os.environ.pop('BAGS_AR_NO_MULTIMEM', None)
    from balancedgroupsoftmax_b200.dist import allreduce_grads, _peer_buckets
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    params = [torch.nn.Parameter(torch.zeros(1236, 1024, device=dev)), torch.nn.Parameter(torch.zeros(1236, device=dev)),
              torch.nn.Parameter(torch.zeros(7, device=dev, dtype=torch.float64)),
              torch.nn.Parameter(torch.zeros(3, device=dev), requires_grad=False)]
    for p_ in params[:3]:
        p_.grad = torch.randn(p_.shape, device=dev, generator=g, dtype=p_.dtype)
    refs = []
    for p_ in params[:3]:
        r_ = p_.grad.clone()
        dist.all_reduce(r_)
        refs.append(r_ / world)
    allreduce_grads(params)
    torch.cuda.synchronize()
    err = max(((p_.grad - r_).abs().max().item() for p_, r_ in zip(params[:3], refs)))
    used = any(b is not None for b in _peer_buckets.values())
    check('allreduce_grads(params) via peer bucket == NCCL mean', err <= 1e-6 and used, 'err %.2e, peer bucket used: %s' % (err, used))

    # NCCL reference timing (graph of 50 all-reduces on the same bucket size)
    flat = torch.zeros(sum(torch.Size(s).numel() for s in shapes), device=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(3):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        stream.synchronize()
        ng = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ng, stream=stream):
            for _ in range(50):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        ng.replay()
        stream.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(4):
            ng.replay()
        b.record(stream)
        stream.synchronize()
        us = a.elapsed_time(b) / 200 * 1e3
        t = torch.tensor([us], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out['us_per_exchange_nccl'] = float(t.item())
        if rank == 0:
            print('nccl: %.2f us per all_reduce(AVG)' % float(t.item()), flush=True)
    ng = None
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
        if '--json' in sys.argv:
            with open(sys.argv[sys.argv.index('--json') + 1], 'w') as f:
                json.dump(out, f, indent=1)
    sys.stdout.flush()
    os._exit(0 if out['ok'] else 1)


if __name__ == '__main__':
    main()
