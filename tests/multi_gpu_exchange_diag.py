"""Multi-GPU diagnosis of the gradient exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tests/multi_gpu_exchange_diag.py [--json out.json]

For the fc_cls bucket (1236 x 1024 + 1236 fp32 = 5.07 MB; mmdet/core/utils/dist_utils.py:9-41 semantics):
  * bit-exactness of bags_grad_allreduce against dist.all_reduce(AVG) on RANDOM gradients (8 back-to-back exchanges),
    for every transport / grid size timed below;
  * the exchange in isolation (CUDA graph of 20 exchanges, CUDA events, max over ranks) per transport and grid size,
    next to NCCL;
  * where the time goes inside the kernel: %globaltimer stamps of block 0 on every rank
    (launch -> local grads visible -> barrier 1 -> reduce + broadcast -> barrier 2).
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
    from balancedgroupsoftmax_b200 import _native
    from balancedgroupsoftmax_b200.dist import PeerGradBucket
    lib = _native.lib()
    out = {'world': world, 'ok': True, 'configs': []}
    shapes = [(1236, 1024), (1236,)]
    assert PeerGradBucket.available(), 'peer path unavailable: ' + PeerGradBucket.why_not
    ROWS = 4096 + 320
    timing = torch.zeros(ROWS * 8, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def gather_max(v):
        t = torch.tensor([float(v)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_all(vals):
        t = torch.tensor([float(v) for v in vals], device=dev)
        outl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [[round(float(x), 2) for x in o.tolist()] for o in outl]

    grids = [int(a) for a in os.environ.get('DIAG_GRIDS', '16,48,148,320').split(',')]
    configs = [('multimem', g) for g in grids] + [('peer', g) for g in grids[1:]]
    for mode, blocks in configs:
        if mode == 'peer':
            os.environ['BAGS_AR_NO_MULTIMEM'] = '1'
        else:
            os.environ.pop('BAGS_AR_NO_MULTIMEM', None)
        lib.bags_reload_env()
        bucket = PeerGradBucket(shapes, dev, max_blocks=blocks)
        if mode == 'multimem' and not bucket.mc_ptr:
            if rank == 0:
                print('no multicast mapping on this box: skipping multimem', flush=True)
            continue
        rec = {'transport': bucket.transport, 'max_blocks': blocks}
        g = torch.Generator(device=dev).manual_seed(4242 + 17 * rank)
        with torch.cuda.stream(stream):
            # ---- random-gradient bit-exactness vs NCCL AVG
            worst = 0.0
            for it in range(8):
                src = torch.randn(bucket.numel, device=dev, generator=g) * (1.0 + it)
                bucket.flat.copy_(src)
                ref = src.clone()
                dist.all_reduce(ref, op=dist.ReduceOp.AVG)
                bucket.allreduce_()
                stream.synchronize()
                worst = max(worst, (bucket.flat - ref).abs().max().item())
            mine = bucket.flat.clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            same = bool(torch.equal(mine, other))
            rec['max_abs_err_vs_nccl_avg'] = gather_max(worst)
            t = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            rec['ranks_bit_identical'] = bool(t.item())
            rec['status_clean'] = bucket.status() == 0
            # NCCL's own summation order differs from rank order beyond 2 ranks: compare to fp32 rounding, not to 0
            tol = 0.0 if world == 2 else 4e-6 * 8
            if not (rec['max_abs_err_vs_nccl_avg'] <= tol + 1e-12 and rec['ranks_bit_identical'] and rec['status_clean']):
                out['ok'] = False
            # ---- isolated timing: 20 exchanges per graph
            tg = torch.cuda.CUDAGraph()
            stream.synchronize()
            with torch.cuda.graph(tg, stream=stream):
                for _ in range(20):
                    bucket.allreduce_()
            tg.replay()
            stream.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5):
                tg.replay()
            b.record(stream)
            stream.synchronize()
            rec['us_per_exchange'] = gather_max(a.elapsed_time(b) / 100 * 1e3)
            # ---- in-kernel phases (block 0 of every rank), eager launches, barrier in between so that ranks start together
            lib.bags_debug_set_timing(timing.data_ptr())
            phases = []
            for it in range(6):
                dist.barrier()
                torch.cuda.synchronize()
                timing.zero_()
                bucket.allreduce_()
                stream.synchronize()
                row = timing[4096 * 8: 4096 * 8 + 8].tolist()
                lastb = timing[(4096 + min(blocks, 38)) * 8:(4096 + min(blocks, 38)) * 8 + 8].tolist()
                if it >= 2:
                    phases.append([(row[i + 1] - row[i]) / 1e3 for i in range(4)] + [(row[4] - row[0]) / 1e3])
            lib.bags_debug_set_timing(None)
            med = [sorted(p[i] for p in phases)[len(phases) // 2] for i in range(5)]
            rec['phases_us_per_rank[launch->grads,barrier1,data,barrier2,total]'] = gather_all(med)
        del tg
        out['configs'].append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del bucket
    os.environ.pop('BAGS_AR_NO_MULTIMEM', None)
    lib.bags_reload_env()

    # NCCL reference timing (graph of 20 all-reduces on the same bucket size)
    flat = torch.zeros(sum(torch.Size(s).numel() for s in shapes), device=dev)
    with torch.cuda.stream(stream):
        for _ in range(3):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        stream.synchronize()
        ng = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ng, stream=stream):
            for _ in range(20):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        ng.replay()
        stream.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            ng.replay()
        b.record(stream)
        stream.synchronize()
        out['us_per_exchange_nccl'] = gather_max(a.elapsed_time(b) / 100 * 1e3)
    ng = None
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(json.dumps({'world': world, 'ok': out['ok'], 'us_per_exchange_nccl': out['us_per_exchange_nccl']}), flush=True)
        if '--json' in sys.argv:
            with open(sys.argv[sys.argv.index('--json') + 1], 'w') as f:
                json.dump(out, f, indent=1)
    sys.stdout.flush()
    os._exit(0 if out['ok'] else 1)


if __name__ == '__main__':
    main()
