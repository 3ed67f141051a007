"""CPU, world_size 2 over gloo: the data-parallel gradient exchange (restatement of
mmdet/core/utils/dist_utils.py:9-41) -- bucket flatten / all-reduce / mean / copy-back."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from balancedgroupsoftmax_b200.dist import allreduce_flat_, allreduce_grads, flat_grad_bucket
        torch.manual_seed(100 + rank)
        w = torch.nn.Parameter(torch.zeros(12, 8))
        b = torch.nn.Parameter(torch.zeros(12))
        frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
        h = torch.nn.Parameter(torch.zeros(4, dtype=torch.float64))
        w.grad = torch.full((12, 8), float(rank + 1))
        b.grad = torch.arange(12, dtype=torch.float32) * (rank + 1)
        h.grad = torch.full((4,), 10.0 * (rank + 1), dtype=torch.float64)
        allreduce_grads([w, b, frozen, h])
        ok = torch.allclose(w.grad, torch.full((12, 8), 1.5)) and \
            torch.allclose(b.grad, torch.arange(12, dtype=torch.float32) * 1.5) and \
            torch.allclose(h.grad, torch.full((4,), 15.0, dtype=torch.float64))
        # non-coalesced path
        w.grad = torch.full((12, 8), float(rank))
        allreduce_grads([w], coalesce=False)
        ok = ok and torch.allclose(w.grad, torch.full((12, 8), 0.5))
        # flat bucket with views: grads written in place, one collective, no copies
        flat, (dW, db) = flat_grad_bucket([(12, 8), (12,)], 'cpu')
        dW.fill_(2.0 * rank)
        db.fill_(4.0 * rank)
        allreduce_flat_(flat)
        ok = ok and torch.allclose(dW, torch.full((12, 8), 1.0)) and torch.allclose(db, torch.full((12,), 2.0))
        ok = ok and dW.data_ptr() == flat.data_ptr()
        # make_grad_bucket: no NVLink peer memory under gloo -> plain bucket + the reference's collective, same result
        from balancedgroupsoftmax_b200.dist import PeerGradBucket, make_grad_bucket
        ok = ok and not PeerGradBucket.available() and bool(PeerGradBucket.why_not)   # (no CUDA / not NCCL)
        bucket, flat2, (gW, gb), exchange = make_grad_bucket([(6, 4), (6,)], 'cpu')
        gW.fill_(float(rank + 1))
        gb.fill_(10.0 * rank)
        exchange()
        ok = ok and bucket is None and torch.allclose(gW, torch.full((6, 4), 1.5)) and \
            torch.allclose(gb, torch.full((6,), 5.0)) and gW.data_ptr() == flat2.data_ptr()
        # NcclGradBucket: the bucket object the data-parallel public API takes when the peer-memory path is unavailable
        # (same interface as PeerGradBucket: views + allreduce_; its stream-forking exchange_overlapped needs a GPU)
        from balancedgroupsoftmax_b200.dist import NcclGradBucket
        nb = NcclGradBucket([(5, 3), (5,)], 'cpu')
        nb.views[0].fill_(3.0 * (rank + 1))
        nb.views[1].copy_(torch.arange(5, dtype=torch.float32) + rank)
        nb.allreduce_()
        ok = ok and torch.allclose(nb.views[0], torch.full((5, 3), 4.5)) and \
            torch.allclose(nb.views[1], torch.arange(5, dtype=torch.float32) + 0.5) and nb.status() == 0 and \
            nb.views[0].data_ptr() == nb.flat.data_ptr() and hasattr(nb, 'exchange_overlapped')
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_grads_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
