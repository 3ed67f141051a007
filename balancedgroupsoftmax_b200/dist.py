"""Data-parallel gradient exchange for the head: one NCCL all-reduce over NVLink.

Restates mmdet/core/utils/dist_utils.py:9-58 (``_allreduce_coalesced`` /
``allreduce_grads``): flatten all trainable grads by dtype into one bucket,
all-reduce, divide by world size, copy back.  Differences that keep the result
identical but remove passes: NCCL's native AVG op replaces the separate ``div_``,
and callers that allocate their grads as views of one flat bucket
(``flat_grad_bucket``) skip the flatten / copy-back entirely.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def _avg_op():
    return dist.ReduceOp.AVG if dist.get_backend() == 'nccl' else None


def allreduce_flat_(flat: torch.Tensor, async_op: bool = False):
    """In-place mean over ranks of one flat bucket."""
    op = _avg_op()
    if op is not None:
        return dist.all_reduce(flat, op=op, async_op=async_op)
    work = dist.all_reduce(flat, async_op=async_op)   # gloo has no AVG
    if async_op:
        work.wait()
    flat.div_(dist.get_world_size())
    return None


def allreduce_grads(params: Iterable[torch.nn.Parameter], coalesce: bool = True) -> None:
    grads = [p.grad.data for p in params if p.requires_grad and p.grad is not None]
    if not grads:
        return
    if not coalesce:
        for g in grads:
            allreduce_flat_(g)
        return
    buckets = OrderedDict()
    for g in grads:
        buckets.setdefault(g.type(), []).append(g)
    for bucket in buckets.values():
        if len(bucket) == 1 and bucket[0].is_contiguous():
            allreduce_flat_(bucket[0].view(-1))
            continue
        flat = _flatten_dense_tensors(bucket)
        allreduce_flat_(flat)
        for g, synced in zip(bucket, _unflatten_dense_tensors(flat, bucket)):
            g.copy_(synced)


def flat_grad_bucket(shapes: List[Tuple[int, ...]], device, dtype=torch.float32):
    """One flat buffer + per-tensor views (write grads straight into the bucket)."""
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    flat = torch.zeros(sum(sizes), device=device, dtype=dtype)
    views, off = [], 0
    for s, n in zip(shapes, sizes):
        views.append(flat[off:off + n].view(*s))
        off += n
    return flat, views
