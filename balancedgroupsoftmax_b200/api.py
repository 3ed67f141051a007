"""Public functional entry point of the hot path.

``bags_head_loss`` = fc_cls projection + Balanced-Group-Softmax loss with a fused
backward, i.e. reference ``head.fc_cls(x)`` followed by ``head.loss(cls_score, None,
labels, ...)`` (convfc_bbox_head.py:166 + gs_bbox_head_with0.py:147-171), as one
autograd node.  ``GSBBoxHeadWith0`` (head.py) is the module-level drop-in built on it.
"""
from __future__ import annotations

import itertools
from typing import Optional, Union

import torch

from . import ops
from .tables import GroupTables

_seed_counter = itertools.count(1)


def bags_head_loss(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], labels: torch.Tensor,
                   tables: Union[GroupTables, ops.DeviceTables], others_sample_ratio: float = 8.0,
                   compute_dtype: torch.dtype = torch.bfloat16, wmask: Optional[torch.Tensor] = None,
                   avg: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                   return_logits: bool = False, grad_bucket=None):
    """losses[G] (fp32, one per bin, already divided by the per-bin avg_factor).

    x [N,K] and weight [C,K] may be fp32 or bf16; ``compute_dtype`` selects bf16 or TF32 tensor-core
    products.  Without ``wmask`` the "others" rows are sampled on the device (``seed`` fixes the draw);
    with ``wmask`` [G,N] uint8 (e.g. recorded from the reference's numpy sampler) it is used as is.
    ``grad_bucket`` (data-parallel training; ``dist.PeerGradBucket`` / ``dist.NcclGradBucket`` over [(C,K), (C,)]): the
    backward writes dW / db into the bucket, exchanges them over NVLink as soon as they are complete and computes dX
    meanwhile; ``weight.grad`` / ``bias.grad`` receive the mean over ranks.
    """
    dt = tables if isinstance(tables, ops.DeviceTables) else ops.DeviceTables.from_tables(tables, x.device)
    if wmask is None:
        if seed is None:
            seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter)) & 0xFFFFFFFFFFFFFFFF
        wmask, avg = ops.sample_others(labels, dt, others_sample_ratio, seed)
    elif avg is None:
        avg = ops.mask_avg(wmask)
    logits = None
    if return_logits:
        logits = torch.empty((x.shape[0], weight.shape[0]), dtype=torch.float32, device=x.device)
    loss = ops.GroupSoftmaxFunction.apply(x, weight, bias, labels, dt, wmask, avg, compute_dtype, logits, grad_bucket)
    return (loss, logits) if return_logits else loss


class GraphedHeadStep(object):
    """One training step of the hot path, captured once as a CUDA graph and replayed.

        sample "others" masks -> fc_cls + BAGS loss (GroupSoftmaxFunction) -> sum_g c_g * loss_g -> backward
        [-> gradient exchange]

    i.e. what ``head.fc_cls(x)``, ``head.loss(...)``, ``parse_losses`` and ``loss.backward()`` do per iteration in the
    reference (convfc_bbox_head.py:166, gs_bbox_head_with0.py:147-171, apis/train.py:17-34, dist_utils.py:53), with
    the same autograd Function as the eager API underneath -- the standard ``torch.cuda.graph`` whole-step capture.
    The per-step host cost drops from a few hundred microseconds of Python / autograd / allocator work to one graph
    launch, which is what a 43 us GPU step needs.

    Static tensors (filled / read by the caller between replays):
        ``x`` [N,K], ``labels`` [N] int64        inputs (e.g. the destination of the H2D copies)
        ``losses`` [G] fp32                      per-bin losses of the last replay
        ``grad_weight``, ``grad_bias``, ``grad_x``   gradients of the last replay (also ``weight.grad`` / ``bias.grad``)
    The sampler's seed advances on the device after every replay, so replays draw different "others" subsets.
    """

    def __init__(self, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter],
                 tables: Union[GroupTables, ops.DeviceTables], n_rois: int, others_sample_ratio: float = 8.0,
                 compute_dtype: torch.dtype = torch.bfloat16, x_dtype: Optional[torch.dtype] = None,
                 loss_weights: Optional[torch.Tensor] = None, need_dx: bool = True, seed: Optional[int] = None,
                 exchange=None, stream: Optional[torch.cuda.Stream] = None, warmup: int = 2, grad_bucket=None):
        dev = weight.device
        if dev.type != 'cuda':
            raise ops.nat.BagsNativeError('GraphedHeadStep runs on a B200 GPU only (weight is on %s)' % dev)
        self.weight, self.bias = weight, bias
        self.dt = tables if isinstance(tables, ops.DeviceTables) else ops.DeviceTables.from_tables(tables, dev)
        self.ratio = float(others_sample_ratio)
        self.compute_dtype = compute_dtype
        self.loss_weights = None if loss_weights is None else loss_weights.to(dev, torch.float32)
        self.exchange = exchange
        # data-parallel: dW / db are written into this bucket (dist.PeerGradBucket / NcclGradBucket), exchanged as soon as
        # they are complete and while dX is computed; weight.grad / bias.grad become the bucket's views (mean over ranks)
        self.grad_bucket = grad_bucket
        self.seed = int(seed if seed is not None else
                        (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter))) & 0xFFFFFFFFFFFFFFFF
        K = weight.shape[1]
        self.x = torch.zeros((n_rois, K), dtype=(x_dtype or weight.dtype), device=dev, requires_grad=need_dx)
        self.labels = torch.zeros((n_rois,), dtype=torch.int64, device=dev)
        self.seed_step = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(max(int(warmup), 1)):       # also lets the library set up its per-stream state
                self._clear_grads()
                self._body()
            self.stream.synchronize()
            self._clear_grads()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.losses = self._body()
            self.seed_step.zero_()
        self.grad_weight = weight.grad
        self.grad_bias = None if bias is None else bias.grad
        self.grad_x = self.x.grad
        torch.cuda.current_stream(dev).wait_stream(self.stream)

    def _clear_grads(self):
        self.weight.grad = None
        if self.bias is not None:
            self.bias.grad = None
        self.x.grad = None

    def _body(self):
        wmask, avg = ops.sample_others(self.labels, self.dt, self.ratio, self.seed, seed_step=self.seed_step)
        losses = ops.GroupSoftmaxFunction.apply(self.x, self.weight, self.bias, self.labels, self.dt, wmask, avg,
                                                self.compute_dtype, None, self.grad_bucket)
        total = losses.sum() if self.loss_weights is None else (losses * self.loss_weights).sum()
        total.backward()
        if self.exchange is not None:
            self.exchange()
        # last node of the graph: never the stream predecessor of the sampler, which reads it before its PDL wait
        self.seed_step.add_(1)
        return losses.detach()

    def replay(self) -> torch.Tensor:
        """Launch the captured step on the current stream; returns the static ``losses`` tensor."""
        self.graph.replay()
        return self.losses

    def __call__(self, x: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None:
            with torch.no_grad():
                self.x.copy_(x, non_blocking=True)
        if labels is not None:
            self.labels.copy_(labels, non_blocking=True)
        return self.replay()
