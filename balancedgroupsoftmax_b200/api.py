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
                   return_logits: bool = False):
    """losses[G] (fp32, one per bin, already divided by the per-bin avg_factor).

    x [N,K] and weight [C,K] may be fp32 or bf16; ``compute_dtype`` selects bf16 or TF32 tensor-core
    products.  Without ``wmask`` the "others" rows are sampled on the device (``seed`` fixes the draw);
    with ``wmask`` [G,N] uint8 (e.g. recorded from the reference's numpy sampler) it is used as is.
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
    loss = ops.GroupSoftmaxFunction.apply(x, weight, bias, labels, dt, wmask, avg, compute_dtype, logits)
    return (loss, logits) if return_logits else loss
