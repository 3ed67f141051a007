"""``force_fp32`` / ``auto_fp16`` method decorators keyed on ``module.fp16_enabled``.

Same observable behaviour as mmdet/core/fp16/decorators.py:9-160: a no-op unless
the owning nn.Module has ``fp16_enabled = True``; then the named tensor
arguments (all when ``apply_to`` is None) are cast half->float (force_fp32) or
float->half (auto_fp16), recursing into lists/tuples/dicts
(mmdet/core/fp16/utils.py:7-23).  ``apply_to`` may be a bare string, in which
case membership is a substring test -- the reference relies on that
(gs_bbox_head_with0.py:239 passes ``apply_to=('cls_score')``).
"""
from __future__ import annotations

import functools
from inspect import getfullargspec

import torch


def cast_tensor_type(inputs, src_type, dst_type):
    if isinstance(inputs, torch.Tensor):
        return inputs.to(dst_type) if inputs.dtype == src_type else inputs
    if isinstance(inputs, (str, bytes)):
        return inputs
    if isinstance(inputs, dict):
        return type(inputs)({k: cast_tensor_type(v, src_type, dst_type) for k, v in inputs.items()})
    if isinstance(inputs, (list, tuple)):
        return type(inputs)(cast_tensor_type(v, src_type, dst_type) for v in inputs)
    return inputs


def _casting_decorator(kind, src, dst, apply_to, cast_output):

    def wrapper(old_func):
        spec = getfullargspec(old_func)

        @functools.wraps(old_func)
        def new_func(*args, **kwargs):
            if not isinstance(args[0], torch.nn.Module):
                raise TypeError('@%s can only be used to decorate the method of nn.Module' % kind)
            if not getattr(args[0], 'fp16_enabled', False):
                return old_func(*args, **kwargs)
            names = spec.args if apply_to is None else apply_to
            new_args = [cast_tensor_type(a, src, dst) if n in names else a
                        for n, a in zip(spec.args[:len(args)], args)]
            new_kwargs = {k: (cast_tensor_type(v, src, dst) if k in names else v) for k, v in kwargs.items()}
            out = old_func(*new_args, **new_kwargs)
            if cast_output:
                out = cast_tensor_type(out, dst, src)
            return out

        return new_func

    return wrapper


def auto_fp16(apply_to=None, out_fp32=False):
    return _casting_decorator('auto_fp16', torch.float, torch.half, apply_to, out_fp32)


def force_fp32(apply_to=None, out_fp16=False):
    return _casting_decorator('force_fp32', torch.half, torch.float, apply_to, out_fp16)
