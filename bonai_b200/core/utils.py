"""multi_apply / unmap (mmdet/core/utils/misc.py:35-67)."""
from functools import partial

import torch


def multi_apply(func, *args, **kwargs):
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def unmap(data, count, inds, fill=0):
    if data.dim() == 1:
        ret = data.new_full((count,), fill)
        ret[inds.type(torch.bool)] = data
    else:
        ret = data.new_full((count,) + data.size()[1:], fill)
        ret[inds.type(torch.bool), :] = data
    return ret
