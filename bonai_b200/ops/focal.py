"""SigmoidFocalLoss binding (mirror of mmcv.ops.sigmoid_focal_loss, reference call site
mmdet/models/losses/focal_loss.py:3,70).  Not executed by the loft_foa config (RPN uses sigmoid
CE) but part of the op surface BASELINE.json names."""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib as L

i32 = ctypes.c_int


class _SigmoidFocalLoss(Function):
    @staticmethod
    def forward(ctx, input, target, gamma, alpha, weight):
        input = input.contiguous().float()
        target = target.contiguous().long()
        n, C = input.shape
        out = torch.empty_like(input)
        w = weight.contiguous().float() if weight is not None else None
        L.call('sigmoid_focal_loss_fwd', L.ptr(input), L.ptr(target), L.ptr(w), L.ll(n), i32(C),
               L.f32(gamma), L.f32(alpha), L.ptr(out), L.stream())
        ctx.save_for_backward(input, target, w)
        ctx.ga = (gamma, alpha)
        return out

    @staticmethod
    def backward(ctx, dloss):
        input, target, w = ctx.saved_tensors
        gamma, alpha = ctx.ga
        n, C = input.shape
        dx = torch.empty_like(input)
        L.call('sigmoid_focal_loss_bwd', L.ptr(input), L.ptr(target), L.ptr(w), L.ll(n), i32(C),
               L.f32(gamma), L.f32(alpha), L.ptr(dloss.contiguous()), L.ptr(dx), L.stream())
        return dx, None, None, None, None


def sigmoid_focal_loss(input, target, gamma=2.0, alpha=0.25, weight=None, reduction='mean'):
    """input [N,C] logits, target [N] int64 in [0,C] (C = background), weight [N] or None."""
    loss = _SigmoidFocalLoss.apply(input, target, float(gamma), float(alpha), weight)
    if reduction == 'none':
        return loss
    if reduction == 'sum':
        return loss.sum()
    if reduction == 'mean':
        return loss.sum() / input.size(0)
    raise ValueError(reduction)
