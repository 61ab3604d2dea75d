"""Fused loss bindings: sums reduced on the device into 1-element tensors; backward consumes the
upstream gradient as a device scalar (no host sync).  Semantics: weight_reduce_loss with
`avg_factor` (mmdet/models/losses/utils.py:26-52): loss = sum(l_i * w_i) * scale."""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib as L

i32 = ctypes.c_int
BCE_LOGITS, L1, SMOOTH_L1 = 0, 1, 2


class _ElemLoss(Function):
    @staticmethod
    def forward(ctx, pred, target, weight, mode, col_off, ncols, beta, scale):
        assert pred.is_contiguous() and pred.dim() == 2
        rows, ld = pred.shape
        target = target.contiguous().float()
        weight = weight.contiguous().float() if weight is not None else None
        out = torch.zeros(1, device=pred.device, dtype=torch.float32)
        L.call('elem_loss_fwd', i32(mode), L.ptr(pred), L.ll(ld), i32(col_off), i32(ncols),
               L.ll(rows), L.ptr(target), L.ptr(weight), L.f32(beta), L.f32(scale), L.ptr(out),
               L.stream())
        ctx.save_for_backward(pred, target, weight)
        ctx.meta = (mode, col_off, ncols, beta, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target, weight = ctx.saved_tensors
        mode, col_off, ncols, beta, scale = ctx.meta
        rows, ld = pred.shape
        if col_off == 0 and ncols == ld:
            dpred = torch.empty_like(pred)
        else:
            dpred = torch.zeros_like(pred)
        L.call('elem_loss_bwd', i32(mode), L.ptr(pred), L.ll(ld), i32(col_off), i32(ncols),
               L.ll(rows), L.ptr(target), L.ptr(weight), L.f32(beta), L.f32(scale),
               L.ptr(g.contiguous().float()), L.ptr(dpred), L.stream())
        return dpred, None, None, None, None, None, None, None


def elem_loss(pred2d, target, weight, mode, scale, col_off=0, ncols=None, beta=1.0):
    """sum_i w_i * f(pred_i, target_i) * scale over pred2d[:, col_off:col_off+ncols]; returns a
    1-element tensor.  target/weight are dense [rows*ncols]."""
    if ncols is None:
        ncols = pred2d.shape[1] - col_off
    if pred2d.shape[0] == 0:
        return pred2d.sum()[None] * 0
    return _ElemLoss.apply(pred2d, target, weight, int(mode), int(col_off), int(ncols), float(beta),
                           float(scale))


class _SoftmaxCE(Function):
    @staticmethod
    def forward(ctx, logits, labels, weight, C, scale):
        assert logits.is_contiguous() and logits.dim() == 2
        n, ld = logits.shape
        labels = labels.contiguous().long()
        weight = weight.contiguous().float() if weight is not None else None
        out = torch.zeros(2, device=logits.device, dtype=torch.float32)
        L.call('softmax_ce_fwd', L.ptr(logits), L.ll(ld), i32(C), L.ll(n), L.ptr(labels),
               L.ptr(weight), L.f32(scale), L.ptr(out), L.stream())
        ctx.save_for_backward(logits, labels, weight)
        ctx.meta = (C, scale)
        return out

    @staticmethod
    def backward(ctx, g):
        logits, labels, weight = ctx.saved_tensors
        C, scale = ctx.meta
        n, ld = logits.shape
        d = torch.zeros_like(logits) if C != ld else torch.empty_like(logits)
        L.call('softmax_ce_bwd', L.ptr(logits), L.ll(ld), i32(C), L.ll(n), L.ptr(labels),
               L.ptr(weight), L.f32(scale), L.ptr(g[:1].contiguous().float()), L.ptr(d), L.stream())
        return d, None, None, None, None


def softmax_ce(logits2d, labels, weight, num_classes, scale):
    """Returns a 2-vector: [sum(CE_i * w_i) * scale, #correct top-1] over logits2d[:, :num_classes]."""
    return _SoftmaxCE.apply(logits2d, labels, weight, int(num_classes), float(scale))
