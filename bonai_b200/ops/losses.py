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


class RpnLevel(ctypes.Structure):
    """Mirror of ``loft_rpn_level_t``."""
    _fields_ = [('out', ctypes.c_void_p), ('labels', ctypes.c_void_p), ('label_w', ctypes.c_void_p),
                ('bbox_t', ctypes.c_void_p), ('bbox_w', ctypes.c_void_p), ('grad', ctypes.c_void_p),
                ('rows', ctypes.c_longlong)]


def rpn_loss_fused(outs2d, targets, num_anchors, mode_bbox, beta, cls_scale, bbox_scale,
                   grads=None, denom=None):
    """AnchorHead.loss over all levels in ONE launch (anchor_head.py:382-497).  outs2d: per level
    the fused head output [rows, ld]; targets: per level (labels, label_weights, bbox_targets,
    bbox_weights) flat; grads: per level a buffer shaped like the output that receives
    d(sum of all RPN loss terms)/d(output) in full (the total loss is the plain sum of its terms,
    detectors/base.py:175-208), or None.  Returns a [2 * n_levels] tensor: per-level cls sums, then
    per-level bbox sums (already scaled).  `denom`: optional 1-element device tensor both scales
    are divided by (the avg_factor `rpn_targets` leaves on the device).  No autograd: the caller
    owns the backward."""
    n = len(outs2d)
    ld = outs2d[0].shape[1]
    arr = (RpnLevel * n)()
    keep = []
    for l, (o, (lab, lw, bt, bw)) in enumerate(zip(outs2d, targets)):
        assert o.is_contiguous() and o.shape[1] == ld
        ts = [t.contiguous().float() for t in (lab, lw, bt, bw)]
        keep.append(ts)
        arr[l].out = o.data_ptr()
        arr[l].labels, arr[l].label_w, arr[l].bbox_t, arr[l].bbox_w = (t.data_ptr() for t in ts)
        if grads is not None:
            assert grads[l].is_contiguous() and grads[l].numel() == o.numel()
            arr[l].grad = grads[l].data_ptr()
        else:
            arr[l].grad = None
        arr[l].rows = o.shape[0]
    sums = torch.empty(2 * n, device=outs2d[0].device, dtype=torch.float32)
    L.call('rpn_loss_fused', arr, i32(n), i32(num_anchors), i32(ld), i32(mode_bbox), L.f32(beta),
           L.f32(cls_scale), L.f32(bbox_scale), L.ptr(denom), L.ptr(sums), L.stream())
    return sums


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
