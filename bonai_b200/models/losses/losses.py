"""Loss modules under the reference's registry names (mmdet/models/losses/): CrossEntropyLoss
(cross_entropy_loss.py:128-200; softmax / sigmoid / mask variants), L1Loss / SmoothL1Loss
(smooth_l1_loss.py:45-136), FocalLoss (focal_loss.py:89-156), accuracy (accuracy.py:4-76).
Forward+reduction is one fused kernel per loss (ops/losses.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..builder import LOSSES
from ...ops import losses as K
from ...ops.focal import sigmoid_focal_loss as _sigmoid_focal_loss


def reduce_loss(loss, reduction):
    reduction_enum = F._Reduction.get_enum(reduction)
    if reduction_enum == 0:
        return loss
    if reduction_enum == 1:
        return loss.mean()
    return loss.sum()


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """mmdet/models/losses/utils.py:26-52."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        loss = reduce_loss(loss, reduction)
    else:
        if reduction == 'mean':
            loss = loss.sum() / avg_factor
        elif reduction != 'none':
            raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def _as2d(t):
    if t.dim() == 1:
        return t.contiguous().view(-1, 1)
    return t.contiguous().view(t.shape[0], -1)


def _scale(n_elems, reduction, avg_factor, loss_weight):
    if avg_factor is not None:
        if reduction != 'mean':
            raise ValueError('avg_factor can only be used with reduction="mean"')
        return loss_weight / float(avg_factor)
    if reduction == 'mean':
        return loss_weight / max(float(n_elems), 1.0)
    if reduction == 'sum':
        return loss_weight
    raise NotImplementedError('reduction="none" is not on the fused CUDA path')


def accuracy(pred, target, topk=1):
    """Top-k accuracy in percent (accuracy.py:4-48)."""
    assert isinstance(topk, (int, tuple))
    return_single = isinstance(topk, int)
    topk = (topk,) if return_single else topk
    maxk = max(topk)
    if pred.size(0) == 0:
        accu = [pred.new_tensor(0.) for _ in topk]
        return accu[0] if return_single else accu
    _, pred_label = pred.topk(maxk, dim=1)
    pred_label = pred_label.t()
    correct = pred_label.eq(target.view(1, -1).expand_as(pred_label))
    res = [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / pred.size(0))
           for k in topk]
    return res[0] if return_single else res


class Accuracy(nn.Module):
    def __init__(self, topk=(1,)):
        super().__init__()
        self.topk = topk

    def forward(self, pred, target):
        return accuracy(pred, target, self.topk)


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None,
                 loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if class_weight is not None:
            raise NotImplementedError('class_weight is not used by the LOFT config')
        self.use_sigmoid = use_sigmoid
        self.use_mask = use_mask
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = class_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if self.use_mask:
            # positional convention of the reference (fcn_mask_head.py:146):
            # loss_mask(mask_pred, mask_targets, labels) -> label=targets, weight=class labels
            assert reduction == 'mean' and avg_factor is None
            return self.loss_weight * mask_cross_entropy(cls_score, label, weight)
        if self.use_sigmoid:
            pred = _as2d(cls_score)
            C = pred.shape[1]
            if pred.shape != label.shape and label.dim() == 1 and C >= 1:
                # _expand_binary_labels (cross_entropy_loss.py:42-55)
                bin_labels = torch.zeros_like(pred)
                valid = label >= 1
                idx = torch.nonzero(valid, as_tuple=False).squeeze(1)
                if idx.numel() > 0:
                    bin_labels[idx, label[idx] - 1] = 1
                w = None if weight is None else weight.view(-1, 1).expand(-1, C).contiguous()
                label = bin_labels
                weight = w
            n = pred.numel()
            return K.elem_loss(pred, label.float().reshape(-1),
                               None if weight is None else weight.float().reshape(-1),
                               K.BCE_LOGITS, _scale(n, reduction, avg_factor, self.loss_weight))
        pred = cls_score.contiguous()
        n = pred.shape[0]
        out = K.softmax_ce(pred, label, weight, pred.shape[1],
                           _scale(n, reduction, avg_factor, self.loss_weight))
        return out[0]


def mask_cross_entropy(pred, target, label):
    """cross_entropy_loss.py:94-125."""
    P, C = pred.shape[:2]
    if C != 1:
        inds = torch.arange(P, device=pred.device)
        sl = pred[inds, label]
    else:
        sl = pred
    sl = sl.reshape(-1, 1).contiguous()
    return K.elem_loss(sl, target.reshape(-1), None, K.BCE_LOGITS, 1.0 / max(sl.numel(), 1))


@LOSSES.register_module()
class L1Loss(nn.Module):
    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        assert pred.size() == target.size() and target.numel() > 0
        p = _as2d(pred)
        return K.elem_loss(p, target.reshape(-1), None if weight is None else weight.reshape(-1),
                           K.L1, _scale(p.numel(), reduction, avg_factor, self.loss_weight))


@LOSSES.register_module()
class SmoothL1Loss(nn.Module):
    def __init__(self, beta=1.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.beta = beta
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None,
                **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        assert self.beta > 0 and pred.size() == target.size() and target.numel() > 0
        p = _as2d(pred)
        return K.elem_loss(p, target.reshape(-1), None if weight is None else weight.reshape(-1),
                           K.SMOOTH_L1, _scale(p.numel(), reduction, avg_factor, self.loss_weight),
                           beta=self.beta)


@LOSSES.register_module()
class FocalLoss(nn.Module):
    """focal_loss.py:89-156 (sigmoid only), backed by the sigmoid_focal_loss CUDA op."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, 'Only sigmoid focal loss supported now.'
        self.use_sigmoid = use_sigmoid
        self.gamma = gamma
        self.alpha = alpha
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        loss = _sigmoid_focal_loss(pred, target, self.gamma, self.alpha, None, 'none')
        if weight is not None:
            if weight.shape != loss.shape:
                if weight.size(0) == loss.size(0):
                    weight = weight.view(-1, 1)
                else:
                    assert weight.numel() == loss.numel()
                    weight = weight.view(loss.size(0), -1)
            assert weight.ndim == loss.ndim
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)
