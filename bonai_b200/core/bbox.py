"""Box-side building blocks of the path with the reference's names and registries
(mmdet/core/bbox/): MaxIoUAssigner (assigners/max_iou_assigner.py:60-212), AssignResult
(assign_result.py), RandomSampler / SamplingResult (samplers/), DeltaXYWHBBoxCoder
(coder/delta_xywh_bbox_coder.py), DeltaXYOffsetCoder (coder/delta_xy_offset_coder.py),
BboxOverlaps2D (iou_calculators/iou2d_calculator.py), bbox2roi / bbox2result (transforms.py).
The O(G x n) IoU + assignment runs in one fused CUDA pass (no IoU matrix, no per-GT Python loop)."""
import ctypes

import numpy as np
import torch

from .. import _lib as L
from ..registry import Registry, build_from_cfg

BBOX_ASSIGNERS = Registry('bbox_assigner')
BBOX_SAMPLERS = Registry('bbox_sampler')
BBOX_CODERS = Registry('bbox_coder')
IOU_CALCULATORS = Registry('IoU calculator')

i32 = ctypes.c_int


def build_assigner(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_ASSIGNERS, default_args)


def build_sampler(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_SAMPLERS, default_args)


def build_bbox_coder(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_CODERS, default_args)


def build_iou_calculator(cfg, default_args=None):
    return build_from_cfg(cfg, IOU_CALCULATORS, default_args)


# ------------------------------------------------------------------------------ IoU
def bbox_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False, eps=1e-6):
    """Dense IoU matrix (host-side API parity; the training path uses the fused assign kernel)."""
    assert mode in ['iou', 'iof'] and not is_aligned
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if rows * cols == 0:
        return bboxes1.new(rows, cols)
    lt = torch.max(bboxes1[:, None, :2], bboxes2[:, :2])
    rb = torch.min(bboxes1[:, None, 2:], bboxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[:, :, 0] * wh[:, :, 1]
    area1 = (bboxes1[:, 2] - bboxes1[:, 0]) * (bboxes1[:, 3] - bboxes1[:, 1])
    if mode == 'iou':
        area2 = (bboxes2[:, 2] - bboxes2[:, 0]) * (bboxes2[:, 3] - bboxes2[:, 1])
        union = area1[:, None] + area2 - overlap
    else:
        union = area1[:, None]
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


@IOU_CALCULATORS.register_module()
class BboxOverlaps2D:
    def __call__(self, bboxes1, bboxes2, mode='iou', is_aligned=False):
        assert bboxes1.size(-1) in [0, 4, 5] and bboxes2.size(-1) in [0, 4, 5]
        return bbox_overlaps(bboxes1[..., :4], bboxes2[..., :4], mode, is_aligned)

    def __repr__(self):
        return self.__class__.__name__ + '()'


# ------------------------------------------------------------------------------ assign
class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts = num_gts
        self.gt_inds = gt_inds
        self.max_overlaps = max_overlaps
        self.labels = labels

    @property
    def num_preds(self):
        return len(self.gt_inds)

    def add_gt_(self, gt_labels):
        self_inds = torch.arange(1, len(gt_labels) + 1, dtype=torch.long,
                                 device=gt_labels.device)
        self.gt_inds = torch.cat([self_inds, self.gt_inds])
        self.max_overlaps = torch.cat([self.max_overlaps.new_ones(len(gt_labels)),
                                       self.max_overlaps])
        if self.labels is not None:
            self.labels = torch.cat([gt_labels, self.labels])


@BBOX_ASSIGNERS.register_module()
class MaxIoUAssigner:
    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True,
                 ignore_iof_thr=-1, ignore_wrt_candidates=True, match_low_quality=True,
                 gpu_assign_thr=-1, iou_calculator=dict(type='BboxOverlaps2D')):
        if isinstance(neg_iou_thr, (tuple, list)):
            raise NotImplementedError('tuple neg_iou_thr is not used by the LOFT config')
        if not gt_max_assign_all:
            raise NotImplementedError('gt_max_assign_all=False is not used by the LOFT config')
        self.pos_iou_thr = pos_iou_thr
        self.neg_iou_thr = neg_iou_thr
        self.min_pos_iou = min_pos_iou
        self.gt_max_assign_all = gt_max_assign_all
        self.ignore_iof_thr = ignore_iof_thr
        self.ignore_wrt_candidates = ignore_wrt_candidates
        self.gpu_assign_thr = gpu_assign_thr      # kept for config parity; never falls back to CPU
        self.match_low_quality = match_low_quality
        self.iou_calculator = build_iou_calculator(iou_calculator)

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        if gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0 and \
                self.ignore_iof_thr > 0:
            raise NotImplementedError('gt_bboxes_ignore is not used by the LOFT config '
                                      '(ignore_iof_thr=-1)')
        n, G = bboxes.size(0), gt_bboxes.size(0)
        dev = bboxes.device
        boxes = bboxes[:, :4].contiguous().float()
        gts = gt_bboxes[:, :4].contiguous().float()
        gt_inds = torch.empty((n,), dtype=torch.long, device=dev)
        max_ov = torch.zeros((n,), dtype=torch.float32, device=dev)
        if n > 0:
            if G == 0:
                gt_inds.zero_()
            else:
                ws_bytes = int(L.lib().loft_iou_assign_workspace(L.ll(n), i32(G)))
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
                L.call('iou_assign', L.ptr(boxes), L.ll(n), L.ptr(gts), i32(G),
                       L.f32(self.pos_iou_thr), L.f32(self.neg_iou_thr), L.f32(self.min_pos_iou),
                       i32(self.match_low_quality), L.ptr(gt_inds), L.ptr(max_ov), L.ptr(ws),
                       ctypes.c_size_t(ws_bytes), L.stream())
        labels = None
        if gt_labels is not None:
            labels = gt_inds.new_full((n,), -1)
            if G > 0 and n > 0:
                pos = gt_inds > 0
                labels = torch.where(pos, gt_labels[(gt_inds - 1).clamp(min=0)], labels)
        return AssignResult(G, gt_inds, max_ov, labels=labels)


# ------------------------------------------------------------------------------ sample
class SamplingResult:
    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags):
        self.pos_inds = pos_inds
        self.neg_inds = neg_inds
        self.pos_bboxes = bboxes[pos_inds]
        self.neg_bboxes = bboxes[neg_inds]
        self.pos_is_gt = gt_flags[pos_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        if gt_bboxes.numel() == 0:
            assert self.pos_assigned_gt_inds.numel() == 0
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, 4)
        else:
            if len(gt_bboxes.shape) < 2:
                gt_bboxes = gt_bboxes.view(-1, 4)
            self.pos_gt_bboxes = gt_bboxes[self.pos_assigned_gt_inds, :]
        self.pos_gt_labels = assign_result.labels[pos_inds] \
            if assign_result.labels is not None else None

    @classmethod
    def from_fused(cls, sel, boxes, gt_idx, is_gt, n_pos, n_neg, gt_bboxes, gt_labels):
        """From the outputs of `loft_rcnn_sample` for one image (rows: positives, then negatives):
        every field is a view of those rows, except the two gathers from the gt arrays."""
        r = cls.__new__(cls)
        n = n_pos + n_neg
        r.pos_inds, r.neg_inds = sel[:n_pos], sel[n_pos:n]
        r.pos_bboxes, r.neg_bboxes = boxes[:n_pos], boxes[n_pos:n]
        r._bboxes = boxes[:n]
        r.pos_is_gt = is_gt[:n_pos]
        r.num_gts = gt_bboxes.shape[0]
        r.pos_assigned_gt_inds = gt_idx[:n_pos]
        if gt_bboxes.numel() == 0:
            r.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, 4)
        else:
            r.pos_gt_bboxes = gt_bboxes.view(-1, 4)[r.pos_assigned_gt_inds, :]
        r.pos_gt_labels = gt_labels[r.pos_assigned_gt_inds] if gt_labels is not None else None
        return r

    @property
    def bboxes(self):
        b = getattr(self, '_bboxes', None)
        return b if b is not None else torch.cat([self.pos_bboxes, self.neg_bboxes])


@BBOX_SAMPLERS.register_module()
class RandomSampler:
    """BaseSampler.sample + RandomSampler (samplers/base_sampler.py:34-101,
    random_sampler.py:31-75).  All random draws go through `random_choice` (parity tests replace
    that method on a sampler INSTANCE to inject the oracle's draws: CUDA and CPU Philox streams
    differ)."""

    def __init__(self, num, pos_fraction, neg_pos_ub=-1, add_gt_as_proposals=True, **kwargs):
        self.num = num
        self.pos_fraction = pos_fraction
        self.neg_pos_ub = neg_pos_ub
        self.add_gt_as_proposals = add_gt_as_proposals
        self.pos_sampler = self
        self.neg_sampler = self

    def random_choice(self, gallery, num):
        assert len(gallery) >= num
        is_tensor = isinstance(gallery, torch.Tensor)
        if not is_tensor:
            gallery = torch.tensor(gallery, dtype=torch.long,
                                   device=torch.cuda.current_device())
        perm = torch.randperm(gallery.numel(), device=gallery.device)[:num]
        rand_inds = gallery[perm]
        if not is_tensor:
            rand_inds = rand_inds.cpu().numpy()
        return rand_inds

    def _sample_pos(self, assign_result, num_expected, **kwargs):
        pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False)
        if pos_inds.numel() != 0:
            pos_inds = pos_inds.squeeze(1)
        if pos_inds.numel() <= num_expected:
            return pos_inds
        return self.random_choice(pos_inds, num_expected)

    def _sample_neg(self, assign_result, num_expected, **kwargs):
        neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False)
        if neg_inds.numel() != 0:
            neg_inds = neg_inds.squeeze(1)
        if len(neg_inds) <= num_expected:
            return neg_inds
        return self.random_choice(neg_inds, num_expected)

    def sample(self, assign_result, bboxes, gt_bboxes, gt_labels=None, **kwargs):
        if len(bboxes.shape) < 2:
            bboxes = bboxes[None, :]
        bboxes = bboxes[:, :4]
        gt_flags = bboxes.new_zeros((bboxes.shape[0],), dtype=torch.uint8)
        if self.add_gt_as_proposals and len(gt_bboxes) > 0:
            if gt_labels is None:
                raise ValueError('gt_labels must be given when add_gt_as_proposals is True')
            bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
            assign_result.add_gt_(gt_labels)
            gt_ones = bboxes.new_ones(gt_bboxes.shape[0], dtype=torch.uint8)
            gt_flags = torch.cat([gt_ones, gt_flags])
        num_expected_pos = int(self.num * self.pos_fraction)
        pos_inds = self._sample_pos(assign_result, num_expected_pos, bboxes=bboxes, **kwargs)
        pos_inds = pos_inds.unique()
        num_sampled_pos = pos_inds.numel()
        num_expected_neg = self.num - num_sampled_pos
        if self.neg_pos_ub >= 0:
            _pos = max(1, num_sampled_pos)
            num_expected_neg = min(int(self.neg_pos_ub * _pos), num_expected_neg)
        neg_inds = self._sample_neg(assign_result, num_expected_neg, bboxes=bboxes, **kwargs)
        neg_inds = neg_inds.unique()
        return SamplingResult(pos_inds, neg_inds, bboxes, gt_bboxes, assign_result, gt_flags)


# ------------------------------------------------------------------------------ coders
def bbox2delta(proposals, gt, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.)):
    """delta_xywh_bbox_coder.py:74-116 -- one fused kernel (means must be 0 on the LOFT path)."""
    assert proposals.size() == gt.size()
    assert all(float(m) == 0.0 for m in means), 'non-zero target_means are not used by LOFT'
    n = proposals.shape[0]
    out = torch.empty((n, 4), device=proposals.device, dtype=torch.float32)
    if n == 0:
        return out
    L.call('bbox_encode', L.ptr(proposals.contiguous().float()), L.ptr(gt.contiguous().float()),
           L.ll(n), L.f32(stds[0]), L.f32(stds[1]), L.f32(stds[2]), L.f32(stds[3]), L.ptr(out),
           L.stream())
    return out


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None,
               wh_ratio_clip=16 / 1000):
    """delta_xywh_bbox_coder.py:119-197 (test-time decode; the RPN training path decodes inside
    the fused rpn_decode kernel)."""
    means = deltas.new_tensor(means).repeat(1, deltas.size(1) // 4)
    stds = deltas.new_tensor(stds).repeat(1, deltas.size(1) // 4)
    d = deltas * stds + means
    dx, dy, dw, dh = d[:, 0::4], d[:, 1::4], d[:, 2::4], d[:, 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dh)
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + pw * dx, py + ph * dy
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1])
        y1 = y1.clamp(min=0, max=max_shape[0])
        x2 = x2.clamp(min=0, max=max_shape[1])
        y2 = y2.clamp(min=0, max=max_shape[0])
    return torch.stack([x1, y1, x2, y2], dim=-1).view_as(deltas)


@BBOX_CODERS.register_module()
class DeltaXYWHBBoxCoder:
    def __init__(self, target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.)):
        self.means = target_means
        self.stds = target_stds

    def encode(self, bboxes, gt_bboxes):
        assert bboxes.size(0) == gt_bboxes.size(0)
        assert bboxes.size(-1) == gt_bboxes.size(-1) == 4
        return bbox2delta(bboxes, gt_bboxes, self.means, self.stds)

    def decode(self, bboxes, pred_bboxes, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_bboxes.size(0) == bboxes.size(0)
        return delta2bbox(bboxes, pred_bboxes, self.means, self.stds, max_shape, wh_ratio_clip)


def offset2delta(proposals, gt, means=(0., 0.), stds=(0.5, 0.5)):
    """delta_xy_offset_coder.py:46-65."""
    assert proposals.size()[0] == gt.size()[0]
    proposals, gt = proposals.float(), gt.float()
    pw = proposals[..., 2] - proposals[..., 0]
    ph = proposals[..., 3] - proposals[..., 1]
    deltas = torch.stack([gt[..., 0] / pw, gt[..., 1] / ph], dim=-1)
    means = deltas.new_tensor(means).unsqueeze(0)
    stds = deltas.new_tensor(stds).unsqueeze(0)
    return deltas.sub_(means).div_(stds)


def delta2offset(rois, deltas, means=(0., 0.), stds=(1., 1.), max_shape=None,
                 wh_ratio_clip=16 / 1000):
    """delta_xy_offset_coder.py:67-88."""
    means = deltas.new_tensor(means).repeat(1, deltas.size(1) // 2)
    stds = deltas.new_tensor(stds).repeat(1, deltas.size(1) // 2)
    d = deltas * stds + means
    dx, dy = d[:, 0::2], d[:, 1::2]
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dx)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dy)
    gx, gy = pw * dx, ph * dy
    if max_shape is not None:
        gx = gx.clamp(min=-max_shape[1], max=max_shape[1])
        gy = gy.clamp(min=-max_shape[0], max=max_shape[0])
    return torch.stack([gx, gy], dim=-1).view_as(deltas)


@BBOX_CODERS.register_module()
class DeltaXYOffsetCoder:
    def __init__(self, target_means=(0., 0.), target_stds=(0.5, 0.5)):
        self.means = target_means
        self.stds = target_stds

    def encode(self, bboxes, gt_offsets):
        assert bboxes.size(0) == gt_offsets.size(0)
        assert gt_offsets.size(-1) == 2
        return offset2delta(bboxes, gt_offsets, self.means, self.stds)

    def decode(self, bboxes, pred_offsets, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_offsets.size(0) == bboxes.size(0)
        return delta2offset(bboxes, pred_offsets, self.means, self.stds, max_shape, wh_ratio_clip)


# ------------------------------------------------------------------------------ transforms
def bbox2roi(bbox_list):
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def roi2bbox(rois):
    bbox_list = []
    img_ids = torch.unique(rois[:, 0].cpu(), sorted=True)
    for img_id in img_ids:
        inds = (rois[:, 0] == img_id.item())
        bbox_list.append(rois[inds, 1:])
    return bbox_list


def bbox2result(bboxes, labels, num_classes):
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    if isinstance(bboxes, torch.Tensor):
        bboxes = bboxes.cpu().numpy()
        labels = labels.cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]
