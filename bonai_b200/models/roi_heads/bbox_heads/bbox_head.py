"""Shared2FCBBoxHead (mmdet/models/roi_heads/bbox_heads/bbox_head.py:12-334,
convfc_bbox_head.py:9-189): flatten -> fc1024+ReLU -> fc1024+ReLU -> fc_cls / fc_reg, as three
tcgen05 GEMMs (fc_cls and fc_reg fused into one 8-wide head).  RoI features are NHWC, so the first
FC reads a re-ordered copy of its weight ((C,H,W) -> (H,W,C) input order), refreshed per step."""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..builder_alias import HEADS, build_loss
from ...init_utils import normal_init, xavier_init
from .... import _lib as L
from ....core import build_bbox_coder, multi_apply
from ....engine import Packed, WeightRef
from ....ops import dense as D
from ....ops import losses as K
from ...losses import accuracy

i32 = ctypes.c_int


def make_hwc_fc(store, fc, C, HW):
    """Packed copy of an fc weight [O, C*HW] with its input axis re-ordered to (HW, C)."""
    O = fc.weight.shape[0]
    dev = store.device
    w = torch.zeros((O, HW * C), device=dev)
    gw = torch.zeros((O, HW * C), device=dev)

    def build():
        L.call('permute_acb', L.ptr(fc.weight._loft.w), L.ptr(w), i32(O), i32(C), i32(HW), i32(0),
               i32(0), L.stream())

    def scatter():
        L.call('permute_acb', L.ptr(gw), L.ptr(fc.weight._loft.grad), i32(O), i32(HW), i32(C),
               i32(1), i32(0), L.stream())

    store.add_packed(Packed(w, None, gw, None, build, scatter))
    return WeightRef(w, gw)


def make_fused_head(store, fcs, width):
    """Several narrow nn.Linear heads stacked into one [width, K] weight (zero padded)."""
    dev = store.device
    Kdim = fcs[0].weight.shape[1]
    w = torch.zeros((width, Kdim), device=dev)
    b = torch.zeros((width,), device=dev)
    gw = torch.zeros((width, Kdim), device=dev)
    gb = torch.zeros((width,), device=dev)
    rows = [fc.weight.shape[0] for fc in fcs]
    assert sum(rows) <= width

    def cp(src, dst, r, c, acc):
        L.call('copy2d', L.ptr(src), L.ll(c), L.ptr(dst), L.ll(c), L.ll(r), i32(c), i32(acc), i32(0),
               L.stream())

    def build():
        o = 0
        for fc, r in zip(fcs, rows):
            cp(fc.weight._loft.w, w[o:], r, Kdim, 0)
            cp(fc.bias, b[o:], 1, r, 0)
            o += r

    def scatter():
        o = 0
        for fc, r in zip(fcs, rows):
            cp(gw[o:], fc.weight._loft.grad, r, Kdim, 1)
            cp(gb[o:], fc.bias._loft.grad, 1, r, 1)
            o += r

    store.add_packed(Packed(w, b, gw, gb, build, scatter))
    return WeightRef(w, gw), b, gb


@HEADS.register_module()
class BBoxHead(nn.Module):
    def __init__(self, with_avg_pool=False, with_cls=True, with_reg=True, roi_feat_size=7,
                 in_channels=256, num_classes=80,
                 bbox_coder=dict(type='DeltaXYWHBBoxCoder', target_means=[0., 0., 0., 0.],
                                 target_stds=[0.1, 0.1, 0.2, 0.2]),
                 reg_class_agnostic=False, reg_decoded_bbox=False,
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0)):
        super().__init__()
        assert with_cls and with_reg and not with_avg_pool and not reg_decoded_bbox
        from torch.nn.modules.utils import _pair
        self.with_avg_pool, self.with_cls, self.with_reg = with_avg_pool, with_cls, with_reg
        self.roi_feat_size = _pair(roi_feat_size)
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.reg_class_agnostic, self.reg_decoded_bbox = reg_class_agnostic, reg_decoded_bbox
        self.fp16_enabled = False
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)

    def _get_target_single(self, pos_bboxes, neg_bboxes, pos_gt_bboxes, pos_gt_labels, cfg):
        num_pos, num_neg = pos_bboxes.size(0), neg_bboxes.size(0)
        num_samples = num_pos + num_neg
        labels = pos_bboxes.new_full((num_samples,), self.num_classes, dtype=torch.long)
        label_weights = pos_bboxes.new_zeros(num_samples)
        bbox_targets = pos_bboxes.new_zeros(num_samples, 4)
        bbox_weights = pos_bboxes.new_zeros(num_samples, 4)
        if num_pos > 0:
            labels[:num_pos] = pos_gt_labels
            label_weights[:num_pos] = 1.0 if cfg.pos_weight <= 0 else cfg.pos_weight
            bbox_targets[:num_pos, :] = self.bbox_coder.encode(pos_bboxes, pos_gt_bboxes)
            bbox_weights[:num_pos, :] = 1
        if num_neg > 0:
            label_weights[-num_neg:] = 1.0
        return labels, label_weights, bbox_targets, bbox_weights

    def get_targets(self, sampling_results, gt_bboxes, gt_labels, rcnn_train_cfg, concat=True):
        r = multi_apply(self._get_target_single, [s.pos_bboxes for s in sampling_results],
                        [s.neg_bboxes for s in sampling_results],
                        [s.pos_gt_bboxes for s in sampling_results],
                        [s.pos_gt_labels for s in sampling_results], cfg=rcnn_train_cfg)
        if concat:
            r = tuple(torch.cat(t, 0) for t in r)
        return r

    def loss(self, cls_score, bbox_pred, rois, labels, label_weights, bbox_targets, bbox_weights,
             reduction_override=None):
        """bbox_head.py:140-185.  Every sampled RoI has weight 1, so avg_factor is the row count
        (the reference obtains it with a `.item()` sync, bbox_head.py:152)."""
        losses = dict()
        fused = getattr(cls_score, '_loft_fused', None)
        if fused is None:
            raise L.LoftError('BBoxHead.loss expects the fused head output of forward()')
        n = fused.shape[0]
        nc = self.num_classes
        if n > 0:
            out = K.softmax_ce(fused, labels, label_weights, nc + 1,
                               self.loss_cls.loss_weight / max(float(n), 1.0))
            losses['loss_cls'] = out[0]
            losses['acc'] = (out[1:2] * (100.0 / n)).detach()
            if self.reg_class_agnostic or nc == 1:
                # one foreground class: the class slice of bbox_pred is columns [0,4) for every
                # positive, negatives carry zero weight (== bbox_pred.view(N,-1,4)[pos, label])
                mode, beta = (K.L1, 1.0) if type(self.loss_bbox).__name__ == 'L1Loss' else \
                    (K.SMOOTH_L1, self.loss_bbox.beta)
                losses['loss_bbox'] = K.elem_loss(
                    fused, bbox_targets.reshape(-1), bbox_weights.reshape(-1), mode,
                    self.loss_bbox.loss_weight / float(bbox_targets.size(0)), col_off=nc + 1,
                    ncols=4, beta=beta)
            else:
                raise NotImplementedError('LOFT path: a single foreground class (building)')
        return losses

    def get_bboxes(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False,
                   cfg=None):
        from ....core.post_processing import multiclass_nms
        if isinstance(cls_score, list):
            cls_score = sum(cls_score) / float(len(cls_score))
        scores = F.softmax(cls_score, dim=1) if cls_score is not None else None
        if bbox_pred is not None:
            bboxes = self.bbox_coder.decode(rois[:, 1:], bbox_pred, max_shape=img_shape)
        else:
            bboxes = rois[:, 1:].clone()
            if img_shape is not None:
                bboxes[:, [0, 2]].clamp_(min=0, max=img_shape[1])
                bboxes[:, [1, 3]].clamp_(min=0, max=img_shape[0])
        if rescale and bboxes.size(0) > 0:
            if isinstance(scale_factor, float):
                bboxes /= scale_factor
            else:
                scale_factor = bboxes.new_tensor(scale_factor)
                bboxes = (bboxes.view(bboxes.size(0), -1, 4) / scale_factor).view(
                    bboxes.size()[0], -1)
        if cfg is None:
            return bboxes, scores
        return multiclass_nms(bboxes, scores, cfg.score_thr, cfg.nms, cfg.max_per_img)


@HEADS.register_module()
class ConvFCBBoxHead(BBoxHead):
    def __init__(self, num_shared_convs=0, num_shared_fcs=0, num_cls_convs=0, num_cls_fcs=0,
                 num_reg_convs=0, num_reg_fcs=0, conv_out_channels=256, fc_out_channels=1024,
                 conv_cfg=None, norm_cfg=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if num_shared_convs or num_cls_convs or num_cls_fcs or num_reg_convs or num_reg_fcs:
            raise NotImplementedError('LOFT path: shared FC layers only (Shared2FCBBoxHead)')
        assert num_shared_fcs > 0
        self.num_shared_fcs = num_shared_fcs
        self.fc_out_channels = fc_out_channels
        self.shared_convs = nn.ModuleList()
        self.shared_fcs = nn.ModuleList()
        last = self.in_channels * self.roi_feat_area
        for i in range(num_shared_fcs):
            self.shared_fcs.append(nn.Linear(last, fc_out_channels))
            last = fc_out_channels
        self.shared_out_channels = last
        self.cls_convs, self.cls_fcs = nn.ModuleList(), nn.ModuleList()
        self.reg_convs, self.reg_fcs = nn.ModuleList(), nn.ModuleList()
        self.relu = nn.ReLU(inplace=True)
        self.fc_cls = nn.Linear(last, self.num_classes + 1)
        out_dim_reg = 4 if self.reg_class_agnostic else 4 * self.num_classes
        self.fc_reg = nn.Linear(last, out_dim_reg)

    def init_weights(self):
        nn.init.normal_(self.fc_cls.weight, 0, 0.01)
        nn.init.constant_(self.fc_cls.bias, 0)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001)
        nn.init.constant_(self.fc_reg.bias, 0)
        for m in self.shared_fcs:
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)

    def loft_prepare(self, store):
        self._specs = []
        for i, fc in enumerate(self.shared_fcs):
            wref = make_hwc_fc(store, fc, self.in_channels, self.roi_feat_area) if i == 0 \
                else fc.weight._loft
            self._specs.append(D.ConvSpec(wref, relu=True, bias=fc.bias,
                                          bias_grad=fc.bias._loft.grad, store=store,
                                          premask_in=(i > 0), grad_premasked=True))
        n_out = self.fc_cls.weight.shape[0] + self.fc_reg.weight.shape[0]
        width = (n_out + 3) // 4 * 4
        wref, b, gb = make_fused_head(store, [self.fc_cls, self.fc_reg], width)
        self._head = D.ConvSpec(wref, bias=b, bias_grad=gb, round_out=False, store=store,
                                premask_in=True)
        D.link_chain(self._specs + [self._head])

    def forward(self, x):
        # x: [K, C, 7, 7] with NHWC storage -> [K, 7*7*C]
        xf = D.nhwc(x).reshape(x.shape[0], -1)
        for fc, spec in zip(self.shared_fcs, self._specs):
            xf = D.linear(xf, spec, triggers=(fc.weight, fc.bias))
        fused = D.linear(xf, self._head, triggers=(self.fc_cls.weight, self.fc_reg.weight))
        nc1 = self.fc_cls.weight.shape[0]
        cls_score = fused[:, :nc1]
        bbox_pred = fused[:, nc1:nc1 + self.fc_reg.weight.shape[0]]
        cls_score._loft_fused = fused
        bbox_pred._loft_fused = fused
        return cls_score, bbox_pred


@HEADS.register_module()
class Shared2FCBBoxHead(ConvFCBBoxHead):
    def __init__(self, fc_out_channels=1024, *args, **kwargs):
        super().__init__(num_shared_convs=0, num_shared_fcs=2, num_cls_convs=0, num_cls_fcs=0,
                         num_reg_convs=0, num_reg_fcs=0, fc_out_channels=fc_out_channels, *args,
                         **kwargs)
