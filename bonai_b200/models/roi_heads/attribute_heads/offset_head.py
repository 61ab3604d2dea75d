"""OffsetHead: the single-branch (no rotation) variant
(mmdet/models/roi_heads/attribute_heads/offset_head.py:24-265)."""
import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from ..builder_alias import HEADS, build_loss
from ..bbox_heads.bbox_head import make_fused_head, make_hwc_fc
from ...init_utils import kaiming_init, normal_init
from .... import _lib as L
from ....core import build_bbox_coder
from ....ops import dense as D
from ....ops import losses as K


@HEADS.register_module()
class OffsetHead(nn.Module):
    def __init__(self, roi_feat_size=7, in_channels=256, num_convs=4, num_fcs=2, reg_num=2,
                 conv_out_channels=256, fc_out_channels=1024, offset_coordinate='rectangle',
                 offset_coder=dict(type='DeltaXYOffsetCoder', target_means=[0.0, 0.0],
                                   target_stds=[0.5, 0.5]),
                 reg_decoded_offset=False, conv_cfg=None, norm_cfg=None,
                 loss_offset=dict(type='SmoothL1Loss', loss_weight=1.0)):
        super().__init__()
        if reg_decoded_offset or offset_coordinate != 'rectangle' or reg_num != 2:
            raise NotImplementedError('LOFT path: rectangular (x, y) offsets, encoded targets')
        self.in_channels, self.conv_out_channels = in_channels, conv_out_channels
        self.fc_out_channels, self.reg_num = fc_out_channels, reg_num
        self.offset_coordinate, self.reg_decoded_offset = offset_coordinate, reg_decoded_offset
        self.offset_coder = build_bbox_coder(offset_coder)
        self.loss_offset = build_loss(loss_offset)
        self.convs = nn.ModuleList()
        for i in range(num_convs):
            cin = in_channels if i == 0 else conv_out_channels
            self.convs.append(nn.Conv2d(cin, conv_out_channels, 3, padding=1))
        self.roi_feat_size = _pair(roi_feat_size)
        area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.fcs = nn.ModuleList()
        for i in range(num_fcs):
            cin = conv_out_channels * area if i == 0 else fc_out_channels
            self.fcs.append(nn.Linear(cin, fc_out_channels))
        self.fc_offset = nn.Linear(fc_out_channels, reg_num)
        self.relu = nn.ReLU()

    def init_weights(self):
        for conv in self.convs:
            kaiming_init(conv)
        for fc in self.fcs:
            kaiming_init(fc, a=1, mode='fan_in', nonlinearity='leaky_relu', distribution='uniform')
        normal_init(self.fc_offset, std=0.01)

    def loft_prepare(self, store):
        self._conv_specs = [D.ConvSpec(c.weight._loft, ksize=3, padding=1, relu=True, bias=c.bias,
                                       bias_grad=c.bias._loft.grad, store=store)
                            for c in self.convs]
        area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self._fc_specs = []
        for i, fc in enumerate(self.fcs):
            wref = make_hwc_fc(store, fc, self.conv_out_channels, area) if i == 0 \
                else fc.weight._loft
            self._fc_specs.append(D.ConvSpec(wref, relu=True, bias=fc.bias,
                                             bias_grad=fc.bias._loft.grad, store=store))
        wref, b, gb = make_fused_head(store, [self.fc_offset], 4)
        self._head = D.ConvSpec(wref, bias=b, bias_grad=gb, round_out=False, store=store)

    def forward(self, x):
        if x.size(0) == 0:
            return x.new_empty(x.size(0), 2)
        for conv, spec in zip(self.convs, self._conv_specs):
            x = D.conv(x, spec, triggers=(conv.weight, conv.bias))
        y = D.nhwc(x).reshape(x.shape[0], -1)
        for fc, spec in zip(self.fcs, self._fc_specs):
            y = D.linear(y, spec, triggers=(fc.weight, fc.bias))
        fused = D.linear(y, self._head, triggers=(self.fc_offset.weight, self.fc_offset.bias))
        offsets = fused[:, :self.reg_num]
        offsets._loft_fused = fused
        return offsets

    def loss(self, offset_pred, offset_targets):
        if offset_pred.size(0) == 0:
            return dict(loss_offset=offset_pred.sum() * 0)
        return dict(loss_offset=self.loss_offset(offset_pred.contiguous(), offset_targets))

    def get_targets(self, sampling_results, gt_offsets, rcnn_train_cfg, concat=True):
        """offset_head.py:120-160: encode(pos_proposals, gt_offsets[assigned])."""
        out = []
        for r, go in zip(sampling_results, gt_offsets):
            if r.pos_bboxes.size(0) == 0:
                out.append(r.pos_bboxes.new_zeros((0, 2)))
            else:
                out.append(self.offset_coder.encode(r.pos_bboxes,
                                                    go.to(r.pos_bboxes.device)[
                                                        r.pos_assigned_gt_inds]))
        return torch.cat(out, 0) if concat else out

    def get_offsets(self, offset_pred, det_bboxes, scale_factor, rescale, img_shape=[1024, 1024]):
        if offset_pred is not None:
            offsets = self.offset_coder.decode(det_bboxes, offset_pred, max_shape=img_shape)
        else:
            offsets = torch.zeros((det_bboxes.size()[0], self.reg_num))
        if isinstance(offsets, torch.Tensor):
            offsets = offsets.cpu().numpy()
        return offsets.astype(np.float32)
