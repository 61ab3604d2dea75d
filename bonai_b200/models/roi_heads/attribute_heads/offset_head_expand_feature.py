"""OffsetHeadExpandFeature = the FOA module (mmdet/models/roi_heads/attribute_heads/
offset_head_expand_feature.py:25-461): the RoI feature is rotated by 0/90/180/270 degrees, each
rotation runs its own 10x(3x3 conv+ReLU) stack, the shared fc1024-fc1024-fc2 maps every branch to
an offset; targets are the GT offset rotated into the branch frame.

B200 path: rotation is a rot90 permutation kernel (== affine_grid+grid_sample to 5e-7, SURVEY 2a
N11); the 40 convs are implicit-GEMM launches over [P,7,7,256] NHWC tiles (5 RoIs per 245-pixel
tile); the four branches share ONE batched pass through the FC layers ([4P,12544] x [12544,1024]);
targets use the closed form of the reference's polar rotation (verified equal, tests/)."""
import ctypes
import math

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
from ....ops import misc as M

i32 = ctypes.c_int


@HEADS.register_module()
class OffsetHeadExpandFeature(nn.Module):
    def __init__(self, roi_feat_size=7, in_channels=256, num_convs=4, num_fcs=2, reg_num=2,
                 conv_out_channels=256, fc_out_channels=1024, expand_feature_num=4,
                 share_expand_fc=False, rotations=[0, 90, 180, 270], offset_coordinate='rectangle',
                 offset_coder=dict(type='DeltaXYOffsetCoder', target_means=[0.0, 0.0],
                                   target_stds=[0.5, 0.5]),
                 reg_decoded_offset=False, conv_cfg=None, norm_cfg=None,
                 loss_offset=dict(type='MSELoss', loss_weight=1.0)):
        super().__init__()
        if not share_expand_fc:
            raise NotImplementedError('LOFT path: share_expand_fc=True '
                                      '(bonai_loft_foa_r50_fpn_basic.py:77)')
        if reg_decoded_offset or offset_coordinate != 'rectangle' or reg_num != 2:
            raise NotImplementedError('LOFT path: rectangular (x, y) offsets, encoded targets')
        for r in rotations[:expand_feature_num]:
            if r not in (0, 90, 180, 270):
                raise NotImplementedError(f'rotation angle: {r}')
        self.in_channels, self.conv_out_channels = in_channels, conv_out_channels
        self.fc_out_channels, self.offset_coordinate = fc_out_channels, offset_coordinate
        self.reg_decoded_offset, self.reg_num = reg_decoded_offset, reg_num
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.expand_feature_num, self.share_expand_fc = expand_feature_num, share_expand_fc
        self.offset_coder = build_bbox_coder(offset_coder)
        self.loss_offset = build_loss(loss_offset)
        self.rotations = rotations
        self.flips = ['h', 'v']
        self.expand_convs = nn.ModuleList()
        for _ in range(expand_feature_num):
            convs = nn.ModuleList()
            for i in range(num_convs):
                cin = in_channels if i == 0 else conv_out_channels
                convs.append(nn.Conv2d(cin, conv_out_channels, 3, padding=1))
            self.expand_convs.append(convs)
        self.roi_feat_size = _pair(roi_feat_size)
        area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.fcs = nn.ModuleList()
        for i in range(num_fcs):
            cin = conv_out_channels * area if i == 0 else fc_out_channels
            self.fcs.append(nn.Linear(cin, fc_out_channels))
        self.fc_offset = nn.Linear(fc_out_channels, reg_num)
        self.relu = nn.ReLU()

    def init_weights(self):
        for convs in self.expand_convs:
            for conv in convs:
                kaiming_init(conv)
        for fc in self.fcs:
            kaiming_init(fc, a=1, mode='fan_in', nonlinearity='leaky_relu', distribution='uniform')
        normal_init(self.fc_offset, std=0.01)

    def loft_prepare(self, store):
        self._conv_specs = [[D.ConvSpec(c.weight._loft, ksize=3, padding=1, relu=True, bias=c.bias,
                                        bias_grad=c.bias._loft.grad, store=store,
                                        premask_in=(i > 0), grad_premasked=True)
                             for i, c in enumerate(convs)] for convs in self.expand_convs]
        # the same conv layer of all branches in one grouped launch (weights are equally strided
        # in the ParamStore's flat buffers)
        self._group_specs = None
        nconv = len(self.expand_convs[0])
        if self.expand_feature_num > 1 and all(len(c) == nconv for c in self.expand_convs) and \
                self.in_channels == self.conv_out_channels:
            gs = []
            for i in range(nconv):
                convs = [self.expand_convs[b][i] for b in range(self.expand_feature_num)]
                gs.append(D.GroupedConvSpec([c.weight._loft for c in convs],
                                            [c.bias for c in convs],
                                            [c.bias._loft.grad for c in convs], relu=True,
                                            store=store, premask_in=(i > 0), grad_premasked=True))
            if all(g.uniform for g in gs):
                self._group_specs = gs
                D.link_chain(gs)
        if self._group_specs is None:
            for specs in self._conv_specs:
                D.link_chain(specs)
        area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self._fc_specs = []
        for i, fc in enumerate(self.fcs):
            wref = make_hwc_fc(store, fc, self.conv_out_channels, area) if i == 0 \
                else fc.weight._loft
            self._fc_specs.append(D.ConvSpec(wref, relu=True, bias=fc.bias,
                                             bias_grad=fc.bias._loft.grad, store=store,
                                             premask_in=(i > 0 or len(self.expand_convs[0]) > 0),
                                             grad_premasked=True))
        wref, b, gb = make_fused_head(store, [self.fc_offset], 4)
        self._head = D.ConvSpec(wref, bias=b, bias_grad=gb, round_out=False, store=store,
                                premask_in=len(self.fcs) > 0)
        D.link_chain(self._fc_specs + [self._head])

    def expand_feature(self, feature, operation_idx):
        if operation_idx >= 4:
            raise NotImplementedError
        return M.rot90(feature, self.rotations[operation_idx] // 90)

    def branch_rotations(self):
        """Quarter turns of the branches if `forward(x, expanded=...)` can take the branch-major
        rotated batch [nb*P, C, S, S] from the caller (ops.roi.take_rows_rot assembles it straight
        from the bbox head's RoI features); None if the head is not on the grouped path."""
        rots = list(self.rotations)[:self.expand_feature_num]
        if self._group_specs is None or self.expand_feature_num > 4 or not rots or rots[0] != 0 or \
                any(r % 90 for r in rots):
            return None
        return [r // 90 for r in rots]

    def forward(self, x, expanded=None):
        if x.size(0) == 0:
            return x.new_empty(x.size(0), 2 * self.expand_feature_num)
        if self._group_specs is not None:
            # branch-major batch [4P, C, 7, 7]; each layer of the 4 branches is ONE launch
            y = expanded if expanded is not None else \
                torch.cat([self.expand_feature(x, idx) for idx in range(self.expand_feature_num)], 0)
            for i, gspec in enumerate(self._group_specs):
                trig = tuple(self.expand_convs[b][i].weight for b in range(self.expand_feature_num))
                y = D.grouped_conv3x3(y, gspec, triggers=trig)
            y = D.nhwc(y).reshape(y.shape[0], -1)                # [4P, 7*7*C], branch-major
        else:
            feats = []
            for idx in range(self.expand_feature_num):
                yb = self.expand_feature(x, idx)
                for conv, spec in zip(self.expand_convs[idx], self._conv_specs[idx]):
                    yb = D.conv(yb, spec, triggers=(conv.weight, conv.bias))
                feats.append(D.nhwc(yb).reshape(yb.shape[0], -1))
            y = torch.cat(feats, 0)                              # [4P, 7*7*C], branch-major
        for fc, spec in zip(self.fcs, self._fc_specs):
            y = D.linear(y, spec, triggers=(fc.weight, fc.bias))
        fused = D.linear(y, self._head, triggers=(self.fc_offset.weight, self.fc_offset.bias))
        offsets = fused[:, :self.reg_num]
        offsets._loft_fused = fused
        return offsets

    def loss(self, offset_pred, offset_targets):
        if offset_pred.size(0) == 0:
            return dict(loss_offset=offset_pred.sum() * 0)
        fused = getattr(offset_pred, '_loft_fused', None)
        name = type(self.loss_offset).__name__
        if fused is not None and name in ('SmoothL1Loss', 'L1Loss'):
            mode, beta = (K.L1, 1.0) if name == 'L1Loss' else (K.SMOOTH_L1, self.loss_offset.beta)
            n = offset_pred.shape[0] * self.reg_num
            return dict(loss_offset=K.elem_loss(fused, offset_targets.reshape(-1), None, mode,
                                                self.loss_offset.loss_weight / n, col_off=0,
                                                ncols=self.reg_num, beta=beta))
        return dict(loss_offset=self.loss_offset(offset_pred.contiguous(), offset_targets))

    # ---- host-side restatement of the reference helpers (used by tests / small inputs) -------
    def offset_coordinate_transform(self, offset, transform_flag='xy2la'):
        if transform_flag == 'xy2la':
            ox, oy = offset
            return [math.sqrt(ox ** 2 + oy ** 2), math.atan2(oy, ox)]
        if transform_flag == 'la2xy':
            length, angle = offset
            return [length * math.cos(angle), length * math.sin(angle)]
        raise NotImplementedError

    def offset_rotate(self, offset, rotate_angle):
        offset = self.offset_coordinate_transform(offset, 'xy2la')
        offset = [offset[0], offset[1] - rotate_angle * math.pi / 180.0]
        return self.offset_coordinate_transform(offset, 'la2xy')

    def get_targets(self, sampling_results, gt_offsets, rcnn_train_cfg, concat=True):
        """offset_head_expand_feature.py:271-344 as one kernel: out is [4*P, 2], branch-major, with
        P the positives of all images in order."""
        assert concat and self.expand_feature_num == 4 and list(self.rotations) == [0, 90, 180, 270]
        props = torch.cat([r.pos_bboxes for r in sampling_results], 0).contiguous().float()
        P = props.shape[0]
        out = torch.empty((4 * P, 2), device=props.device, dtype=torch.float32)
        if P == 0:
            return out
        inds, offs, o = [], [], 0
        for r, go in zip(sampling_results, gt_offsets):
            inds.append(r.pos_assigned_gt_inds + o)
            offs.append(go.to(props.device).float())
            o += go.shape[0]
        inds = torch.cat(inds).contiguous()
        offs = torch.cat(offs, 0).contiguous()
        stds = self.offset_coder.stds
        assert all(float(m) == 0.0 for m in self.offset_coder.means)
        L.call('offset_target', L.ptr(props), L.ptr(offs), L.ptr(inds), L.ll(P), L.f32(stds[0]),
               L.f32(stds[1]), L.ptr(out), L.stream())
        return out

    def offset_fusion(self, offset_pred, model='max'):
        """offset_head_expand_feature.py:346-413 (test time)."""
        split = offset_pred.split(int(offset_pred.shape[0] / self.expand_feature_num), dim=0)
        main = split[0]
        if model != 'max' or self.expand_feature_num != 4:
            raise NotImplementedError
        vx = torch.stack([split[0][:, 0], split[1][:, 1], split[2][:, 0], split[3][:, 1]], dim=1)
        vy = torch.stack([split[0][:, 1], split[1][:, 0], split[2][:, 1], split[3][:, 0]], dim=1)
        vals = torch.stack([vx.abs().max(dim=1)[0], vy.abs().max(dim=1)[0]], dim=1)
        polarity = torch.where(main > 0, torch.ones_like(main), -torch.ones_like(main))
        return vals * polarity

    def get_offsets(self, offset_pred, det_bboxes, scale_factor, rescale, img_shape=[1024, 1024]):
        """offset_head_expand_feature.py:415-448: fusion of the four branches + decode -- one
        launch (ops.infer.offset_fusion_decode) when the head has its four rotations."""
        if offset_pred is not None:
            if self.expand_feature_num == 4 and offset_pred.is_cuda and \
                    all(float(m) == 0.0 for m in self.offset_coder.means):
                from ....ops.infer import offset_fusion_decode
                pred = offset_pred if offset_pred.stride(1) == 1 else offset_pred.contiguous()
                offsets = offset_fusion_decode(pred, det_bboxes, self.offset_coder.stds, img_shape)
            else:
                offset_pred = self.offset_fusion(offset_pred)
                offsets = self.offset_coder.decode(det_bboxes, offset_pred, max_shape=img_shape)
        else:
            offsets = torch.zeros((det_bboxes.size()[0], self.reg_num))
        if isinstance(offsets, torch.Tensor):
            offsets = offsets.cpu().numpy()
        return offsets.astype(np.float32)
