"""FCNMaskHead (mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:18-308): 4x(3x3 conv + ReLU)
@14x14 -> 2x2/2 deconv + ReLU -> 1x1 logits @28x28.  The convs are implicit-GEMM launches, the
deconv is one GEMM whose epilogue scatters the four sub-pixels, the logits conv is padded to a
4-wide row; targets come from the uint8 mask-target sampler; the loss is one fused BCE reduction."""
import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from ..builder_alias import HEADS, build_loss
from ...init_utils import ConvModule
from .... import _lib as L
from ....core import mask_target
from ....engine import Packed, WeightRef
from ....ops import dense as D
from ....ops import losses as K

i32 = ctypes.c_int


@HEADS.register_module()
class FCNMaskHead(nn.Module):
    def __init__(self, num_convs=4, roi_feat_size=14, in_channels=256, conv_kernel_size=3,
                 conv_out_channels=256, num_classes=80, class_agnostic=False,
                 upsample_cfg=dict(type='deconv', scale_factor=2), conv_cfg=None, norm_cfg=None,
                 loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0)):
        super().__init__()
        self.upsample_cfg = dict(upsample_cfg)
        if self.upsample_cfg['type'] != 'deconv' or self.upsample_cfg.get('scale_factor', 2) != 2 \
                or conv_kernel_size != 3 or norm_cfg is not None or conv_cfg is not None:
            raise NotImplementedError('LOFT path: 3x3 convs + 2x deconv mask head')
        self.num_convs = num_convs
        self.roi_feat_size = _pair(roi_feat_size)
        self.in_channels, self.conv_kernel_size = in_channels, conv_kernel_size
        self.conv_out_channels = conv_out_channels
        self.upsample_method = 'deconv'
        self.scale_factor = 2
        self.num_classes, self.class_agnostic = num_classes, class_agnostic
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.fp16_enabled = False
        self.loss_mask = build_loss(loss_mask)
        self.convs = nn.ModuleList()
        for i in range(num_convs):
            cin = in_channels if i == 0 else conv_out_channels
            self.convs.append(ConvModule(cin, conv_out_channels, 3, padding=1))
        up_in = conv_out_channels if num_convs > 0 else in_channels
        self.upsample = nn.ConvTranspose2d(up_in, conv_out_channels, 2, stride=2)
        out_channels = 1 if class_agnostic else num_classes
        self.conv_logits = nn.Conv2d(conv_out_channels, out_channels, 1)
        self.relu = nn.ReLU(inplace=True)
        self.debug_imgs = None

    def init_weights(self):
        for m in [self.upsample, self.conv_logits]:
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            nn.init.constant_(m.bias, 0)

    def loft_prepare(self, store):
        dev = store.device
        self._conv_specs = []
        for i, cm in enumerate(self.convs):
            c = cm.conv
            self._conv_specs.append(D.ConvSpec(c.weight._loft, ksize=3, padding=1, relu=True,
                                               bias=c.bias, bias_grad=c.bias._loft.grad,
                                               store=store, premask_in=(i > 0),
                                               grad_premasked=True))
        up = self.upsample                       # weight [Cin, Co, 2, 2]
        Ci, Co = up.weight.shape[:2]
        w = torch.zeros((4 * Co, Ci), device=dev)
        gw = torch.zeros((4 * Co, Ci), device=dev)
        b4 = torch.zeros((4 * Co,), device=dev)

        # The ParamStore keeps 4-D weights channels_last: [Cin, Co, 2, 2] is physically
        # [ci][i][j][co], so the packed [(i,j,co)][ci] operand is one 2-D transpose away.
        def build():
            st = L.stream()
            L.call('permute_acb', L.ptr(up.weight._loft.w), L.ptr(w), i32(1), i32(Ci), i32(4 * Co),
                   i32(0), i32(0), st)
            L.call('copy2d', L.ptr(up.bias), L.ll(0), L.ptr(b4), L.ll(Co), L.ll(4), i32(Co), i32(0),
                   i32(0), st)

        def scatter():
            L.call('permute_acb', L.ptr(gw), L.ptr(up.weight._loft.grad), i32(1), i32(4 * Co),
                   i32(Ci), i32(1), i32(0), L.stream())

        store.add_packed(Packed(w, b4, gw, None, build, scatter))
        self._up_spec = D.ConvSpec(WeightRef(w, gw), relu=True, bias=b4,
                                   bias_grad=up.bias._loft.grad, store=store,
                                   premask_in=len(self.convs) > 0, grad_premasked=True)
        lg = self.conv_logits
        n_out = lg.weight.shape[0]
        width = (n_out + 3) // 4 * 4
        lw = torch.zeros((width, Co), device=dev)
        lb = torch.zeros((width,), device=dev)
        lgw = torch.zeros((width, Co), device=dev)
        lgb = torch.zeros((width,), device=dev)

        def cp(src, dst, r, c, acc):
            L.call('copy2d', L.ptr(src), L.ll(c), L.ptr(dst), L.ll(c), L.ll(r), i32(c), i32(acc),
                   i32(0), L.stream())

        def build2():
            cp(lg.weight._loft.w, lw, n_out, Co, 0)
            cp(lg.bias, lb, 1, n_out, 0)

        def scatter2():
            cp(lgw, lg.weight._loft.grad, n_out, Co, 1)
            cp(lgb, lg.bias._loft.grad, 1, n_out, 1)

        store.add_packed(Packed(lw, lb, lgw, lgb, build2, scatter2))
        self._logit_spec = D.ConvSpec(WeightRef(lw, lgw), ksize=1, bias=lb, bias_grad=lgb,
                                      round_out=False, store=store, premask_in=True)
        self._n_out = n_out
        self._logit_spec.n_out = n_out
        D.link_chain(self._conv_specs + [self._up_spec, self._logit_spec])

    def forward(self, x):
        for cm, spec in zip(self.convs, self._conv_specs):
            x = D.conv(x, spec, triggers=(cm.conv.weight, cm.conv.bias))
        if D.deconv_logits_ok(self._up_spec, self._logit_spec, x):
            # upsample + logits with the deconv output left in its GEMM layout (no pixel-shuffle
            # store, no space-to-depth copy of its gradient)
            fused = D.deconv_logits(x, self._up_spec, self._logit_spec,
                                    triggers=(self.upsample.weight, self.upsample.bias,
                                              self.conv_logits.weight))
        else:
            x = D.deconv2x2(x, self._up_spec, triggers=(self.upsample.weight, self.upsample.bias))
            if D.narrow_head_ok(self._logit_spec, x):   # 256 -> 1 channel: HBM-bound, no GEMM tile
                fused = D.narrow_head(x, self._logit_spec, triggers=(self.conv_logits.weight,))
            else:
                fused = D.conv(x, self._logit_spec, triggers=(self.conv_logits.weight,))
        mask_pred = fused[:, :self._n_out]
        mask_pred._loft_fused = fused
        return mask_pred

    def get_targets(self, sampling_results, gt_masks, rcnn_train_cfg):
        pos_proposals = [res.pos_bboxes for res in sampling_results]
        pos_assigned_gt_inds = [res.pos_assigned_gt_inds for res in sampling_results]
        return mask_target(pos_proposals, pos_assigned_gt_inds, gt_masks, rcnn_train_cfg)

    def loss(self, mask_pred, mask_targets, labels):
        loss = dict()
        if mask_pred.size(0) == 0:
            loss['loss_mask'] = mask_pred.sum() * 0
            return loss
        fused = getattr(mask_pred, '_loft_fused', None)
        if fused is not None and (self.class_agnostic or self._n_out == 1):
            # mask_cross_entropy on pred[:, 0]: column 0 of the 4-wide fused rows, mean reduction
            out2d = fused.permute(0, 2, 3, 1).reshape(-1, fused.shape[1])
            n = out2d.shape[0]
            loss['loss_mask'] = K.elem_loss(out2d, mask_targets.reshape(-1), None, K.BCE_LOGITS,
                                            self.loss_mask.loss_weight / n, col_off=0, ncols=1)
        elif self.class_agnostic:
            loss['loss_mask'] = self.loss_mask(mask_pred, mask_targets, torch.zeros_like(labels))
        else:
            loss['loss_mask'] = self.loss_mask(mask_pred, mask_targets, labels)
        return loss

    def get_seg_masks(self, mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape,
                      scale_factor, rescale, to_numpy=True):
        """Test-time mask paste with the interface of the reference's FCNMaskHead.get_seg_masks
        (fcn_mask_head.py:151-237): `cls_segms[label]` = list of [img_h, img_w] bitmaps, one per
        detection.  The work is one zero-fill and one launch of the paste kernel
        (ops.infer.paste_masks: only each detection's box window is resampled), then ONE
        device->host copy; `to_numpy=False` keeps the [N, img_h, img_w] bitmaps on the device
        (returned as the third element of a (cls_segms, labels, masks) tuple)."""
        from ....ops.infer import paste_masks
        if not isinstance(mask_pred, torch.Tensor):
            mask_pred = det_bboxes.new_tensor(mask_pred)
        if rescale:
            img_h, img_w = (int(v) for v in ori_shape[:2])
        else:
            img_h = int(np.round(ori_shape[0] * scale_factor))
            img_w = int(np.round(ori_shape[1] * scale_factor))
            scale_factor = 1.0
        if not isinstance(scale_factor, (float, torch.Tensor)):
            scale_factor = det_bboxes.new_tensor(scale_factor)
        bboxes = det_bboxes[:, :4] / scale_factor
        N = mask_pred.shape[0]
        fused = getattr(mask_pred, '_loft_fused', None)
        if (self.class_agnostic or self._n_out == 1) and fused is not None:
            # channel 0 of the fused [N, 4, M, M] head output (NHWC storage): pixel stride 4
            logits = fused.permute(0, 2, 3, 1)[..., 0]
        elif self.class_agnostic:
            logits = mask_pred[:, 0]
        else:
            logits = mask_pred[torch.arange(N, device=mask_pred.device), det_labels]
        if logits.stride(1) != logits.shape[2] * logits.stride(2):
            logits = logits.contiguous()
        masks = paste_masks(logits, bboxes, img_h, img_w, float(rcnn_test_cfg.mask_thr_binary))
        cls_segms = [[] for _ in range(self.num_classes)]
        if not to_numpy:
            return cls_segms, det_labels, masks
        host = masks.cpu().numpy()
        for i, lab in enumerate(det_labels.tolist()):
            cls_segms[lab].append(host[i])
        return cls_segms
