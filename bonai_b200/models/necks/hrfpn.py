"""HRFPN neck (mmdet/models/necks/hrfpn.py:12-102; HRNetV2p): the four HRNet branches are
bilinearly upsampled to the finest resolution and concatenated (32+64+128+256 = 480 channels), a
1x1 conv reduces them to `out_channels`, average pooling by 2^i builds the pyramid, and one 3x3
conv per level produces the outputs.  The reduction and the five output convs are fused tcgen05
launches (bias in the epilogue); the resampling / pooling are bandwidth-bound ATen ops."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..backbones.hrnet import _RoundTF32
from ..builder import NECKS
from ..init_utils import ConvModule
from ...ops import dense as D


@NECKS.register_module()
class HRFPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs=5, pooling_type='AVG', conv_cfg=None,
                 norm_cfg=None, with_cp=False, stride=1):
        super().__init__()
        assert isinstance(in_channels, list)
        if conv_cfg is not None or norm_cfg is not None or stride != 1:
            raise NotImplementedError('LOFT path: plain HRFPN (no norm / custom conv, stride 1)')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs, self.with_cp = len(in_channels), num_outs, with_cp
        self.reduction_conv = ConvModule(sum(in_channels), out_channels, 1, act_cfg=None)
        self.fpn_convs = nn.ModuleList(
            ConvModule(out_channels, out_channels, 3, padding=1, stride=stride, act_cfg=None)
            for _ in range(num_outs))
        self.pooling = F.max_pool2d if pooling_type == 'MAX' else F.avg_pool2d

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):          # caffe2_xavier_init = kaiming_uniform(a=1, fan_in)
                nn.init.kaiming_uniform_(m.weight, a=1, mode='fan_in', nonlinearity='leaky_relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def loft_prepare(self, store):
        c = self.reduction_conv.conv
        self._red = D.ConvSpec(c.weight._loft, ksize=1, bias=c.bias, bias_grad=c.bias._loft.grad,
                               store=store)
        self._out = [D.ConvSpec(m.conv.weight._loft, ksize=3, padding=1, bias=m.conv.bias,
                                bias_grad=m.conv.bias._loft.grad, store=store)
                     for m in self.fpn_convs]

    def forward(self, inputs):
        assert len(inputs) == self.num_ins
        outs = [inputs[0]]
        for i in range(1, self.num_ins):
            outs.append(F.interpolate(inputs[i], scale_factor=2 ** i, mode='bilinear',
                                      align_corners=False))
        cat = _RoundTF32.apply(torch.cat(outs, dim=1).contiguous(memory_format=torch.channels_last))
        c = self.reduction_conv.conv
        out = D.conv(cat, self._red, triggers=(c.weight, c.bias))
        pyr = [out]
        for i in range(1, self.num_outs):
            pyr.append(_RoundTF32.apply(self.pooling(out, kernel_size=2 ** i, stride=2 ** i)))
        res = []
        for i in range(self.num_outs):
            m = self.fpn_convs[i].conv
            res.append(D.conv(pyr[i], self._out[i], triggers=(m.weight, m.bias)))
        return tuple(res)
