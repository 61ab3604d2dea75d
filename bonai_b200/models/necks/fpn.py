"""FPN neck (mmdet/models/necks/fpn.py:9-216).  The top-down pathway is fused into the lateral
1x1 convs: lateral i-1's epilogue adds the nearest-upsampled (already merged) lateral i, so the
`laterals[i-1] += F.interpolate(laterals[i])` pass (fpn.py:181-186) never touches HBM twice."""
import torch.nn as nn

from ..builder import NECKS
from ..init_utils import ConvModule, xavier_init
from ...ops import dense as D
from ...ops import misc as M


@NECKS.register_module()
class FPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1,
                 add_extra_convs=False, extra_convs_on_inputs=True, relu_before_extra_convs=False,
                 no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest')):
        super().__init__()
        assert isinstance(in_channels, list)
        if add_extra_convs or norm_cfg is not None or conv_cfg is not None or act_cfg is not None:
            raise NotImplementedError('LOFT path: plain FPN (no extra convs / norm / act)')
        if dict(upsample_cfg) != dict(mode='nearest'):
            raise NotImplementedError('LOFT path: nearest top-down upsampling only')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs = len(in_channels), num_outs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.no_norm_on_lateral = no_norm_on_lateral
        self.fp16_enabled = False
        self.upsample_cfg = dict(upsample_cfg)
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels)
            assert num_outs == end_level - start_level
        self.start_level, self.end_level = start_level, end_level
        self.add_extra_convs = add_extra_convs
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, act_cfg=None))
            self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, padding=1,
                                             act_cfg=None))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def loft_prepare(self, store):
        n = len(self.lateral_convs)
        self._lat, self._out = [], []
        for i in range(n):
            c = self.lateral_convs[i].conv
            self._lat.append(D.ConvSpec(c.weight._loft, ksize=1, bias=c.bias, store=store,
                                        bias_grad=c.bias._loft.grad, res_upsample=(i < n - 1)))
            c = self.fpn_convs[i].conv
            self._out.append(D.ConvSpec(c.weight._loft, ksize=3, padding=1, bias=c.bias,
                                        bias_grad=c.bias._loft.grad, store=store))

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        n = len(self.lateral_convs)
        lat = [None] * n
        for i in range(n - 1, -1, -1):
            x = inputs[i + self.start_level]
            c = self.lateral_convs[i].conv
            res = None
            if i < n - 1:
                res = lat[i + 1]
                assert x.shape[2] == 2 * res.shape[2] and x.shape[3] == 2 * res.shape[3], \
                    'FPN levels must differ by exactly 2x (pad inputs to a multiple of 32)'
            lat[i] = D.conv(x, self._lat[i], residual=res, triggers=(c.weight, c.bias))
        outs = []
        for i in range(n):
            c = self.fpn_convs[i].conv
            outs.append(D.conv(lat[i], self._out[i], triggers=(c.weight, c.bias)))
        while len(outs) < self.num_outs:
            outs.append(M.subsample2(outs[-1]))        # F.max_pool2d(x, 1, stride=2), fpn.py:199
        return tuple(outs)
