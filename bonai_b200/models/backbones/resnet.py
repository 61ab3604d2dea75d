"""ResNet backbone (mmdet/models/backbones/resnet.py:95-664, models/utils/res_layer.py:5-102) with
the reference's parameter names / constructor, running on the fused tcgen05 kernels:
every conv is one implicit-GEMM launch whose epilogue applies the eval-mode BatchNorm affine,
the residual add and the ReLU.  BN runs in eval mode (norm_eval=True, resnet.py:640-649) -- the
only mode the loft_foa config uses (bonai_loft_foa_r50_fpn_basic.py:12)."""
import torch
import torch.nn as nn

from ..builder import BACKBONES
from ..init_utils import constant_init, kaiming_init
from ...ops import dense as D
from ...ops import misc as M


class ResLayer(nn.Sequential):
    def __init__(self, block, inplanes, planes, num_blocks, stride=1, norm_cfg=None, **kwargs):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(inplanes, planes * block.expansion, kernel_size=1, stride=stride,
                          bias=False),
                _make_bn(planes * block.expansion, norm_cfg))
        layers = [block(inplanes=inplanes, planes=planes, stride=stride, downsample=downsample,
                        norm_cfg=norm_cfg, **kwargs)]
        inplanes = planes * block.expansion
        for _ in range(1, num_blocks):
            layers.append(block(inplanes=inplanes, planes=planes, stride=1, norm_cfg=norm_cfg,
                                **kwargs))
        super().__init__(*layers)


def _make_bn(c, norm_cfg):
    cfg = dict(norm_cfg or dict(type='BN'))
    if cfg.pop('type') not in ('BN', 'BN2d'):
        raise NotImplementedError('only BN is supported on the LOFT path')
    requires_grad = cfg.pop('requires_grad', True)
    bn = nn.BatchNorm2d(c, **cfg)
    for p in bn.parameters():
        p.requires_grad = requires_grad
    return bn


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 with_cp=False, conv_cfg=None, norm_cfg=dict(type='BN'), dcn=None, plugins=None):
        super().__init__()
        assert style in ['pytorch', 'caffe']
        if dcn is not None or plugins is not None or conv_cfg is not None or dilation != 1:
            raise NotImplementedError('dcn / plugins / dilation are not on the LOFT path')
        self.inplanes, self.planes, self.stride, self.style = inplanes, planes, stride, style
        self.with_cp = with_cp
        if style == 'pytorch':
            self.conv1_stride, self.conv2_stride = 1, stride
        else:
            self.conv1_stride, self.conv2_stride = stride, 1
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=self.conv1_stride, bias=False)
        self.bn1 = _make_bn(planes, norm_cfg)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=self.conv2_stride, padding=1, bias=False)
        self.bn2 = _make_bn(planes, norm_cfg)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _make_bn(planes * 4, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    @property
    def norm3(self):
        return self.bn3

    def loft_prepare(self, store):
        def spec(conv, bn, relu, **kw):
            return D.ConvSpec(conv.weight._loft, ksize=conv.kernel_size[0], stride=conv.stride[0],
                              padding=conv.padding[0], relu=relu, bn=bn._loft_bn,
                              bn_trainable=bn.weight.requires_grad, store=store, **kw)
        # conv1 -> conv2 -> conv3 is a single-consumer chain: each ReLU's backward mask is applied
        # by the next conv's dgrad epilogue instead of a separate pass
        self._s1 = spec(self.conv1, self.bn1, True, grad_premasked=True)
        self._s2 = spec(self.conv2, self.bn2, True, premask_in=True, grad_premasked=True)
        self._s3 = spec(self.conv3, self.bn3, True, premask_in=True)   # ReLU after the residual add
        self._sd = spec(self.downsample[0], self.downsample[1], False) \
            if self.downsample is not None else None
        self._trainable = self.conv1.weight.requires_grad
        store.fold_bn(self.conv1.weight, self.bn1)
        store.fold_bn(self.conv2.weight, self.bn2)
        store.fold_bn(self.conv3.weight, self.bn3)
        if self.downsample is not None:
            store.fold_bn(self.downsample[0].weight, self.downsample[1])
        if self._trainable:
            # beta gradients of bn1 / bn2 come out of conv2's / conv3's dgrad epilogues
            D.link_chain([self._s1, self._s2, self._s3])

    def forward(self, x):
        if not self._trainable or not torch.is_grad_enabled():
            out = D.conv_nograd(x, self._s1)
            out = D.conv_nograd(out, self._s2)
            identity = x if self._sd is None else D.conv_nograd(x, self._sd)
            return D.conv_nograd(out, self._s3, residual=identity)
        out = D.conv(x, self._s1, triggers=(self.conv1.weight,))
        out = D.conv(out, self._s2, triggers=(self.conv2.weight,))
        identity = x if self._sd is None else \
            D.conv(x, self._sd, triggers=(self.downsample[0].weight,))
        return D.conv(out, self._s3, residual=identity, triggers=(self.conv3.weight,))


class BasicBlock(nn.Module):
    """ResNet-18/34 are not on the LOFT path (the HRNet stages use backbones/hrnet.BasicBlock)."""
    expansion = 1

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError('BasicBlock ResNets (depth 18/34) are not on the LOFT path')


@BACKBONES.register_module()
class ResNet(nn.Module):
    arch_settings = {
        18: (BasicBlock, (2, 2, 2, 2)), 34: (BasicBlock, (3, 4, 6, 3)),
        50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)),
        152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, in_channels=3, stem_channels=64, base_channels=64, num_stages=4,
                 strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3),
                 style='pytorch', deep_stem=False, avg_down=False, frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False,
                 zero_init_residual=True):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f'invalid depth {depth} for resnet')
        if deep_stem or avg_down or dcn is not None or plugins is not None or conv_cfg is not None:
            raise NotImplementedError('deep_stem / avg_down / dcn / plugins are not on the LOFT path')
        if not norm_eval:
            raise NotImplementedError('LOFT path: BatchNorm runs in eval mode (norm_eval=True, '
                                      'bonai_loft_foa_r50_fpn_basic.py:12)')
        if frozen_stages < 0:
            raise NotImplementedError('LOFT path: the 7x7 stem is forward-only (frozen_stages>=0)')
        assert 1 <= num_stages <= 4 and len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        self.depth, self.num_stages = depth, num_stages
        self.strides, self.dilations, self.out_indices = strides, dilations, out_indices
        self.style, self.frozen_stages = style, frozen_stages
        self.norm_cfg, self.norm_eval, self.with_cp = norm_cfg, norm_eval, with_cp
        self.zero_init_residual = zero_init_residual
        self.block, stage_blocks = self.arch_settings[depth]
        self.stage_blocks = stage_blocks[:num_stages]
        self.inplanes = stem_channels
        self.conv1 = nn.Conv2d(in_channels, stem_channels, 7, stride=2, padding=3, bias=False)
        self.bn1 = _make_bn(stem_channels, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.res_layers = []
        for i, num_blocks in enumerate(self.stage_blocks):
            planes = base_channels * 2 ** i
            layer = ResLayer(self.block, self.inplanes, planes, num_blocks, stride=strides[i],
                             norm_cfg=norm_cfg, style=style, with_cp=with_cp)
            self.inplanes = planes * self.block.expansion
            name = f'layer{i + 1}'
            self.add_module(name, layer)
            self.res_layers.append(name)
        self._freeze_stages()
        self.feat_dim = self.block.expansion * base_channels * 2 ** (len(self.stage_blocks) - 1)

    @property
    def norm1(self):
        return self.bn1

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.bn1.eval()
            for m in [self.conv1, self.bn1]:
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            raise NotImplementedError('checkpoint download is unavailable here; use '
                                      'pretrained=None and load_state_dict')
        if pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                kaiming_init(m)
            elif isinstance(m, nn.modules.batchnorm._BatchNorm):
                constant_init(m, 1)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    constant_init(m.bn3, 0)

    def loft_prepare(self, store):
        from ... import _lib as L
        import ctypes
        w = self.conv1.weight                          # [64,3,7,7], physically [64,7,7,3]
        store.fold_bn(w, self.bn1)
        assert w.shape[1] <= 4, 'the direct stem conv packs the image as NHWC4'
        self._stem_w = torch.zeros((w.shape[0], M.STEM_K), device=store.device)

        def build():
            # the TF32 copy is logically [Cout,C,7,7] over [Cout,7,7,C] memory (engine.ParamStore)
            M.pack_stem_weight(w._loft.w.permute(0, 2, 3, 1), self._stem_w)

        from ...engine import Packed
        store.add_packed(Packed(self._stem_w, None, None, None, build, lambda: None))

    def forward(self, x):
        bn = self.bn1._loft_bn
        with torch.no_grad():
            x = M.stem_conv(x, self._stem_w, None, bn.shift)
            x = M.maxpool3x3s2(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
