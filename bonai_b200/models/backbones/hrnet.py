"""HRNet backbone (mmdet/models/backbones/hrnet.py:12-537) with the reference's constructor and
parameter names, on the fused tcgen05 kernels: every conv (+ eval-mode BN folded into its weights,
+ residual, + ReLU) is one implicit-GEMM launch.  BASELINE.json configs[3] (LOFT+FOA over
HRNetV2p-W32, configs/hrnet/mask_rcnn_hrnetv2p_w32_1x_coco.py:1-36) stresses the dense core with
many small-channel (32 / 64 / 128 / 256) 3x3 convs at four resolutions and the cross-resolution
exchange of every HRModule (hrnet.py:115-195).

Differences from the reference implementation, none visible in results: BN runs folded (norm_eval
is the reference's default, hrnet.py:262, and the only mode supported here); the exchange sums are
accumulated in the conv epilogues where the incoming term is a conv output at the target
resolution (the strided 3x3 down paths), and by a nearest-upsample-add otherwise."""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from ..builder import BACKBONES
from ..init_utils import constant_init, kaiming_init
from ... import _lib as L
from ...engine import Packed, WeightRef
from ...ops import dense as D
from .resnet import Bottleneck, _make_bn

i32 = ctypes.c_int


def _spec(conv, bn, relu, store, **kw):
    return D.ConvSpec(conv.weight._loft, ksize=conv.kernel_size[0], stride=conv.stride[0],
                      padding=conv.padding[0], relu=relu, bn=bn._loft_bn,
                      bn_trainable=bn.weight.requires_grad, store=store, **kw)


class _ConvBN(nn.Sequential):
    """conv (bias=False) -> BN [-> ReLU] as laid out by the reference (nn.Sequential of conv, norm
    and optionally nn.ReLU, hrnet.py:127-160,349-386): one fused launch."""

    def __init__(self, cin, cout, k, stride, norm_cfg, relu):
        mods = [nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=False),
                _make_bn(cout, norm_cfg)]
        if relu:
            mods.append(nn.ReLU(inplace=False))
        super().__init__(*mods)
        self._relu = relu

    def loft_prepare(self, store):
        self._s = _spec(self[0], self[1], self._relu, store)
        store.fold_bn(self[0].weight, self[1])

    def forward(self, x, residual=None):
        return D.conv(x, self._s, residual=residual, triggers=(self[0].weight,))


class BasicBlock(nn.Module):
    """resnet.py:12-92: conv3x3-BN-ReLU-conv3x3-BN, + identity, ReLU."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 with_cp=False, conv_cfg=None, norm_cfg=dict(type='BN'), dcn=None, plugins=None):
        super().__init__()
        if dcn is not None or plugins is not None or conv_cfg is not None or dilation != 1:
            raise NotImplementedError('dcn / plugins / dilation are not on the LOFT path')
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _make_bn(planes, norm_cfg)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _make_bn(planes, norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def loft_prepare(self, store):
        self._s1 = _spec(self.conv1, self.bn1, True, store, grad_premasked=True)
        self._s2 = _spec(self.conv2, self.bn2, True, store, premask_in=True)   # ReLU after the add
        self._sd = _spec(self.downsample[0], self.downsample[1], False, store) \
            if self.downsample is not None else None
        store.fold_bn(self.conv1.weight, self.bn1)
        store.fold_bn(self.conv2.weight, self.bn2)
        if self.downsample is not None:
            store.fold_bn(self.downsample[0].weight, self.downsample[1])
        if self.conv1.weight.requires_grad and self.stride == 1:
            D.link_chain([self._s1, self._s2])

    def forward(self, x):
        out = D.conv(x, self._s1, triggers=(self.conv1.weight,))
        identity = x if self._sd is None else \
            D.conv(x, self._sd, triggers=(self.downsample[0].weight,))
        return D.conv(out, self._s2, residual=identity, triggers=(self.conv2.weight,))


class _RoundTF32(Function):
    """Identity whose forward rounds to the TF32 grid: tensors produced by ATen ops (the exchange
    sums) are read next by tensor-core convs, which TRUNCATE unrounded operands."""

    @staticmethod
    def forward(ctx, x):
        xn = D.nhwc(x)
        y = D.new_nhwc(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
        n = xn.numel()
        L.call('copy2d', L.ptr(xn), L.ll(n), L.ptr(y.permute(0, 2, 3, 1)), L.ll(n), L.ll(1),
               i32(n), i32(0), i32(1), L.stream())
        return y

    @staticmethod
    def backward(ctx, g):
        return g


class HRModule(nn.Module):
    """hrnet.py:12-195: `num_branches` parallel stacks of blocks, then the exchange: output i =
    ReLU(sum_j f_ij(x_j)) with f_ii = identity, f_ij (j > i) = 1x1 conv + BN + nearest upsample by
    2^(j-i), f_ij (j < i) = (i-j) strided 3x3 conv + BN (+ ReLU between them)."""

    def __init__(self, num_branches, blocks, num_blocks, in_channels, num_channels,
                 multiscale_output=True, with_cp=False, conv_cfg=None, norm_cfg=dict(type='BN')):
        super().__init__()
        if num_branches != len(num_blocks):
            raise ValueError(f'NUM_BRANCHES({num_branches}) != NUM_BLOCKS({len(num_blocks)})')
        if num_branches != len(num_channels):
            raise ValueError(f'NUM_BRANCHES({num_branches}) != NUM_CHANNELS({len(num_channels)})')
        if num_branches != len(in_channels):
            raise ValueError(f'NUM_BRANCHES({num_branches}) != NUM_INCHANNELS({len(in_channels)})')
        self.in_channels = in_channels
        self.num_branches = num_branches
        self.multiscale_output = multiscale_output
        self.norm_cfg = norm_cfg
        self.branches = nn.ModuleList(
            self._make_one_branch(i, blocks, num_blocks, num_channels) for i in range(num_branches))
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(inplace=False)

    def _make_one_branch(self, idx, block, num_blocks, num_channels, stride=1):
        downsample = None
        cout = num_channels[idx] * block.expansion
        if stride != 1 or self.in_channels[idx] != cout:
            downsample = nn.Sequential(
                nn.Conv2d(self.in_channels[idx], cout, 1, stride=stride, bias=False),
                _make_bn(cout, self.norm_cfg))
        layers = [block(self.in_channels[idx], num_channels[idx], stride, downsample=downsample,
                        norm_cfg=self.norm_cfg)]
        self.in_channels[idx] = cout
        for _ in range(1, num_blocks[idx]):
            layers.append(block(self.in_channels[idx], num_channels[idx], norm_cfg=self.norm_cfg))
        return nn.Sequential(*layers)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        nb, c = self.num_branches, self.in_channels
        fuse = []
        for i in range(nb if self.multiscale_output else 1):
            row = []
            for j in range(nb):
                if j > i:
                    m = _ConvBN(c[j], c[i], 1, 1, self.norm_cfg, relu=False)
                    m.add_module('2', nn.Upsample(scale_factor=2 ** (j - i), mode='nearest'))
                    row.append(m)
                elif j == i:
                    row.append(None)
                else:
                    downs = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        downs.append(_ConvBN(c[j], c[i] if last else c[j], 3, 2, self.norm_cfg,
                                             relu=not last))
                    row.append(nn.Sequential(*downs))
            fuse.append(nn.ModuleList(row))
        return nn.ModuleList(fuse)

    def forward(self, x):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        x = [self.branches[i](x[i]) for i in range(self.num_branches)]
        out = []
        for i in range(len(self.fuse_layers)):
            # terms from finer branches end in a conv at this resolution: each one adds the
            # running sum in its epilogue; coarser branches are upsampled and added afterwards
            y = x[i]
            for j in range(i):
                chain = self.fuse_layers[i][j]
                t = x[j]
                for k, m in enumerate(chain):
                    t = m(t, residual=y if k == len(chain) - 1 else None)
                y = t
            for j in range(i + 1, self.num_branches):
                m = self.fuse_layers[i][j]
                t = D.conv(x[j], m._s, triggers=(m[0].weight,))
                y = y + F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
            out.append(_RoundTF32.apply(torch.relu(y)))
        return out


class _StemConvFn(Function):
    """3x3 / stride 2 conv over the NCHW input image (3 -> 64) + folded BN + ReLU, trainable
    (hrnet.py:281-289): im2col straight from the image into a K-padded matrix, one GEMM; the
    backward is one masked reduction (beta) and one weight-gradient GEMM -- the image needs no
    gradient."""

    @staticmethod
    def forward(ctx, img, stem, *triggers):
        N, C, H, W = img.shape
        Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        img = img.contiguous().float()
        kp = stem._kpad
        col = torch.empty((N * Ho * Wo, kp), device=img.device, dtype=torch.float32)
        st = L.stream()
        L.call('im2col', L.ptr(img), L.ptr(col), i32(N), i32(H), i32(W), i32(C), i32(3), i32(3),
               i32(2), i32(1), i32(kp), i32(1), st)
        cout = stem._w.shape[0]
        y = D.new_nhwc(N, cout, Ho, Wo, img.device)
        e = L.make_epilogue(shift=stem._bn.shift, relu=True, round_out=True)
        L.call('gemm_fprop', L.ptr(col), L.ptr(stem._w), L.ptr(y.permute(0, 2, 3, 1)),
               L.ll(N * Ho * Wo), i32(kp), i32(cout), L.ll(kp), L.ll(kp), L.ll(cout), i32(Ho),
               i32(Wo), ctypes.byref(e), st)
        ctx.stem = stem
        ctx.save_for_backward(col, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        stem = ctx.stem
        col, y = ctx.saved_tensors
        stem._store.queue_finalize()
        st = L.stream()
        dyn, yn = D.nhwc(dy), D.nhwc(y)
        P, cout = col.shape[0], stem._w.shape[0]
        dz = torch.empty_like(dyn)
        dbeta = stem._bn.dbeta if stem._trainable_bn else None
        L.call('act_bwd', L.ptr(dyn), L.ptr(yn), None, None, None, None, L.ptr(dz), None, None,
               L.ptr(dbeta) if dbeta is not None else None, L.ll(P), i32(cout), i32(1), st)
        L.call('gemm_wgrad', L.ptr(dz), L.ptr(col), L.ptr(stem._gw), L.ll(P), i32(stem._kpad),
               i32(cout), L.ll(cout), L.ll(stem._kpad), L.ll(stem._kpad), st)
        return (None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


@BACKBONES.register_module()
class HRNet(nn.Module):
    blocks_dict = {'BASIC': BasicBlock, 'BOTTLENECK': Bottleneck}

    def __init__(self, extra, in_channels=3, conv_cfg=None, norm_cfg=dict(type='BN'),
                 norm_eval=True, with_cp=False, zero_init_residual=False):
        super().__init__()
        if conv_cfg is not None:
            raise NotImplementedError('conv_cfg is not on the LOFT path')
        if not norm_eval:
            raise NotImplementedError('LOFT path: BatchNorm runs in eval mode (norm_eval=True, the '
                                      'default of hrnet.py:262)')
        self.extra, self.norm_cfg, self.norm_eval = extra, norm_cfg, norm_eval
        self.with_cp, self.zero_init_residual = with_cp, zero_init_residual
        self.conv1 = nn.Conv2d(in_channels, 64, 3, stride=2, padding=1, bias=False)
        self.bn1 = _make_bn(64, norm_cfg)
        self.conv2 = nn.Conv2d(64, 64, 3, stride=2, padding=1, bias=False)
        self.bn2 = _make_bn(64, norm_cfg)
        self.relu = nn.ReLU(inplace=True)

        cfg1 = self.stage1_cfg = extra['stage1']
        block = self.blocks_dict[cfg1['block']]
        c1 = cfg1['num_channels'][0]
        self.layer1 = self._make_layer(block, 64, c1, cfg1['num_blocks'][0])
        pre = [c1 * block.expansion]
        for s in (2, 3, 4):
            cfg = extra[f'stage{s}']
            setattr(self, f'stage{s}_cfg', cfg)
            block = self.blocks_dict[cfg['block']]
            chans = [c * block.expansion for c in cfg['num_channels']]
            setattr(self, f'transition{s - 1}', self._make_transition_layer(pre, chans))
            stage, pre = self._make_stage(cfg, chans)
            setattr(self, f'stage{s}', stage)

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def _make_transition_layer(self, pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                layers.append(_ConvBN(pre[i], cur[i], 3, 1, self.norm_cfg, relu=True)
                              if cur[i] != pre[i] else None)
            else:
                downs = []
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    downs.append(_ConvBN(cin, cout, 3, 2, self.norm_cfg, relu=True))
                layers.append(nn.Sequential(*downs))
        return nn.ModuleList(layers)

    def _make_layer(self, block, inplanes, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                _make_bn(planes * block.expansion, self.norm_cfg))
        layers = [block(inplanes, planes, stride, downsample=downsample, norm_cfg=self.norm_cfg)]
        inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(inplanes, planes, norm_cfg=self.norm_cfg))
        return nn.Sequential(*layers)

    def _make_stage(self, cfg, in_channels, multiscale_output=True):
        block = self.blocks_dict[cfg['block']]
        mods = []
        for i in range(cfg['num_modules']):
            reset = not (not multiscale_output and i == cfg['num_modules'] - 1)
            mods.append(HRModule(cfg['num_branches'], block, cfg['num_blocks'], in_channels,
                                 cfg['num_channels'], reset, norm_cfg=self.norm_cfg))
        return nn.Sequential(*mods), in_channels

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
                elif isinstance(m, BasicBlock):
                    constant_init(m.bn2, 0)

    # ------------------------------------------------------------------ kernel-side weights
    def loft_prepare(self, store):
        self._store = store
        w = self.conv1.weight                      # [64,3,3,3], physically [64,3,3,3(cin)]
        store.fold_bn(w, self.bn1)
        K = w.shape[1] * 9
        self._kpad = 32                            # K = 27 padded to one 32-float k-block
        cout = w.shape[0]
        self._w = torch.zeros((cout, self._kpad), device=store.device)
        self._gw = torch.zeros((cout, self._kpad), device=store.device)
        self._bn = self.bn1._loft_bn
        self._trainable_bn = self.bn1.weight.requires_grad

        def cp(src, lds, dst, ldd, acc, rnd):
            L.call('copy2d', L.ptr(src), L.ll(lds), L.ptr(dst), L.ll(ldd), L.ll(cout), i32(K),
                   i32(acc), i32(rnd), L.stream())

        def build():
            cp(w._loft.w, K, self._w, self._kpad, 0, 1)

        def scatter():
            if w._loft.grad is not None:
                cp(self._gw, self._kpad, w._loft.grad, K, 1, 0)
        store.add_packed(Packed(self._w, None, self._gw, None, build, scatter))
        self._s2 = _spec(self.conv2, self.bn2, True, store)
        store.fold_bn(self.conv2.weight, self.bn2)

    def forward(self, x):
        x = _StemConvFn.apply(x, self, self.conv1.weight)
        x = D.conv(x, self._s2, triggers=(self.conv2.weight,))
        x = self.layer1(x)
        y_list = [x]
        for s in (2, 3, 4):
            trans = getattr(self, f'transition{s - 1}')
            x_list = []
            for i in range(getattr(self, f'stage{s}_cfg')['num_branches']):
                t = trans[i]
                if t is None:
                    x_list.append(y_list[i])
                    continue
                v = y_list[-1]          # hrnet.py:504-527 (stage 2: the single layer1 output)
                for m in ([t] if isinstance(t, _ConvBN) else t):
                    v = m(v)
                x_list.append(v)
            y_list = x_list
            for mod in getattr(self, f'stage{s}'):
                y_list = mod(y_list)
        return y_list

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
