"""mmcv.cnn stand-in: ConvModule, layer builders and weight-init helpers (mmcv v1.0.5 semantics)."""
import numpy as np
import torch.nn as nn

from _shim_dummy import install_getattr as _ig


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def uniform_init(module, a=0, b=1, bias=0):
    nn.init.uniform_(module.weight, a, b)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def caffe2_xavier_init(module, bias=0):
    kaiming_init(module, a=1, mode='fan_in', nonlinearity='leaky_relu', distribution='uniform')


def bias_init_with_prob(prior_prob):
    return float(-np.log((1 - prior_prob) / prior_prob))


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is None:
        cfg_ = dict(type='Conv2d')
    else:
        cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type not in ('Conv', 'Conv2d'):
        raise KeyError(f'oracle shim: conv layer {layer_type} not supported')
    return nn.Conv2d(*args, **kwargs, **cfg_)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type not in ('BN', 'BN2d'):
        raise KeyError(f'oracle shim: norm layer {layer_type} not supported')
    name = 'bn' + str(postfix)
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    layer = nn.BatchNorm2d(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name, layer


def build_upsample_layer(cfg, *args, **kwargs):
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type == 'deconv':
        return nn.ConvTranspose2d(*args, **kwargs, **cfg_)
    if layer_type in ('nearest', 'bilinear'):
        return nn.Upsample(*args, mode=layer_type, **kwargs, **cfg_)
    raise KeyError(f'oracle shim: upsample layer {layer_type} not supported')


def build_activation_layer(cfg):
    cfg_ = dict(cfg)
    t = cfg_.pop('type')
    return getattr(nn, t)(**cfg_)


class ConvModule(nn.Module):
    """conv -> norm -> act block; bias='auto' means bias iff no norm."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'),
                 inplace=True, with_spectral_norm=False, padding_mode='zeros',
                 order=('conv', 'norm', 'act')):
        super().__init__()
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.inplace = inplace
        self.order = order
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                     stride=stride, padding=padding, dilation=dilation,
                                     groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.dilation, self.groups = dilation, groups
        if self.with_norm:
            norm_channels = out_channels if order.index('norm') > order.index('conv') \
                else in_channels
            self.norm_name, norm = build_norm_layer(norm_cfg, norm_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_['type'] not in ['Tanh', 'PReLU', 'Sigmoid']:
                act_cfg_.setdefault('inplace', inplace)
            self.activate = build_activation_layer(act_cfg_)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name)

    def init_weights(self):
        if self.with_activation and self.act_cfg['type'] == 'LeakyReLU':
            nonlinearity, a = 'leaky_relu', self.act_cfg.get('negative_slope', 0.01)
        else:
            nonlinearity, a = 'relu', 0
        kaiming_init(self.conv, a=a, nonlinearity=nonlinearity)
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x, activate=True, norm=True):
        for layer in self.order:
            if layer == 'conv':
                x = self.conv(x)
            elif layer == 'norm' and norm and self.with_norm:
                x = self.norm(x)
            elif layer == 'act' and activate and self.with_activation:
                x = self.activate(x)
        return x


from . import bricks  # noqa: E402,F401

_ig(globals(), 'mmcv.cnn')
