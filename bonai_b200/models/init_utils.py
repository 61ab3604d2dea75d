"""Weight-init helpers with mmcv.cnn semantics (kaiming / xavier / normal / constant), used by the
reference's init_weights (resnet.py:591-621, fpn.py:158-162, rpn_head.py:32-36, ...)."""
import torch.nn as nn


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if distribution == 'uniform':
        nn.init.xavier_uniform_(module.weight, gain=gain)
    else:
        nn.init.xavier_normal_(module.weight, gain=gain)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


class ConvModule(nn.Module):
    """Parameter container with mmcv.cnn.ConvModule's attribute layout (`.conv`, so state_dict
    keys read `...lateral_convs.0.conv.weight`).  conv -> (no norm) -> optional
    ReLU; the math runs in the fused tcgen05 kernels, not here."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True):
        super().__init__()
        if norm_cfg is not None or conv_cfg is not None:
            raise NotImplementedError('LOFT path: ConvModule without norm / custom conv only')
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, bias=bias)
        self.with_activation = act_cfg is not None
        self.with_norm = False
        kaiming_init(self.conv, nonlinearity='relu')
