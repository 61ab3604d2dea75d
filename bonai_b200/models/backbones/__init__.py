from .resnet import ResNet, Bottleneck, ResLayer
from .hrnet import HRNet, HRModule, BasicBlock

__all__ = ['ResNet', 'Bottleneck', 'BasicBlock', 'ResLayer', 'HRNet', 'HRModule']
