from .resnet import ResNet, Bottleneck, BasicBlock, ResLayer

__all__ = ['ResNet', 'Bottleneck', 'BasicBlock', 'ResLayer']
