import torch.distributed as dist

from _shim_dummy import install_getattr as _ig
from ..utils import Registry

HOOKS = Registry('hook')


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def load_checkpoint(*a, **k):
    raise NotImplementedError('oracle shim: load_checkpoint (use pretrained=None)')


class Hook:
    pass


class OptimizerHook(Hook):
    def __init__(self, grad_clip=None):
        self.grad_clip = grad_clip


_ig(globals(), 'mmcv.runner')
