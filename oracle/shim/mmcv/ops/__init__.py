"""mmcv.ops stand-in.  Only the ops on the LOFT path are real; they delegate to torchvision, which
implements the same Detectron-lineage algorithms (the reference itself flips RoI layers to
use_torchvision=True for CPU inference at mmdet/apis/inference.py:102-109).  SURVEY.md App. A."""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.ops as tvo
from torch.nn.modules.utils import _pair

from _shim_dummy import install_getattr as _ig
# import submodules eagerly, THEN rebind function names (a later `import mmcv.ops.nms` would
# otherwise rebind the attribute `nms` from the function to the module)
from . import roi_align as _roi_align_mod, nms as _nms_mod, carafe, merge_cells  # noqa: F401

Conv2d = nn.Conv2d
ConvTranspose2d = nn.ConvTranspose2d
MaxPool2d = nn.MaxPool2d
Linear = nn.Linear

RoIAlign = _roi_align_mod.RoIAlign
roi_align = _roi_align_mod.roi_align
nms = _nms_mod.nms
batched_nms = _nms_mod.batched_nms
soft_nms = _nms_mod.soft_nms


def sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25, weight=None, reduction='mean'):
    """mmcv CUDA op semantics: target[N] int64 in [0, C] (C = background); see SURVEY App. A."""
    C = pred.size(1)
    t = F.one_hot(target, C + 1)[:, :C].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, t, reduction='none') * fw
    if weight is not None:
        loss = loss * weight.view(-1, 1)
    if reduction == 'mean':
        return loss.sum() / pred.size(0)
    if reduction == 'sum':
        return loss.sum()
    return loss


def get_compiler_version():
    return 'shim'


def get_compiling_cuda_version():
    return 'shim'


_ig(globals(), 'mmcv.ops')
