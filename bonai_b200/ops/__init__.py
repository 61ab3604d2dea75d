"""Op surface mirroring what the reference imports from mmcv.ops (mmdet/ops/__init__.py:5-16):
RoIAlign / roi_align, nms / batched_nms, sigmoid_focal_loss, Conv2d ... backed by the sm_100a
kernels of libloft_b200.so.  RoI layers are looked up here by name, exactly like
`getattr(mmcv.ops, cfg['type'])` in roi_extractors/base_roi_extractor.py:49-55."""
import torch.nn as nn

from .roi import RoIAlign, roi_align, multilevel_roi_align, mask_target_sample
from .nms import nms, batched_nms, nms_sorted, nms_segmented, soft_nms
from .focal import sigmoid_focal_loss
from .infer import paste_masks, offset_fusion_decode

Conv2d = nn.Conv2d            # parameter containers; their math runs through ops.dense
ConvTranspose2d = nn.ConvTranspose2d
Linear = nn.Linear
MaxPool2d = nn.MaxPool2d

__all__ = ['RoIAlign', 'roi_align', 'multilevel_roi_align', 'mask_target_sample', 'nms',
           'batched_nms', 'nms_sorted', 'soft_nms', 'sigmoid_focal_loss', 'paste_masks',
           'offset_fusion_decode', 'Conv2d', 'ConvTranspose2d',
           'Linear', 'MaxPool2d']
