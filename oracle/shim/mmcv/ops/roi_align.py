import torch.nn as nn
import torchvision.ops as tvo
from torch.nn.modules.utils import _pair


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
              aligned=True):
    assert pool_mode == 'avg'
    return tvo.roi_align(input, rois, _pair(output_size), spatial_scale, sampling_ratio, aligned)


class RoIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio,
                         self.pool_mode, self.aligned)
