"""SingleRoIExtractor (mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:9-84,
single_level_roi_extractor.py:9-80).  The reference runs one RoIAlign + boolean gather/scatter
(+ an `inds.any()` host sync) per FPN level; here the level mapping is computed inside a single
multi-level kernel launch."""
import torch
import torch.nn as nn

from ..builder_alias import ROI_EXTRACTORS
from .... import ops


class BaseRoIExtractor(nn.Module):
    def __init__(self, roi_layer, out_channels, featmap_strides):
        super().__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = dict(layer_cfg)
        layer_type = cfg.pop('type')
        assert hasattr(ops, layer_type), f'{layer_type} is not an op of bonai_b200.ops'
        layer_cls = getattr(ops, layer_type)          # looked up by name like mmcv.ops
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def roi_rescale(self, rois, scale_factor):
        cx = (rois[:, 1] + rois[:, 3]) * 0.5
        cy = (rois[:, 2] + rois[:, 4]) * 0.5
        w = (rois[:, 3] - rois[:, 1]) * scale_factor
        h = (rois[:, 4] - rois[:, 2]) * scale_factor
        return torch.stack((rois[:, 0], cx - w * 0.5, cy - h * 0.5, cx + w * 0.5, cy + h * 0.5),
                           dim=-1)


@ROI_EXTRACTORS.register_module()
class SingleRoIExtractor(BaseRoIExtractor):
    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super().__init__(roi_layer, out_channels, featmap_strides)
        self.finest_scale = finest_scale

    def map_roi_levels(self, rois, num_levels):
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def forward(self, feats, rois, roi_scale_factor=None):
        out_size = self.roi_layers[0].output_size
        assert out_size[0] == out_size[1]
        if roi_scale_factor is not None:
            rois = self.roi_rescale(rois, roi_scale_factor)
        return ops.multilevel_roi_align(list(feats[:self.num_inputs]), rois, out_size[0],
                                        self.featmap_strides, self.finest_scale)
