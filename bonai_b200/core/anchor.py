"""AnchorGenerator and helpers (mmdet/core/anchor/anchor_generator.py:8-329, utils.py:4-46,
builder.py:1-7).  Anchors are a pure function of the feature-map sizes, so the grids are generated
once per size on the device and cached (the reference regenerates them every step)."""
import torch
from torch.nn.modules.utils import _pair

from ..registry import Registry, build_from_cfg

ANCHOR_GENERATORS = Registry('Anchor generator')


def build_anchor_generator(cfg, default_args=None):
    return build_from_cfg(cfg, ANCHOR_GENERATORS, default_args)


@ANCHOR_GENERATORS.register_module()
class AnchorGenerator:
    def __init__(self, strides, ratios, scales=None, base_sizes=None, scale_major=True,
                 octave_base_scale=None, scales_per_octave=None, centers=None, center_offset=0.):
        if center_offset != 0:
            assert centers is None, 'center cannot be set when center_offset != 0'
        if not (0 <= center_offset <= 1):
            raise ValueError(f'center_offset should be in range [0, 1], {center_offset} is given.')
        if centers is not None:
            assert len(centers) == len(strides)
        self.strides = [_pair(s) for s in strides]
        self.base_sizes = [min(s) for s in self.strides] if base_sizes is None else base_sizes
        assert len(self.base_sizes) == len(self.strides)
        assert ((octave_base_scale is not None and scales_per_octave is not None) ^
                (scales is not None)), \
            'scales and octave_base_scale with scales_per_octave cannot be set at the same time'
        if scales is not None:
            self.scales = torch.Tensor(scales)
        else:
            import numpy as np
            octave_scales = np.array([2 ** (i / scales_per_octave)
                                      for i in range(scales_per_octave)])
            self.scales = torch.Tensor(octave_scales * octave_base_scale)
        self.octave_base_scale = octave_base_scale
        self.scales_per_octave = scales_per_octave
        self.ratios = torch.Tensor(ratios)
        self.scale_major = scale_major
        self.centers = centers
        self.center_offset = center_offset
        self.base_anchors = self.gen_base_anchors()
        self._cache = {}

    @property
    def num_base_anchors(self):
        return [b.size(0) for b in self.base_anchors]

    @property
    def num_levels(self):
        return len(self.strides)

    def gen_base_anchors(self):
        out = []
        for i, base_size in enumerate(self.base_sizes):
            center = self.centers[i] if self.centers is not None else None
            out.append(self.gen_single_level_base_anchors(base_size, self.scales, self.ratios,
                                                          center))
        return out

    def gen_single_level_base_anchors(self, base_size, scales, ratios, center=None):
        w = h = base_size
        if center is None:
            x_center, y_center = self.center_offset * w, self.center_offset * h
        else:
            x_center, y_center = center
        h_ratios = torch.sqrt(ratios)
        w_ratios = 1 / h_ratios
        if self.scale_major:
            ws = (w * w_ratios[:, None] * scales[None, :]).view(-1)
            hs = (h * h_ratios[:, None] * scales[None, :]).view(-1)
        else:
            ws = (w * scales[:, None] * w_ratios[None, :]).view(-1)
            hs = (h * scales[:, None] * h_ratios[None, :]).view(-1)
        return torch.stack([x_center - 0.5 * ws, y_center - 0.5 * hs, x_center + 0.5 * ws,
                            y_center + 0.5 * hs], dim=-1)

    def single_level_grid_anchors(self, base_anchors, featmap_size, stride=(16, 16), device='cuda'):
        feat_h, feat_w = featmap_size
        base_anchors = base_anchors.to(device)
        shift_x = torch.arange(0, feat_w, device=device).to(base_anchors) * stride[0]
        shift_y = torch.arange(0, feat_h, device=device).to(base_anchors) * stride[1]
        xx = shift_x.repeat(len(shift_y))
        yy = shift_y.view(-1, 1).repeat(1, len(shift_x)).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1)
        return (base_anchors[None, :, :] + shifts[:, None, :]).view(-1, 4)

    def grid_anchors(self, featmap_sizes, device='cuda'):
        assert self.num_levels == len(featmap_sizes)
        key = (tuple(tuple(int(v) for v in s) for s in featmap_sizes), str(device))
        if key not in self._cache:
            self._cache[key] = [
                self.single_level_grid_anchors(self.base_anchors[i], featmap_sizes[i],
                                               self.strides[i], device=device)
                for i in range(self.num_levels)]
        return self._cache[key]

    def valid_flags(self, featmap_sizes, pad_shape, device='cuda'):
        assert self.num_levels == len(featmap_sizes)
        out = []
        for i in range(self.num_levels):
            stride = self.strides[i]
            feat_h, feat_w = featmap_sizes[i]
            h, w = pad_shape[:2]
            valid_h = min(-(-h // stride[1]), feat_h)
            valid_w = min(-(-w // stride[0]), feat_w)
            vx = torch.zeros(feat_w, dtype=torch.bool, device=device)
            vy = torch.zeros(feat_h, dtype=torch.bool, device=device)
            vx[:valid_w] = 1
            vy[:valid_h] = 1
            xx = vx.repeat(len(vy))
            yy = vy.view(-1, 1).repeat(1, len(vx)).view(-1)
            valid = xx & yy
            out.append(valid[:, None].expand(valid.size(0),
                                             self.num_base_anchors[i]).contiguous().view(-1))
        return out

    def __repr__(self):
        return (f'{self.__class__.__name__}(strides={self.strides}, ratios={self.ratios}, '
                f'scales={self.scales}, base_sizes={self.base_sizes}, '
                f'scale_major={self.scale_major}, center_offset={self.center_offset})')


def images_to_levels(target, num_levels):
    target = torch.stack(target, 0)
    out, start = [], 0
    for n in num_levels:
        out.append(target[:, start:start + n])
        start += n
    return out


def anchor_inside_flags(flat_anchors, valid_flags, img_shape, allowed_border=0):
    img_h, img_w = img_shape[:2]
    if allowed_border >= 0:
        return (valid_flags & (flat_anchors[:, 0] >= -allowed_border) &
                (flat_anchors[:, 1] >= -allowed_border) &
                (flat_anchors[:, 2] < img_w + allowed_border) &
                (flat_anchors[:, 3] < img_h + allowed_border))
    return valid_flags
