"""BaseDetector (mmdet/models/detectors/base.py:14-243): forward dispatch, loss parsing, train_step."""
from abc import ABCMeta
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn


class BaseDetector(nn.Module, metaclass=ABCMeta):
    def __init__(self):
        super().__init__()
        self.fp16_enabled = False

    @property
    def with_neck(self):
        return hasattr(self, 'neck') and self.neck is not None

    @property
    def with_shared_head(self):
        return hasattr(self, 'roi_head') and self.roi_head.with_shared_head

    @property
    def with_bbox(self):
        return ((hasattr(self, 'roi_head') and self.roi_head.with_bbox)
                or (hasattr(self, 'bbox_head') and self.bbox_head is not None))

    @property
    def with_mask(self):
        return ((hasattr(self, 'roi_head') and self.roi_head.with_mask)
                or (hasattr(self, 'mask_head') and self.mask_head is not None))

    def forward_test(self, imgs, img_metas, **kwargs):
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError(f'{name} must be a list, but got {type(var)}')
        if len(imgs) != len(img_metas):
            raise ValueError(f'num of augmentations ({len(imgs)}) != num of image meta '
                             f'({len(img_metas)})')
        assert imgs[0].size(0) == 1, 'only batch size 1 is supported at test time'
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        raise NotImplementedError('test-time augmentation is not on the LOFT path')

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def _parse_losses(self, losses):
        """base.py:175-208 -- but ONE packed device reduction / all-reduce and one read-back
        instead of a collective + `.item()` per key."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        packed = torch.stack([v.detach().reshape(()) for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized():
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        vals = packed.tolist()                         # the step's only host read-back
        for k, v in zip(list(log_vars.keys()), vals):
            log_vars[k] = v
        return loss, log_vars

    def train_step(self, data, optimizer):
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))

    def val_step(self, data, optimizer):
        return self.train_step(data, optimizer)
