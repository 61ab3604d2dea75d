"""TwoStageDetector (mmdet/models/detectors/two_stage.py:9-214)."""
import os

import torch
import torch.nn as nn

from ..builder import DETECTORS, build_backbone, build_head, build_neck
from ...engine import get_store
from .base import BaseDetector


@DETECTORS.register_module()
class TwoStageDetector(BaseDetector):
    def __init__(self, backbone, neck=None, rpn_head=None, roi_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None):
        super().__init__()
        self.backbone = build_backbone(backbone)
        if neck is not None:
            self.neck = build_neck(neck)
        if rpn_head is not None:
            rpn_train_cfg = train_cfg.rpn if train_cfg is not None else None
            rpn_head_ = dict(rpn_head)
            rpn_head_.update(train_cfg=rpn_train_cfg, test_cfg=test_cfg.rpn)
            self.rpn_head = build_head(rpn_head_)
        if roi_head is not None:
            rcnn_train_cfg = train_cfg.rcnn if train_cfg is not None else None
            roi_head_ = dict(roi_head)
            roi_head_.update(train_cfg=rcnn_train_cfg)
            roi_head_.update(test_cfg=test_cfg.rcnn)
            self.roi_head = build_head(roi_head_)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.init_weights(pretrained=pretrained)

    @property
    def with_rpn(self):
        return hasattr(self, 'rpn_head') and self.rpn_head is not None

    @property
    def with_roi_head(self):
        return hasattr(self, 'roi_head') and self.roi_head is not None

    def init_weights(self, pretrained=None):
        self.backbone.init_weights(pretrained=pretrained)
        if self.with_neck:
            if isinstance(self.neck, nn.Sequential):
                for m in self.neck:
                    m.init_weights()
            else:
                self.neck.init_weights()
        if self.with_rpn:
            self.rpn_head.init_weights()
        if self.with_roi_head:
            self.roi_head.init_weights(pretrained)

    def _loft_trunk(self, store):
        t = self.__dict__.get('_trunk', False)
        if t is False:
            from ...trunk import Trunk
            ok = self.with_rpn and self.with_neck and self.training and Trunk.eligible(self)
            t = self.__dict__['_trunk'] = Trunk(self, store) if ok else None
        return t

    def prefetch(self, data):
        """Start the input-only work of the NEXT step (RPN anchor targets: assignment + sampling,
        ~3 ms of launch-thread time with a dozen host syncs) on the side stream now, while the GPU
        is still busy with this step's backward.  `data` is the next batch as it will be passed
        to forward_train (device-resident or staged by Trainer.stage)."""
        if not (self.with_rpn and hasattr(self.rpn_head, 'prefetch_targets')):
            return
        gt = data.get('gt_bboxes')
        img = data.get('img')
        if not gt or img is None or not all(t.is_cuda for t in gt):
            return
        self.rpn_head.prefetch_targets(gt, data['img_metas'], img.shape[-2:],
                                       data.get('ready_event'))

    def extract_feat(self, img):
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None,
                      gt_masks=None, proposals=None, **kwargs):
        store = get_store(self, img.device if img.is_cuda else None)
        store.begin_step()
        ready_event = kwargs.pop('ready_event', None)    # inputs staged on a copy stream
        prefetch_next = kwargs.pop('prefetch_next', None)    # the NEXT batch (Trainer.train_step)
        if ready_event is not None:
            torch.cuda.current_stream(store.device).wait_event(ready_event)
        if not img.is_cuda:
            img = img.to(store.device, non_blocking=True)
        dev = store.device
        gt_bboxes = [b.to(dev, non_blocking=True) for b in gt_bboxes]
        gt_labels = [l.to(dev, non_blocking=True) for l in gt_labels]
        for k, v in list(kwargs.items()):
            if isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], torch.Tensor):
                kwargs[k] = [t.to(dev, non_blocking=True) for t in v]
        # static part (backbone + FPN + RPN convs): recorded CUDA-graph programs when the model is
        # the R50-FPN-RPN composition (bonai_b200.trunk), else module by module.  The RPN targets
        # (assignment + sampling, several host syncs) are built on a side stream: with the trunk
        # they are issued AFTER the one-launch forward graph so the syncs hide under it.
        rpn_outs = None
        trunk = self._loft_trunk(store) if torch.is_grad_enabled() else None
        prefetch = self.with_rpn and hasattr(self.rpn_head, 'prefetch_targets') and \
            len(gt_bboxes) > 0 and not self.rpn_head.has_prefetched(gt_bboxes)
        proposal_cfg = self.train_cfg.get('rpn_proposal', self.test_cfg.rpn) if self.with_rpn \
            else None
        pre_proposals = None
        if trunk is not None:
            x, fused = trunk(img, img_metas, proposal_cfg)
            rpn_outs = self.rpn_head.outs_from_fused(fused)
            pre_proposals = trunk.proposals()
            if prefetch:
                self.rpn_head.prefetch_targets(gt_bboxes, img_metas, img.shape[-2:], ready_event)
        else:
            if prefetch:
                self.rpn_head.prefetch_targets(gt_bboxes, img_metas, img.shape[-2:], ready_event)
            x = self.extract_feat(img)
        losses = dict()
        if self.with_rpn:
            # with the recorded trunk the RPN loss is one launch that also writes the gradient of
            # the head outputs into the buffers the RPN backward program reads (no autograd nodes)
            direct = trunk is not None and trunk.direct_rpn_grads()
            rpn_losses, proposal_list = self.rpn_head.forward_train(
                x, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=gt_bboxes_ignore,
                proposal_cfg=proposal_cfg, rpn_outs=rpn_outs, proposals=pre_proposals,
                grad_out=trunk.current.gR if direct else None,
                after_loss=(lambda d: trunk.early_rpn_backward(d, 'side'))
                if trunk is not None and not direct else None)
        else:
            proposal_list = proposals
            rpn_losses = {}
            direct = False
        # input-only work of the NEXT step (its RPN anchor targets) goes to the side stream now:
        # the launch thread is about to wait for this step's forward graph in the RoI sampler
        # anyway, and at the start of the next step nothing then stands between the optimizer
        # launch and the next forward graph
        if prefetch_next is not None:
            self.prefetch(prefetch_next)
        rpn_bwd_at_sample = direct and os.environ.get('LOFT_RPN_BWD_AT', 'sample') == 'sample' and \
            hasattr(self.roi_head, 'offset_head')
        if rpn_bwd_at_sample:
            kwargs['after_sample'] = trunk.rpn_backward_direct
        roi_losses = self.roi_head.forward_train(x, img_metas, proposal_list, gt_bboxes, gt_labels,
                                                 gt_bboxes_ignore, gt_masks, **kwargs)
        if trunk is not None and rpn_losses:
            # RPN part of the backward: GEMM work for the GPU while the launch thread is busy --
            # right after the RoI sampler's host sync (default), or here while it sums the losses
            # and starts the autograd engine (bonai_b200.trunk.Trunk.early_rpn_backward)
            if direct:
                if not rpn_bwd_at_sample:
                    trunk.rpn_backward_direct()
            else:
                rpn_losses = trunk.early_rpn_backward(rpn_losses, 'end')
        losses.update(rpn_losses)
        losses.update(roi_losses)
        return losses

    @torch.no_grad()
    def simple_test_batch(self, img, img_metas, max_dets=None):
        """Batched test-time forward with device-resident results (see
        LoftRoIHead.simple_test_batch): img [B,3,H,W], one meta per tile."""
        store = get_store(self, img.device if img.is_cuda else None)
        store.refresh_weights()
        if not img.is_cuda:
            img = img.to(store.device, non_blocking=True)
        x = self.extract_feat(img)
        proposal_list = self.rpn_head.simple_test_rpn(x, img_metas)
        return self.roi_head.simple_test_batch(x, proposal_list, img_metas, max_dets=max_dets)

    @torch.no_grad()
    def simple_test(self, img, img_metas, proposals=None, rescale=False):
        """TwoStageDetector.simple_test (two_stage.py:187-199)."""
        assert self.with_bbox, 'Bbox head must be implemented.'
        store = get_store(self, img.device if img.is_cuda else None)
        store.refresh_weights()
        if not img.is_cuda:
            img = img.to(store.device, non_blocking=True)
        x = self.extract_feat(img)
        if proposals is None:
            proposal_list = self.rpn_head.simple_test_rpn(x, img_metas)
        else:
            proposal_list = proposals
        return self.roi_head.simple_test(x, proposal_list, img_metas, rescale=rescale)
