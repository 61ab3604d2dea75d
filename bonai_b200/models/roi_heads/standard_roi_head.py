"""BaseRoIHead / StandardRoIHead (mmdet/models/roi_heads/base_roi_head.py:8-154,
standard_roi_head.py:10-290): assign + sample per image, bbox branch, mask branch."""
import os
from abc import ABCMeta

import torch
import torch.nn as nn

from .builder_alias import HEADS, build_head, build_roi_extractor, build_shared_head
from ...core import bbox2roi, build_assigner, build_sampler


class BaseRoIHead(nn.Module, metaclass=ABCMeta):
    def __init__(self, bbox_roi_extractor=None, bbox_head=None, mask_roi_extractor=None,
                 mask_head=None, shared_head=None, train_cfg=None, test_cfg=None):
        super().__init__()
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        if shared_head is not None:
            self.shared_head = build_shared_head(shared_head)
        if bbox_head is not None:
            self.init_bbox_head(bbox_roi_extractor, bbox_head)
        if mask_head is not None:
            self.init_mask_head(mask_roi_extractor, mask_head)
        self.init_assigner_sampler()

    @property
    def with_bbox(self):
        return hasattr(self, 'bbox_head') and self.bbox_head is not None

    @property
    def with_mask(self):
        return hasattr(self, 'mask_head') and self.mask_head is not None

    @property
    def with_offset(self):
        return hasattr(self, 'offset_head') and self.offset_head is not None

    @property
    def with_shared_head(self):
        return hasattr(self, 'shared_head') and self.shared_head is not None


@HEADS.register_module()
class StandardRoIHead(BaseRoIHead):
    def init_assigner_sampler(self):
        self.bbox_assigner = None
        self.bbox_sampler = None
        if self.train_cfg:
            self.bbox_assigner = build_assigner(self.train_cfg.assigner)
            self.bbox_sampler = build_sampler(self.train_cfg.sampler, context=self)

    def init_bbox_head(self, bbox_roi_extractor, bbox_head):
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        self.bbox_head = build_head(bbox_head)

    def init_mask_head(self, mask_roi_extractor, mask_head):
        if mask_roi_extractor is not None:
            self.mask_roi_extractor = build_roi_extractor(mask_roi_extractor)
            self.share_roi_extractor = False
        else:
            self.share_roi_extractor = True
            self.mask_roi_extractor = self.bbox_roi_extractor
        self.mask_head = build_head(mask_head)

    def init_weights(self, pretrained):
        if self.with_shared_head:
            self.shared_head.init_weights(pretrained=pretrained)
        if self.with_bbox:
            self.bbox_roi_extractor.init_weights()
            self.bbox_head.init_weights()
        if self.with_mask:
            self.mask_head.init_weights()
            if not self.share_roi_extractor:
                self.mask_roi_extractor.init_weights()

    def assign_and_sample(self, x, img_metas, proposal_list, gt_bboxes, gt_labels,
                          gt_bboxes_ignore=None):
        num_imgs = len(img_metas)
        if gt_bboxes_ignore is None:
            gt_bboxes_ignore = [None for _ in range(num_imgs)]
        from ...core.bbox import RandomSampler
        if type(self.bbox_sampler) is RandomSampler and self.bbox_sampler.neg_pos_ub < 0:
            if self._fused_sampling_ok(proposal_list, gt_bboxes, gt_bboxes_ignore):
                return self._assign_and_sample_fused(proposal_list, gt_bboxes, gt_labels)
            return self._assign_and_sample_batched(proposal_list, gt_bboxes, gt_labels,
                                                   gt_bboxes_ignore)
        sampling_results = []
        for i in range(num_imgs):
            assign_result = self.bbox_assigner.assign(proposal_list[i], gt_bboxes[i],
                                                      gt_bboxes_ignore[i], gt_labels[i])
            sampling_results.append(self.bbox_sampler.sample(
                assign_result, proposal_list[i], gt_bboxes[i], gt_labels[i]))
        return sampling_results

    def _fused_sampling_ok(self, proposal_list, gt_bboxes, gt_bboxes_ignore):
        """The one-launch sampler covers the LOFT configuration: MaxIoUAssigner without ignore
        regions, RandomSampler(add_gt_as_proposals) whose `random_choice` has not been replaced
        (parity tests inject the oracle's draws through it), equally sized proposal blocks."""
        from ...core.bbox import MaxIoUAssigner
        smp = self.bbox_sampler
        if os.environ.get('LOFT_FUSED_SAMPLER', '1') == '0' or 'random_choice' in vars(smp):
            return False
        if type(self.bbox_assigner) is not MaxIoUAssigner or not smp.add_gt_as_proposals:
            return False
        if any(g is not None and g.numel() > 0 for g in gt_bboxes_ignore):
            return False
        K = proposal_list[0].shape[0]
        if any(p.shape[0] != K or not p.is_cuda for p in proposal_list):
            return False
        return K + max(int(g.shape[0]) for g in gt_bboxes) <= 4096 and smp.num <= 1024

    def _assign_and_sample_fused(self, proposal_list, gt_bboxes, gt_labels):
        """assign + RandomSampler.sample for every image in three launches per image + one for
        the batch (`loft_iou_assign`, `loft_rcnn_sample`) and ONE host read-back (the per-image
        positive / negative counts); same candidate sets and sampling law as the reference
        (base_sampler.py:34-101, random_sampler.py:31-75), own random stream."""
        import ctypes
        from ... import _lib as L
        from ...core.bbox import SamplingResult
        i32 = ctypes.c_int
        asg, smp = self.bbox_assigner, self.bbox_sampler
        n_img = len(proposal_list)
        dev = proposal_list[0].device
        K, ld = int(proposal_list[0].shape[0]), int(proposal_list[0].shape[1])
        base = proposal_list[0]
        step = K * ld * base.element_size()
        if all(p.is_contiguous() and p.data_ptr() == base.data_ptr() + i * step
               for i, p in enumerate(proposal_list)):
            props_ptr = L.ptr(base)                      # slices of one [B,K,5] block: no copy
        else:
            stacked = torch.stack([p.contiguous() for p in proposal_list])
            props_ptr = L.ptr(stacked)
        nvs = [getattr(p, '_loft_num_valid', None) for p in proposal_list]
        nv = torch.stack([v.reshape(()) for v in nvs]).to(torch.int32) \
            if all(v is not None for v in nvs) else None
        gts = [g[:, :4].contiguous().float() for g in gt_bboxes]
        Gs = [int(g.shape[0]) for g in gts]
        gt_all = torch.cat(gts) if sum(Gs) else torch.zeros((1, 4), device=dev)
        offs = [0]
        for g in Gs:
            offs.append(offs[-1] + g)
        gt_off = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
        gt_inds = torch.empty((n_img, K), dtype=torch.long, device=dev)
        max_ov = torch.empty((n_img, K), dtype=torch.float32, device=dev)
        for i in range(n_img):
            if Gs[i] == 0 or K == 0:
                continue
            boxes = proposal_list[i][:, :4].contiguous()
            ws_bytes = int(L.lib().loft_iou_assign_workspace(L.ll(K), i32(Gs[i])))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            L.call('iou_assign', L.ptr(boxes), L.ll(K), L.ptr(gts[i]), i32(Gs[i]),
                   L.f32(asg.pos_iou_thr), L.f32(asg.neg_iou_thr), L.f32(asg.min_pos_iou),
                   i32(asg.match_low_quality), L.ptr(gt_inds[i]), L.ptr(max_ov[i]), L.ptr(ws),
                   ctypes.c_size_t(ws_bytes), L.stream())
        num = int(smp.num)
        sel = torch.empty((n_img, num), dtype=torch.long, device=dev)
        out_boxes = torch.empty((n_img, num, 4), dtype=torch.float32, device=dev)
        out_gt = torch.empty((n_img, num), dtype=torch.long, device=dev)
        out_isgt = torch.empty((n_img, num), dtype=torch.uint8, device=dev)
        cnt = torch.empty((n_img, 2), dtype=torch.int32, device=dev)
        self._sample_calls = getattr(self, '_sample_calls', 0) + 1
        seed = (int(torch.initial_seed()) * 1000003 + self._sample_calls) & ((1 << 64) - 1)
        L.call('rcnn_sample', props_ptr, L.ll(K * ld), i32(K), i32(ld), L.ptr(nv), L.ptr(gt_inds),
               L.ptr(gt_all), L.ptr(gt_off), i32(n_img), i32(max(Gs)), i32(num),
               i32(int(smp.num * smp.pos_fraction)), ctypes.c_ulonglong(seed), L.ptr(sel),
               L.ptr(out_boxes), L.ptr(out_gt), L.ptr(out_isgt), L.ptr(cnt), L.stream())
        counts = cnt.tolist()                                                  # the only sync
        return [SamplingResult.from_fused(sel[i], out_boxes[i], out_gt[i], out_isgt[i],
                                          counts[i][0], counts[i][1], gt_bboxes[i], gt_labels[i])
                for i in range(n_img)]

    def _assign_and_sample_batched(self, proposal_list, gt_bboxes, gt_labels, gt_bboxes_ignore):
        """Same result as assign + RandomSampler.sample per image (base_sampler.py:34-101), but
        all images share ONE host synchronisation: the assignments are computed for every image
        first, the positive / negative candidate lists come from a single `nonzero` over the
        stacked flags, and `unique()` of an index subset is a device-side sort."""
        from ...core.bbox import SamplingResult
        sampler = self.bbox_sampler
        n_img = len(proposal_list)
        ars, boxes_l, flags_l = [], [], []
        for i in range(n_img):
            ar = self.bbox_assigner.assign(proposal_list[i], gt_bboxes[i], gt_bboxes_ignore[i],
                                           gt_labels[i])
            nv = getattr(proposal_list[i], '_loft_num_valid', None)
            if nv is not None:     # fixed-size proposal block: rows >= nv are padding -> ignore
                ar.gt_inds = torch.where(
                    torch.arange(ar.gt_inds.numel(), device=ar.gt_inds.device) < nv, ar.gt_inds,
                    torch.full_like(ar.gt_inds, -1))
            bboxes = proposal_list[i][:, :4]
            gt_flags = bboxes.new_zeros((bboxes.shape[0],), dtype=torch.uint8)
            if sampler.add_gt_as_proposals and len(gt_bboxes[i]) > 0:
                bboxes = torch.cat([gt_bboxes[i], bboxes], dim=0)
                ar.add_gt_(gt_labels[i])
                gt_flags = torch.cat([bboxes.new_ones(gt_bboxes[i].shape[0], dtype=torch.uint8),
                                      gt_flags])
            ars.append(ar)
            boxes_l.append(bboxes)
            flags_l.append(gt_flags)
        sizes = [ar.gt_inds.numel() for ar in ars]
        allg = torch.cat([ar.gt_inds for ar in ars])
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        # code 0: positive, 1: negative candidates.  Per-image counts first (one read-back, the
        # only host sync of the RoI sampling), then ONE size-known nonzero for everything: rows
        # come out ordered by (code, position) = [pos img0, pos img1, ..., neg img0, neg img1, ...]
        flags = torch.stack([allg > 0, allg == 0])
        counts = torch.stack([flags[:, offs[i]:offs[i + 1]].sum(1) for i in range(n_img)], 1)
        counts = counts.reshape(-1).tolist()                                       # the only sync
        idx = torch.nonzero_static(flags, size=sum(counts))
        pos = idx[:, 1]
        chunks = torch.split(pos, counts)
        results = []
        num_expected_pos = int(sampler.num * sampler.pos_fraction)
        for i in range(n_img):
            off = offs[i]
            pos_inds = chunks[i] - off
            neg_inds = chunks[n_img + i] - off
            if pos_inds.numel() > num_expected_pos:
                pos_inds = sampler.random_choice(pos_inds, num_expected_pos).sort()[0]
            num_expected_neg = sampler.num - pos_inds.numel()
            if neg_inds.numel() > num_expected_neg:
                neg_inds = sampler.random_choice(neg_inds, num_expected_neg).sort()[0]
            results.append(SamplingResult(pos_inds, neg_inds, boxes_l[i], gt_bboxes[i], ars[i],
                                          flags_l[i]))
        return results

    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels,
                      gt_bboxes_ignore=None, gt_masks=None):
        sampling_results = self.assign_and_sample(x, img_metas, proposal_list, gt_bboxes, gt_labels,
                                                  gt_bboxes_ignore)
        losses = dict()
        if self.with_bbox:
            bbox_results = self._bbox_forward_train(x, sampling_results, gt_bboxes, gt_labels,
                                                    img_metas)
            losses.update(bbox_results['loss_bbox'])
        if self.with_mask:
            mask_results = self._mask_forward_train(x, sampling_results,
                                                    bbox_results['bbox_feats'], gt_masks,
                                                    img_metas)
            if mask_results['loss_mask'] is not None:
                losses.update(mask_results['loss_mask'])
        return losses

    def _bbox_forward(self, x, rois):
        bbox_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        if self.with_shared_head:
            bbox_feats = self.shared_head(bbox_feats)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats)

    def _bbox_forward_train(self, x, sampling_results, gt_bboxes, gt_labels, img_metas):
        rois = bbox2roi([res.bboxes for res in sampling_results])
        bbox_results = self._bbox_forward(x, rois)
        bbox_targets = self.bbox_head.get_targets(sampling_results, gt_bboxes, gt_labels,
                                                  self.train_cfg)
        loss_bbox = self.bbox_head.loss(bbox_results['cls_score'], bbox_results['bbox_pred'], rois,
                                        *bbox_targets)
        bbox_results.update(loss_bbox=loss_bbox)
        return bbox_results

    def _mask_forward_train(self, x, sampling_results, bbox_feats, gt_masks, img_metas):
        if not self.share_roi_extractor:
            pos_rois = bbox2roi([res.pos_bboxes for res in sampling_results])
            if pos_rois.shape[0] == 0:
                return dict(loss_mask=None)
            mask_results = self._mask_forward(x, pos_rois)
        else:
            pos_inds = []
            device = bbox_feats.device
            for res in sampling_results:
                pos_inds.append(torch.ones(res.pos_bboxes.shape[0], device=device,
                                           dtype=torch.bool))
                pos_inds.append(torch.zeros(res.neg_bboxes.shape[0], device=device,
                                            dtype=torch.bool))
            pos_inds = torch.cat(pos_inds)
            if pos_inds.shape[0] == 0:
                return dict(loss_mask=None)
            mask_results = self._mask_forward(x, pos_inds=pos_inds, bbox_feats=bbox_feats)
        mask_targets = self.mask_head.get_targets(sampling_results, gt_masks, self.train_cfg)
        pos_labels = torch.cat([res.pos_gt_labels for res in sampling_results])
        loss_mask = self.mask_head.loss(mask_results['mask_pred'], mask_targets, pos_labels)
        mask_results.update(loss_mask=loss_mask, mask_targets=mask_targets)
        return mask_results

    def _mask_forward(self, x, rois=None, pos_inds=None, bbox_feats=None):
        assert ((rois is not None) ^ (pos_inds is not None and bbox_feats is not None))
        if rois is not None:
            mask_feats = self.mask_roi_extractor(x[:self.mask_roi_extractor.num_inputs], rois)
            if self.with_shared_head:
                mask_feats = self.shared_head(mask_feats)
        else:
            assert bbox_feats is not None
            mask_feats = bbox_feats[pos_inds]
        mask_pred = self.mask_head(mask_feats)
        return dict(mask_pred=mask_pred, mask_feats=mask_feats)

    # ------------------------------------------------------------------ inference (test_mixins.py)
    def simple_test_bboxes(self, x, img_metas, proposals, rcnn_test_cfg, rescale=False):
        """BBoxTestMixin.simple_test_bboxes (test_mixins.py:53-72)."""
        rois = bbox2roi(proposals)
        bbox_results = self._bbox_forward(x, rois)
        img_shape = img_metas[0]['img_shape']
        scale_factor = img_metas[0]['scale_factor']
        return self.bbox_head.get_bboxes(rois, bbox_results['cls_score'].contiguous(),
                                         bbox_results['bbox_pred'].contiguous(), img_shape,
                                         scale_factor, rescale=rescale, cfg=rcnn_test_cfg)

    def simple_test_mask(self, x, img_metas, det_bboxes, det_labels, rescale=False):
        """MaskTestMixin.simple_test_mask (test_mixins.py:152-177)."""
        ori_shape = img_metas[0]['ori_shape']
        scale_factor = img_metas[0]['scale_factor']
        if det_bboxes.shape[0] == 0:
            return [[] for _ in range(self.mask_head.num_classes)]
        if rescale and not isinstance(scale_factor, float):
            scale_factor = torch.from_numpy(scale_factor).to(det_bboxes.device)
        _bboxes = det_bboxes[:, :4] * scale_factor if rescale else det_bboxes
        mask_rois = bbox2roi([_bboxes])
        mask_results = self._mask_forward(x, mask_rois)
        return self.mask_head.get_seg_masks(mask_results['mask_pred'], _bboxes,
                                            det_labels, self.test_cfg, ori_shape, scale_factor,
                                            rescale)

    def simple_test(self, x, proposal_list, img_metas, proposals=None, rescale=False):
        """StandardRoIHead.simple_test (standard_roi_head.py:218-237)."""
        from ...core import bbox2result
        det_bboxes, det_labels = self.simple_test_bboxes(x, img_metas, proposal_list, self.test_cfg,
                                                         rescale=rescale)
        bbox_results = bbox2result(det_bboxes, det_labels, self.bbox_head.num_classes)
        if not self.with_mask:
            return bbox_results
        return bbox_results, self.simple_test_mask(x, img_metas, det_bboxes, det_labels,
                                                   rescale=rescale)
