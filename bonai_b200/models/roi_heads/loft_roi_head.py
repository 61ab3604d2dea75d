"""LoftRoIHead (mmdet/models/roi_heads/loft_roi_head.py:22-227): StandardRoIHead + the
roof-to-footprint offset branch."""
import os

import torch

from .builder_alias import HEADS, build_head, build_roi_extractor
from .standard_roi_head import StandardRoIHead
from ...core import bbox2roi


@HEADS.register_module()
class LoftRoIHead(StandardRoIHead):
    def __init__(self, offset_roi_extractor=None, offset_head=None, **kwargs):
        assert offset_head is not None
        super().__init__(**kwargs)
        if offset_head is not None:
            self.init_offset_head(offset_roi_extractor, offset_head)
        self.with_vis_feat = False

    def init_offset_head(self, offset_roi_extractor, offset_head):
        self.offset_roi_extractor = build_roi_extractor(offset_roi_extractor)
        self.offset_head = build_head(offset_head)

    def init_weights(self, pretrained):
        super().init_weights(pretrained)
        self.offset_head.init_weights()

    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels,
                      gt_bboxes_ignore=None, gt_masks=None, gt_offsets=None, after_sample=None):
        sampling_results = self.assign_and_sample(x, img_metas, proposal_list, gt_bboxes, gt_labels,
                                                  gt_bboxes_ignore)
        self._last_sampling_results = sampling_results
        if after_sample is not None:
            # the sampler's one host sync has just returned: the GPU is idle and the launch
            # thread has ~200 small launches ahead of it.  The caller queues independent GPU work
            # here (the RPN part of the backward, one graph launch, ~1.3 ms) to cover them.
            after_sample()
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
        if self.with_offset:
            offset_results = self._offset_forward_train(x, sampling_results,
                                                        bbox_results['bbox_feats'], gt_offsets,
                                                        img_metas,
                                                        offset_feats=bbox_results.get('offset_feats'),
                                                        offset_expanded=bbox_results.get(
                                                            'offset_expanded'))
            if offset_results['loss_offset'] is not None:
                losses.update(offset_results['loss_offset'])
        return losses

    def _mask_forward_train(self, x, sampling_results, bbox_feats, gt_masks, img_metas):
        """loft_roi_head.py:162-194 (no empty-positives early-out, unlike the parent)."""
        pos_rois = bbox2roi([res.pos_bboxes for res in sampling_results])
        mask_results = self._mask_forward(x, pos_rois)
        mask_targets = self.mask_head.get_targets(sampling_results, gt_masks, self.train_cfg)
        pos_labels = torch.cat([res.pos_gt_labels for res in sampling_results])
        loss_mask = self.mask_head.loss(mask_results['mask_pred'], mask_targets, pos_labels)
        mask_results.update(loss_mask=loss_mask, mask_targets=mask_targets)
        return mask_results

    def _shares_bbox_rois(self):
        """True when the offset extractor is the same RoIAlign as the bbox extractor (it is in
        every LOFT config: bonai_loft_foa_r50_fpn_basic.py:37-42,69-74), so that the offset
        features of the positive RoIs ARE the positives' rows of the bbox features."""
        a, b = self.bbox_roi_extractor, self.offset_roi_extractor
        if os.environ.get('LOFT_SHARE_OFFSET_ROI', '1') == '0' or self.with_shared_head:
            return False
        if type(a) is not type(b) or len(a.roi_layers) != len(b.roi_layers):
            return False
        return (repr(a.roi_layers) == repr(b.roi_layers) and
                list(a.featmap_strides) == list(b.featmap_strides) and
                a.out_channels == b.out_channels and
                getattr(a, 'finest_scale', None) == getattr(b, 'finest_scale', None))

    def _bbox_forward_train(self, x, sampling_results, gt_bboxes, gt_labels, img_metas):
        """StandardRoIHead._bbox_forward_train (standard_roi_head.py:111-124); additionally hands
        the positives' rows of the RoI features to the offset branch (`_TakeRows`) instead of
        running the offset extractor over the same RoIs again."""
        if not (self.with_offset and torch.is_grad_enabled() and self._shares_bbox_rois()):
            return super()._bbox_forward_train(x, sampling_results, gt_bboxes, gt_labels, img_metas)
        from ...ops.roi import take_rows, take_rows_rot
        rois = bbox2roi([res.bboxes for res in sampling_results])
        bbox_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        rows, off = [], 0
        for res in sampling_results:          # per image: positives first, then negatives
            n_pos, n_all = int(res.pos_bboxes.shape[0]), int(res.bboxes.shape[0])
            rows.append(torch.arange(off, off + n_pos, device=rois.device))
            off += n_all
        rows = torch.cat(rows)
        offset_feats = offset_expanded = None
        if rows.numel() > 0:
            ks = getattr(self.offset_head, 'branch_rotations', lambda: None)()
            if ks is not None:
                # gather + the FOA branch rotations + their concatenation in one launch; branch 0
                # is the identity, so the un-rotated features are its first P rows
                bbox_feats, offset_expanded = take_rows_rot(bbox_feats, rows, ks)
                offset_feats = offset_expanded[:rows.numel()]
            else:
                bbox_feats, offset_feats = take_rows(bbox_feats, rows)
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        bbox_results = dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats,
                            offset_feats=offset_feats, offset_expanded=offset_expanded)
        bbox_targets = self.bbox_head.get_targets(sampling_results, gt_bboxes, gt_labels,
                                                  self.train_cfg)
        loss_bbox = self.bbox_head.loss(cls_score, bbox_pred, rois, *bbox_targets)
        bbox_results.update(loss_bbox=loss_bbox)
        return bbox_results

    def _offset_forward_train(self, x, sampling_results, bbox_feats, gt_offsets, img_metas,
                              offset_feats=None, offset_expanded=None):
        if offset_expanded is not None:
            offset_results = dict(offset_pred=self.offset_head(offset_feats,
                                                               expanded=offset_expanded),
                                  offset_feats=offset_feats)
        elif offset_feats is not None:
            offset_results = dict(offset_pred=self.offset_head(offset_feats),
                                  offset_feats=offset_feats)
        else:
            pos_rois = bbox2roi([res.pos_bboxes for res in sampling_results])
            offset_results = self._offset_forward(x, pos_rois)
        offset_targets = self.offset_head.get_targets(sampling_results, gt_offsets, self.train_cfg)
        loss_offset = self.offset_head.loss(offset_results['offset_pred'], offset_targets)
        offset_results.update(loss_offset=loss_offset, offset_targets=offset_targets)
        return offset_results

    def _offset_forward(self, x, rois=None, pos_inds=None, bbox_feats=None):
        assert ((rois is not None) ^ (pos_inds is not None and bbox_feats is not None))
        if rois is not None:
            offset_feats = self.offset_roi_extractor(x[:self.offset_roi_extractor.num_inputs], rois)
        else:
            assert bbox_feats is not None
            offset_feats = bbox_feats[pos_inds]
        offset_pred = self.offset_head(offset_feats)
        return dict(offset_pred=offset_pred, offset_feats=offset_feats)

    # ------------------------------------------------------------------ batched device inference
    @torch.no_grad()
    def simple_test_batch(self, x, proposal_list, img_metas, max_dets=None):
        """Test-time RoI stage for a BATCH of tiles with every result left on the device: the
        per-image steps of LoftRoIHead.simple_test (loft_roi_head.py:196-227, test_mixins.py:53-72,
        152-177, 211-241) with the dense work batched across images -- one RoIAlign + bbox-head
        pass over all proposals, one mask-head and one FOA pass over all detections -- and only the
        per-image pieces (soft-NMS, mask paste) in a loop.  rescale=False semantics (scale 1).
        Returns per image (dets [k,5], labels [k], masks bool [k,H,W], offsets [k,2])."""
        from ...ops.infer import offset_fusion_decode, paste_masks
        cfg = self.test_cfg
        n_img = len(img_metas)
        rois = bbox2roi([p[:, :4] for p in proposal_list])
        bb = self._bbox_forward(x, rois)
        counts = [p.shape[0] for p in proposal_list]
        cls_l = bb['cls_score'].split(counts, 0)
        reg_l = bb['bbox_pred'].split(counts, 0)
        roi_l = rois.split(counts, 0)
        dets, labels = [], []
        for i in range(n_img):
            d, l = self.bbox_head.get_bboxes(roi_l[i], cls_l[i].contiguous(), reg_l[i].contiguous(),
                                             img_metas[i]['img_shape'], 1.0, rescale=False, cfg=cfg)
            if max_dets is not None:
                d, l = d[:max_dets], l[:max_dets]
            dets.append(d)
            labels.append(l)
        k = [d.shape[0] for d in dets]
        out = []
        if sum(k) == 0:
            for i in range(n_img):
                h, w = img_metas[i]['img_shape'][:2]
                out.append((dets[i], labels[i], dets[i].new_zeros((0, h, w), dtype=torch.bool),
                            dets[i].new_zeros((0, 2))))
            return out
        det_rois = bbox2roi([d[:, :4] for d in dets])
        mask_pred = self._mask_forward(x, det_rois)['mask_pred']
        off_feats = self.offset_roi_extractor(x[:self.offset_roi_extractor.num_inputs], det_rois)
        off_pred = self.offset_head(off_feats)                     # [4 * sum k, 2] branch-major
        fused_m = getattr(mask_pred, '_loft_fused', None)
        logits = fused_m.permute(0, 2, 3, 1)[..., 0] if fused_m is not None else mask_pred[:, 0]
        K = sum(k)
        off4 = off_pred.reshape(4, K, -1)
        o = 0
        for i in range(n_img):
            h, w = img_metas[i]['img_shape'][:2]
            sl = slice(o, o + k[i])
            masks = paste_masks(logits[sl], dets[i], h, w, float(cfg.mask_thr_binary))
            pred_i = off4[:, sl].reshape(4 * k[i], -1)
            offs = offset_fusion_decode(pred_i.contiguous(), dets[i], self.offset_head.offset_coder.stds,
                                        [1024, 1024])      # get_offsets' default img_shape
            out.append((dets[i], labels[i], masks, offs))
            o += k[i]
        return out

    # ------------------------------------------------------------------ inference
    def simple_test_offset(self, x, img_metas, det_bboxes, det_labels, rescale=False):
        """OffsetTestMixin.simple_test_offset (test_mixins.py:211-241)."""
        scale_factor = img_metas[0]['scale_factor']
        if det_bboxes.shape[0] == 0:
            return [[] for _ in range(2)]
        if rescale and not isinstance(scale_factor, float):
            scale_factor = torch.from_numpy(scale_factor).to(det_bboxes.device)
        _bboxes = det_bboxes[:, :4] * scale_factor if rescale else det_bboxes
        offset_rois = bbox2roi([_bboxes])
        offset_feats = self.offset_roi_extractor(x[:self.offset_roi_extractor.num_inputs],
                                                 offset_rois)
        offset_pred = self.offset_head(offset_feats)
        return self.offset_head.get_offsets(offset_pred.contiguous(), _bboxes, scale_factor, rescale)

    def simple_test(self, x, proposal_list, img_metas, proposals=None, rescale=False):
        """LoftRoIHead.simple_test (loft_roi_head.py:196-227): (bbox, segm, offset) results."""
        from ...core import bbox2result
        assert self.with_bbox, 'Bbox head must be implemented.'
        det_bboxes, det_labels = self.simple_test_bboxes(x, img_metas, proposal_list, self.test_cfg,
                                                         rescale=rescale)
        bbox_results = bbox2result(det_bboxes, det_labels, self.bbox_head.num_classes)
        segm_results = None
        if self.with_mask:
            segm_results = self.simple_test_mask(x, img_metas, det_bboxes, det_labels,
                                                 rescale=rescale)
        offset_results = self.simple_test_offset(x, img_metas, det_bboxes, det_labels,
                                                 rescale=rescale)
        return bbox_results, segm_results, offset_results
