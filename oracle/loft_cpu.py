"""ORACLE (test infrastructure, not product code) -- portable CPU restatement of the reference's
LOFT/FOA training step in plain PyTorch fp32, written functionally over the reference's
``state_dict`` (same key names / shapes).

Every function cites the reference file:line (paths under jwwangchn/BONAI @ aeafa46) it follows.
The restatement is pinned against the UNMODIFIED reference executed over the import shim
(oracle/shim, oracle/make_golden.py -> tests/golden/*.npz; tests/test_oracle_vs_reference.py runs
both side by side when /root/reference is present).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs) may
import this module; nothing under bonai_b200/ does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_cpu

# ----------------------------------------------------------------------------- config constants
# configs/_base_/models/bonai_loft_foa_r50_fpn_basic.py
CFG = dict(
    anchor_scales=[8], anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[4, 8, 16, 32, 64],   # :23-27
    rpn_assign=dict(pos=0.7, neg=0.3, min_pos=0.3),                                        # :86-93
    rpn_sampler=dict(num=512, pos_fraction=0.5, add_gt=False),                             # :94-99
    rpn_proposal=dict(nms_pre=3000, nms_post=3000, nms_thr=0.7, min_bbox_size=0),          # :103-109
    rcnn_assign=dict(pos=0.5, neg=0.5, min_pos=0.5),                                       # :111-118
    rcnn_sampler=dict(num=1024, pos_fraction=0.25, add_gt=True),                           # :119-124
    mask_size=28,                                                                          # :125
    roi_strides=[4, 8, 16, 32], finest_scale=56,
    rpn_stds=(1.0, 1.0, 1.0, 1.0), rcnn_stds=(0.1, 0.1, 0.2, 0.2),                         # :28-31,48-51
    rotations=[0, 90, 180, 270], offset_stds=(0.5, 0.5), loss_offset_weight=16.0,          # :75-82
    num_classes=1,
)


# ----------------------------------------------------------------------------- backbone / neck
def _bn_eval(x, p, prefix, eps=1e-5):
    """BatchNorm2d in eval mode (norm_eval=True, backbones/resnet.py:640-649)."""
    return F.batch_norm(x, p[prefix + '.running_mean'], p[prefix + '.running_var'],
                        p[prefix + '.weight'], p[prefix + '.bias'], False, 0.0, eps)


def _bottleneck(x, p, prefix, stride, has_down):
    """Bottleneck.forward, style='pytorch' (stride on the 3x3), resnet.py:260-300."""
    out = F.relu(_bn_eval(F.conv2d(x, p[prefix + '.conv1.weight']), p, prefix + '.bn1'))
    out = F.relu(_bn_eval(F.conv2d(out, p[prefix + '.conv2.weight'], stride=stride, padding=1),
                          p, prefix + '.bn2'))
    out = _bn_eval(F.conv2d(out, p[prefix + '.conv3.weight']), p, prefix + '.bn3')
    identity = x
    if has_down:
        identity = _bn_eval(F.conv2d(x, p[prefix + '.downsample.0.weight'], stride=stride),
                            p, prefix + '.downsample.1')
    return F.relu(out + identity)


def resnet50(img, p, prefix='backbone'):
    """ResNet.forward, depth 50, out_indices (0,1,2,3), resnet.py:623-638."""
    x = F.conv2d(img, p[prefix + '.conv1.weight'], stride=2, padding=3)
    x = F.relu(_bn_eval(x, p, prefix + '.bn1'))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, (nblocks, stride) in enumerate(zip([3, 4, 6, 3], [1, 2, 2, 2])):
        for b in range(nblocks):
            x = _bottleneck(x, p, f'{prefix}.layer{li + 1}.{b}', stride if b == 0 else 1, b == 0)
        outs.append(x)
    return outs


def fpn(feats, p, prefix='neck'):
    """FPN.forward (num_outs=5, no extra convs), necks/fpn.py:164-216."""
    lat = [F.conv2d(f, p[f'{prefix}.lateral_convs.{i}.conv.weight'],
                    p[f'{prefix}.lateral_convs.{i}.conv.bias']) for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    outs = [F.conv2d(l, p[f'{prefix}.fpn_convs.{i}.conv.weight'],
                     p[f'{prefix}.fpn_convs.{i}.conv.bias'], padding=1) for i, l in enumerate(lat)]
    outs.append(F.max_pool2d(outs[-1], 1, stride=2))   # fpn.py:197-199
    return outs


# ----------------------------------------------------------------------------- anchors / coders
def base_anchors(base_size, scales, ratios):
    """AnchorGenerator.gen_single_level_base_anchors (center_offset=0), anchor_generator.py:142-185."""
    scales = torch.tensor(scales, dtype=torch.float32)
    ratios = torch.tensor(ratios, dtype=torch.float32)
    h_ratios = torch.sqrt(ratios)
    w_ratios = 1 / h_ratios
    ws = (base_size * w_ratios[:, None] * scales[None, :]).view(-1)
    hs = (base_size * h_ratios[:, None] * scales[None, :]).view(-1)
    return torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], dim=-1)


def grid_anchors(featmap_sizes, cfg=CFG):
    """AnchorGenerator.grid_anchors / single_level_grid_anchors, anchor_generator.py:206-271."""
    out = []
    for (fh, fw), stride in zip(featmap_sizes, cfg['anchor_strides']):
        ba = base_anchors(stride, cfg['anchor_scales'], cfg['anchor_ratios'])
        sx = torch.arange(0, fw, dtype=torch.float32) * stride
        sy = torch.arange(0, fh, dtype=torch.float32) * stride
        xx = sx.repeat(len(sy))
        yy = sy.view(-1, 1).repeat(1, len(sx)).view(-1)
        shifts = torch.stack([xx, yy, xx, yy], dim=-1)
        out.append((ba[None, :, :] + shifts[:, None, :]).view(-1, 4))
    return out


def bbox2delta(proposals, gt, stds):
    """delta_xywh_bbox_coder.py:74-116 (means 0)."""
    px = (proposals[..., 0] + proposals[..., 2]) * 0.5
    py = (proposals[..., 1] + proposals[..., 3]) * 0.5
    pw = proposals[..., 2] - proposals[..., 0]
    ph = proposals[..., 3] - proposals[..., 1]
    gx = (gt[..., 0] + gt[..., 2]) * 0.5
    gy = (gt[..., 1] + gt[..., 3]) * 0.5
    gw = gt[..., 2] - gt[..., 0]
    gh = gt[..., 3] - gt[..., 1]
    deltas = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph)],
                         dim=-1)
    return deltas.sub_(deltas.new_zeros(4).unsqueeze(0)).div_(deltas.new_tensor(stds).unsqueeze(0))


def delta2bbox(rois, deltas, stds, max_shape=None, wh_ratio_clip=16 / 1000):
    """delta_xywh_bbox_coder.py:119-197 (means 0)."""
    means = deltas.new_zeros(4).repeat(1, deltas.size(1) // 4)
    stds_t = deltas.new_tensor(stds).repeat(1, deltas.size(1) // 4)
    d = deltas * stds_t + means
    dx, dy, dw, dh = d[:, 0::4], d[:, 1::4], d[:, 2::4], d[:, 3::4]
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dh)
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = px + pw * dx
    gy = py + ph * dy
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    if max_shape is not None:
        x1 = x1.clamp(min=0, max=max_shape[1])
        y1 = y1.clamp(min=0, max=max_shape[0])
        x2 = x2.clamp(min=0, max=max_shape[1])
        y2 = y2.clamp(min=0, max=max_shape[0])
    return torch.stack([x1, y1, x2, y2], dim=-1).view_as(deltas)


def offset2delta(proposals, gt, stds=(0.5, 0.5)):
    """delta_xy_offset_coder.py:46-65."""
    pw = proposals[..., 2] - proposals[..., 0]
    ph = proposals[..., 3] - proposals[..., 1]
    deltas = torch.stack([gt[..., 0] / pw, gt[..., 1] / ph], dim=-1)
    return deltas.sub_(deltas.new_zeros(2).unsqueeze(0)).div_(deltas.new_tensor(stds).unsqueeze(0))


def delta2offset(rois, deltas, stds=(0.5, 0.5), max_shape=None):
    """delta_xy_offset_coder.py:67-88."""
    d = deltas * deltas.new_tensor(stds).repeat(1, deltas.size(1) // 2)
    dx, dy = d[:, 0::2], d[:, 1::2]
    pw = (rois[:, 2] - rois[:, 0]).unsqueeze(1).expand_as(dx)
    ph = (rois[:, 3] - rois[:, 1]).unsqueeze(1).expand_as(dy)
    gx, gy = pw * dx, ph * dy
    if max_shape is not None:
        gx = gx.clamp(min=-max_shape[1], max=max_shape[1])
        gy = gy.clamp(min=-max_shape[0], max=max_shape[0])
    return torch.stack([gx, gy], dim=-1).view_as(deltas)


# ----------------------------------------------------------------------------- assign / sample
def bbox_overlaps(b1, b2, eps=1e-6):
    """bbox_overlaps(mode='iou', is_aligned=False), iou2d_calculator.py:39-130."""
    rows, cols = b1.size(0), b2.size(0)
    if rows * cols == 0:
        return b1.new_zeros(rows, cols)
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[:, :, 0] * wh[:, :, 1]
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = a1[:, None] + a2 - overlap
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


def max_iou_assign(bboxes, gt_bboxes, pos_thr, neg_thr, min_pos, gt_labels=None):
    """MaxIoUAssigner.assign / assign_wrt_overlaps (match_low_quality, gt_max_assign_all),
    max_iou_assigner.py:60-212.  Returns (gt_inds[n] (0 bg, -1 ignore, i+1 matched),
    max_overlaps[n], labels[n] or None)."""
    overlaps = bbox_overlaps(gt_bboxes, bboxes)
    num_gts, num_bboxes = overlaps.size(0), overlaps.size(1)
    gt_inds = overlaps.new_full((num_bboxes,), -1, dtype=torch.long)
    if num_gts == 0 or num_bboxes == 0:
        max_overlaps = overlaps.new_zeros((num_bboxes,))
        if num_gts == 0:
            gt_inds[:] = 0
        labels = None if gt_labels is None else overlaps.new_full((num_bboxes,), -1,
                                                                  dtype=torch.long)
        return gt_inds, max_overlaps, labels
    max_overlaps, argmax_overlaps = overlaps.max(dim=0)
    gt_max_overlaps, _ = overlaps.max(dim=1)
    gt_inds[(max_overlaps >= 0) & (max_overlaps < neg_thr)] = 0
    pos = max_overlaps >= pos_thr
    gt_inds[pos] = argmax_overlaps[pos] + 1
    for i in range(num_gts):                               # :193-199
        if gt_max_overlaps[i] >= min_pos:
            gt_inds[overlaps[i, :] == gt_max_overlaps[i]] = i + 1
    labels = None
    if gt_labels is not None:
        labels = gt_inds.new_full((num_bboxes,), -1)
        pos_inds = torch.nonzero(gt_inds > 0, as_tuple=False).squeeze()
        if pos_inds.numel() > 0:
            labels[pos_inds] = gt_labels[gt_inds[pos_inds] - 1]
    return gt_inds, max_overlaps, labels


def _random_choice(gallery, num, forced):
    """RandomSampler.random_choice, random_sampler.py:31-55 (torch.randperm on gallery.device)."""
    if forced is not None:
        return forced.pop(0)
    perm = torch.randperm(gallery.numel())[:num]
    return gallery[perm]


def random_sample(gt_inds, bboxes, gt_bboxes, gt_labels, labels, num, pos_fraction, add_gt,
                  forced=None, record=None):
    """BaseSampler.sample + RandomSampler + SamplingResult, base_sampler.py:34-101,
    random_sampler.py:57-75, sampling_result.py:25-54, assign_result.py:190-204."""
    if add_gt and gt_bboxes.size(0) > 0:
        bboxes = torch.cat([gt_bboxes, bboxes], dim=0)
        self_inds = torch.arange(1, gt_bboxes.size(0) + 1, dtype=torch.long)
        gt_inds = torch.cat([self_inds, gt_inds])
        if labels is not None:
            labels = torch.cat([gt_labels, labels])
    num_expected_pos = int(num * pos_fraction)
    pos_inds = torch.nonzero(gt_inds > 0, as_tuple=False)
    if pos_inds.numel() != 0:
        pos_inds = pos_inds.squeeze(1)
    if pos_inds.numel() > num_expected_pos:
        pos_inds = _random_choice(pos_inds, num_expected_pos, forced)
        if record is not None:
            record.append(pos_inds.clone())
    pos_inds = pos_inds.unique()
    num_expected_neg = num - pos_inds.numel()
    neg_inds = torch.nonzero(gt_inds == 0, as_tuple=False)
    if neg_inds.numel() != 0:
        neg_inds = neg_inds.squeeze(1)
    if len(neg_inds) > num_expected_neg:
        neg_inds = _random_choice(neg_inds, num_expected_neg, forced)
        if record is not None:
            record.append(neg_inds.clone())
    neg_inds = neg_inds.unique()
    res = dict(pos_inds=pos_inds, neg_inds=neg_inds, pos_bboxes=bboxes[pos_inds],
               neg_bboxes=bboxes[neg_inds], pos_assigned_gt_inds=gt_inds[pos_inds] - 1)
    if gt_bboxes.numel() == 0:
        res['pos_gt_bboxes'] = torch.empty_like(gt_bboxes).view(-1, 4)
    else:
        res['pos_gt_bboxes'] = gt_bboxes[res['pos_assigned_gt_inds'], :]
    res['pos_gt_labels'] = labels[pos_inds] if labels is not None else None
    res['bboxes'] = torch.cat([res['pos_bboxes'], res['neg_bboxes']])
    return res


# ----------------------------------------------------------------------------- RPN
def rpn_forward(feats, p, prefix='rpn_head'):
    """RPNHead.forward_single on every level (shared weights), rpn_head.py:38-44."""
    cls, reg = [], []
    for x in feats:
        x = F.relu(F.conv2d(x, p[prefix + '.rpn_conv.weight'], p[prefix + '.rpn_conv.bias'],
                            padding=1))
        cls.append(F.conv2d(x, p[prefix + '.rpn_cls.weight'], p[prefix + '.rpn_cls.bias']))
        reg.append(F.conv2d(x, p[prefix + '.rpn_reg.weight'], p[prefix + '.rpn_reg.bias']))
    return cls, reg


def rpn_loss(cls_scores, bbox_preds, gt_bboxes, cfg=CFG, forced=None, record=None):
    """AnchorHead.loss / get_targets / _get_targets_single / loss_single,
    anchor_head.py:180-497 with RPNHead.loss rpn_head.py:46-77
    (allowed_border=-1 => every anchor is inside; sampling=True)."""
    sizes = [c.shape[-2:] for c in cls_scores]
    mlvl = grid_anchors(sizes, cfg)
    num_lvl = [a.size(0) for a in mlvl]
    flat = torch.cat(mlvl)
    n_img = cls_scores[0].size(0)
    labels_l, lw_l, bt_l, bw_l = [], [], [], []
    num_pos_tot = num_neg_tot = 0
    for i in range(n_img):
        a = cfg['rpn_assign']
        gt_inds, _, _ = max_iou_assign(flat, gt_bboxes[i], a['pos'], a['neg'], a['min_pos'])
        s = cfg['rpn_sampler']
        sr = random_sample(gt_inds, flat, gt_bboxes[i], None, None, s['num'], s['pos_fraction'],
                           s['add_gt'], forced, record)
        n = flat.size(0)
        labels = flat.new_full((n,), 0, dtype=torch.long)        # background_label = 0
        label_weights = flat.new_zeros(n)
        bbox_targets = torch.zeros_like(flat)
        bbox_weights = torch.zeros_like(flat)
        pos, neg = sr['pos_inds'], sr['neg_inds']
        if len(pos) > 0:
            bbox_targets[pos, :] = bbox2delta(sr['pos_bboxes'], sr['pos_gt_bboxes'],
                                              cfg['rpn_stds'])
            bbox_weights[pos, :] = 1.0
            labels[pos] = 1                                        # anchor_head.py:252-254
            label_weights[pos] = 1.0
        if len(neg) > 0:
            label_weights[neg] = 1.0
        num_pos_tot += max(pos.numel(), 1)
        num_neg_tot += max(neg.numel(), 1)
        labels_l.append(labels)
        lw_l.append(label_weights)
        bt_l.append(bbox_targets)
        bw_l.append(bbox_weights)
    num_total_samples = num_pos_tot + num_neg_tot

    def to_levels(t):                                              # anchor/utils.py:4-17
        t = torch.stack(t, 0)
        out, s = [], 0
        for n in num_lvl:
            out.append(t[:, s:s + n])
            s += n
        return out

    labels_l, lw_l, bt_l, bw_l = map(to_levels, (labels_l, lw_l, bt_l, bw_l))
    loss_cls, loss_bbox = [], []
    for l in range(len(cls_scores)):
        lab = labels_l[l].reshape(-1)
        lw = lw_l[l].reshape(-1)
        cs = cls_scores[l].permute(0, 2, 3, 1).reshape(-1, 1)
        # binary_cross_entropy with _expand_binary_labels, cross_entropy_loss.py:42-91
        bin_lab = (lab >= 1).float().view(-1, 1)
        l_cls = F.binary_cross_entropy_with_logits(cs, bin_lab, reduction='none')
        loss_cls.append((l_cls * lw.view(-1, 1)).sum() / num_total_samples)
        bp = bbox_preds[l].permute(0, 2, 3, 1).reshape(-1, 4)
        l_box = torch.abs(bp - bt_l[l].reshape(-1, 4)) * bw_l[l].reshape(-1, 4)   # l1_loss
        loss_bbox.append(l_box.sum() / num_total_samples)
    return loss_cls, loss_bbox


def rpn_get_bboxes(cls_scores, bbox_preds, img_shapes, cfg=CFG, stable_sort=False, nms_fn=None):
    """AnchorHead.get_bboxes -> RPNHead._get_bboxes_single, rpn_head.py:79-168."""
    pc = cfg['rpn_proposal']
    sizes = [c.shape[-2:] for c in cls_scores]
    mlvl = grid_anchors(sizes, cfg)
    out = []
    for i in range(cls_scores[0].size(0)):
        lvl_ids, sc_l, bp_l, an_l = [], [], [], []
        for l in range(len(cls_scores)):
            scores = cls_scores[l][i].detach().permute(1, 2, 0).reshape(-1).sigmoid()
            bp = bbox_preds[l][i].detach().permute(1, 2, 0).reshape(-1, 4)
            anchors = mlvl[l]
            if pc['nms_pre'] > 0 and scores.shape[0] > pc['nms_pre']:
                ranked, rank_inds = scores.sort(descending=True, stable=stable_sort)
                topk = rank_inds[:pc['nms_pre']]
                scores = ranked[:pc['nms_pre']]
                bp = bp[topk, :]
                anchors = anchors[topk, :]
            sc_l.append(scores)
            bp_l.append(bp)
            an_l.append(anchors)
            lvl_ids.append(scores.new_full((scores.size(0),), l, dtype=torch.long))
        scores = torch.cat(sc_l)
        anchors = torch.cat(an_l)
        bp = torch.cat(bp_l)
        proposals = delta2bbox(anchors, bp, cfg['rpn_stds'], max_shape=img_shapes[i])
        ids = torch.cat(lvl_ids)
        fn = nms_fn or ops_cpu.batched_nms
        dets, _ = fn(proposals, scores, ids, pc['nms_thr'])
        out.append(dets[:pc['nms_post']])
    return out


# ----------------------------------------------------------------------------- RoI heads
def bbox2roi(bbox_list):
    """core/bbox/transforms.py:54-73."""
    rois = []
    for i, b in enumerate(bbox_list):
        if b.size(0) > 0:
            rois.append(torch.cat([b.new_full((b.size(0), 1), i), b[:, :4]], dim=-1))
        else:
            rois.append(b.new_zeros((0, 5)))
    return torch.cat(rois, 0)


def map_roi_levels(rois, num_levels, finest_scale=56):
    """single_level_roi_extractor.py:32-51."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lvls.clamp(min=0, max=num_levels - 1).long()


def roi_extract(feats, rois, out_size, cfg=CFG, roi_align_fn=None):
    """SingleRoIExtractor.forward, single_level_roi_extractor.py:53-80."""
    fn = roi_align_fn or ops_cpu.roi_align
    nl = len(cfg['roi_strides'])
    out = feats[0].new_zeros(rois.size(0), feats[0].size(1), out_size, out_size)
    lvls = map_roi_levels(rois, nl, cfg['finest_scale'])
    for i in range(nl):
        inds = lvls == i
        if inds.any():
            out[inds] = fn(feats[i], rois[inds, :], out_size, 1.0 / cfg['roi_strides'][i], 0, True)
    return out


def bbox_head_forward(x, p, prefix='roi_head.bbox_head'):
    """Shared2FCBBoxHead.forward, convfc_bbox_head.py:135-173."""
    x = x.flatten(1)
    for i in range(2):
        x = F.relu(F.linear(x, p[f'{prefix}.shared_fcs.{i}.weight'],
                            p[f'{prefix}.shared_fcs.{i}.bias']))
    return (F.linear(x, p[prefix + '.fc_cls.weight'], p[prefix + '.fc_cls.bias']),
            F.linear(x, p[prefix + '.fc_reg.weight'], p[prefix + '.fc_reg.bias']))


def bbox_head_loss(cls_score, bbox_pred, samples, cfg=CFG):
    """BBoxHead.get_targets + loss, bbox_head.py:84-185; CE cross_entropy_loss.py:9-39;
    accuracy accuracy.py:4-48."""
    nc = cfg['num_classes']
    labels_l, lw_l, bt_l, bw_l = [], [], [], []
    for sr in samples:
        npos, nneg = sr['pos_bboxes'].size(0), sr['neg_bboxes'].size(0)
        n = npos + nneg
        labels = sr['pos_bboxes'].new_full((n,), nc, dtype=torch.long)
        lw = sr['pos_bboxes'].new_zeros(n)
        bt = sr['pos_bboxes'].new_zeros(n, 4)
        bw = sr['pos_bboxes'].new_zeros(n, 4)
        if npos > 0:
            labels[:npos] = sr['pos_gt_labels']
            lw[:npos] = 1.0
            bt[:npos, :] = bbox2delta(sr['pos_bboxes'], sr['pos_gt_bboxes'], cfg['rcnn_stds'])
            bw[:npos, :] = 1
        if nneg > 0:
            lw[-nneg:] = 1.0
        labels_l.append(labels)
        lw_l.append(lw)
        bt_l.append(bt)
        bw_l.append(bw)
    labels, lw, bt, bw = map(lambda t: torch.cat(t, 0), (labels_l, lw_l, bt_l, bw_l))
    losses = {}
    avg_factor = max(torch.sum(lw > 0).float().item(), 1.)
    if cls_score.numel() > 0:
        ce = F.cross_entropy(cls_score, labels, reduction='none')
        losses['loss_cls'] = (ce * lw).sum() / avg_factor
        pred_label = cls_score.topk(1, dim=1)[1].t()
        correct = pred_label.eq(labels.view(1, -1).expand_as(pred_label))
        losses['acc'] = correct[:1].reshape(-1).float().sum(0, keepdim=True).mul_(
            100.0 / cls_score.size(0))
    pos_inds = (labels >= 0) & (labels < nc)
    if pos_inds.any():
        pos_pred = bbox_pred.view(bbox_pred.size(0), -1, 4)[pos_inds.type(torch.bool),
                                                            labels[pos_inds.type(torch.bool)]]
        l1 = torch.abs(pos_pred - bt[pos_inds]) * bw[pos_inds]
        losses['loss_bbox'] = l1.sum() / bt.size(0)                 # bbox_head.py:181
    else:
        losses['loss_bbox'] = bbox_pred.sum() * 0
    return losses


def mask_head_forward(x, p, prefix='roi_head.mask_head'):
    """FCNMaskHead.forward, fcn_mask_head.py:117-126."""
    for i in range(4):
        x = F.relu(F.conv2d(x, p[f'{prefix}.convs.{i}.conv.weight'],
                            p[f'{prefix}.convs.{i}.conv.bias'], padding=1))
    x = F.relu(F.conv_transpose2d(x, p[prefix + '.upsample.weight'], p[prefix + '.upsample.bias'],
                                  stride=2))
    return F.conv2d(x, p[prefix + '.conv_logits.weight'], p[prefix + '.conv_logits.bias'])


def mask_target(samples, gt_masks, cfg=CFG, roi_align_fn=None):
    """mask_target / mask_target_single + BitmapMasks.crop_and_resize,
    core/mask/mask_target.py:6-62, core/mask/structures.py:261-291.
    gt_masks: list of uint8 arrays/tensors [G,H,W]."""
    fn = roi_align_fn or ops_cpu.roi_align
    ms = cfg['mask_size']
    out = []
    for sr, m in zip(samples, gt_masks):
        props = sr['pos_bboxes']
        if props.size(0) == 0:
            out.append(props.new_zeros((0, ms, ms)))
            continue
        m = torch.as_tensor(m)
        maxh, maxw = m.shape[-2:]
        pn = props.detach().cpu().numpy().copy()
        pn[:, [0, 2]] = np.clip(pn[:, [0, 2]], 0, maxw)
        pn[:, [1, 3]] = np.clip(pn[:, [1, 3]], 0, maxh)
        b = torch.from_numpy(pn)
        nb = b.shape[0]
        rois = torch.cat([torch.arange(nb, dtype=b.dtype)[:, None], b], dim=1)
        sel = m.index_select(0, sr['pos_assigned_gt_inds']).to(dtype=rois.dtype)
        t = fn(sel[:, None, :, :], rois, ms, 1.0, 0, True).squeeze(1)
        out.append((t >= 0.5).float())
    return torch.cat(out) if out else torch.zeros((0, ms, ms))


def rotate_feature(x, angle_deg):
    """OffsetHeadExpandFeature.expand_feature (affine_grid + grid_sample),
    offset_head_expand_feature.py:163-196."""
    theta = torch.zeros((x.size(0), 2, 3))
    a = angle_deg * math.pi / 180.0
    theta[:, 0, 0] = math.cos(a)
    theta[:, 0, 1] = math.sin(-a)
    theta[:, 1, 0] = math.sin(a)
    theta[:, 1, 1] = math.cos(a)
    grid = F.affine_grid(theta, x.size(), align_corners=False)
    return F.grid_sample(x, grid, align_corners=False)


def offset_head_forward(x, p, cfg=CFG, prefix='roi_head.offset_head'):
    """OffsetHeadExpandFeature.forward (share_expand_fc=True), offset_head_expand_feature.py:134-161."""
    if x.size(0) == 0:
        return x.new_empty(0, 2 * len(cfg['rotations']))
    inp = x.clone()
    outs = []
    for idx, rot in enumerate(cfg['rotations']):
        y = rotate_feature(inp, rot)
        for c in range(10):
            y = F.relu(F.conv2d(y, p[f'{prefix}.expand_convs.{idx}.{c}.weight'],
                                p[f'{prefix}.expand_convs.{idx}.{c}.bias'], padding=1))
        y = y.view(y.size(0), -1)
        for i in range(2):
            y = F.relu(F.linear(y, p[f'{prefix}.fcs.{i}.weight'], p[f'{prefix}.fcs.{i}.bias']))
        outs.append(F.linear(y, p[prefix + '.fc_offset.weight'], p[prefix + '.fc_offset.bias']))
    return torch.cat(outs, 0)


def _offset_rotate(offset, angle_deg):
    """offset_rotate via polar coordinates in float64 `math`, offset_head_expand_feature.py:207-247."""
    ox, oy = offset
    length = math.sqrt(ox ** 2 + oy ** 2)
    ang = math.atan2(oy, ox) - angle_deg * math.pi / 180.0
    return [length * math.cos(ang), length * math.sin(ang)]


def offset_targets(samples, gt_offsets, cfg=CFG):
    """OffsetHeadExpandFeature.get_targets / _offset_target_single,
    offset_head_expand_feature.py:271-344."""
    out = []
    for rot in cfg['rotations']:
        per_img = []
        for sr, go in zip(samples, gt_offsets):
            props = sr['pos_bboxes']
            if props.size(0) == 0:
                per_img.append(props.new_zeros((0, 2)))
                continue
            inds = sr['pos_assigned_gt_inds'].cpu().numpy()
            rows = [_offset_rotate(go[inds[i]].tolist(), rot) for i in range(props.size(0))]
            pg = torch.from_numpy(np.stack(np.array(rows))).float()
            if rot in (90, 270):
                t = offset2delta(props, pg[:, [1, 0]], cfg['offset_stds'])[:, [1, 0]]
            else:
                t = offset2delta(props, pg, cfg['offset_stds'])
            per_img.append(t)
        out.append(torch.cat(per_img, 0))
    return torch.cat(out, 0)


def smooth_l1_mean(pred, target, beta=1.0):
    """smooth_l1_loss with reduction='mean', smooth_l1_loss.py:8-26."""
    diff = torch.abs(pred - target)
    return torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta).mean()


def offset_fusion_max(offset_pred):
    """offset_fusion(model='max'), 4 branches, offset_head_expand_feature.py:346-413."""
    s = offset_pred.split(offset_pred.shape[0] // 4, dim=0)
    vx = torch.stack([s[0][:, 0], s[1][:, 1], s[2][:, 0], s[3][:, 1]], dim=1).abs().max(dim=1)[0]
    vy = torch.stack([s[0][:, 1], s[1][:, 0], s[2][:, 1], s[3][:, 0]], dim=1).abs().max(dim=1)[0]
    pol = torch.where(s[0] > 0, torch.ones_like(s[0]), -torch.ones_like(s[0]))
    return torch.stack([vx, vy], dim=1) * pol


# ----------------------------------------------------------------------------- whole step
def forward_train(p, img, gt_bboxes, gt_labels, gt_masks, gt_offsets, cfg=CFG, forced=None,
                  record=None, stable_sort=False, aux=None, forced_proposals=None,
                  roi_align_fn=None, nms_fn=None):
    """TwoStageDetector.forward_train + LoftRoIHead.forward_train, detectors/two_stage.py:105-167,
    roi_heads/loft_roi_head.py:44-194.  Returns the reference's loss dict.

    forced / record: lists of sampled index tensors consumed / produced in call order (RPN pos, RPN
    neg per image, then RCNN pos, neg per image) so a second implementation can be teacher-forced.
    """
    n_img = img.size(0)
    img_shape = tuple(img.shape[-2:])
    c = resnet50(img, p)
    feats = fpn(c, p)
    cls_scores, bbox_preds = rpn_forward(feats, p)
    losses = {}
    l_cls, l_box = rpn_loss(cls_scores, bbox_preds, gt_bboxes, cfg, forced, record)
    losses['loss_rpn_cls'], losses['loss_rpn_bbox'] = l_cls, l_box
    if forced_proposals is not None:
        proposals = forced_proposals
    else:
        proposals = rpn_get_bboxes(cls_scores, bbox_preds, [img_shape] * n_img, cfg, stable_sort,
                                   nms_fn)
    samples = []
    for i in range(n_img):
        a = cfg['rcnn_assign']
        props = proposals[i][:, :4]
        gt_inds, _, labels = max_iou_assign(props, gt_bboxes[i], a['pos'], a['neg'], a['min_pos'],
                                            gt_labels[i])
        s = cfg['rcnn_sampler']
        samples.append(random_sample(gt_inds, props, gt_bboxes[i], gt_labels[i], labels, s['num'],
                                     s['pos_fraction'], s['add_gt'], forced, record))
    # bbox branch, standard_roi_head.py:135-161
    rois = bbox2roi([sr['bboxes'] for sr in samples])
    bbox_feats = roi_extract(feats[:4], rois, 7, cfg, roi_align_fn)
    cls_score, bbox_pred = bbox_head_forward(bbox_feats, p)
    losses.update(bbox_head_loss(cls_score, bbox_pred, samples, cfg))
    # mask branch, loft_roi_head.py:162-194
    pos_rois = bbox2roi([sr['pos_bboxes'] for sr in samples])
    mask_feats = roi_extract(feats[:4], pos_rois, 14, cfg, roi_align_fn)
    mask_pred = mask_head_forward(mask_feats, p) if pos_rois.size(0) > 0 else \
        mask_feats.new_zeros((0, 1, 28, 28))
    m_targets = mask_target(samples, gt_masks, cfg, roi_align_fn)
    pos_labels = torch.cat([sr['pos_gt_labels'] for sr in samples])
    if mask_pred.size(0) == 0:
        losses['loss_mask'] = mask_pred.sum() * 0
    else:
        n = mask_pred.size(0)
        sl = mask_pred[torch.arange(n), pos_labels].squeeze(1)       # cross_entropy_loss.py:94-125
        losses['loss_mask'] = F.binary_cross_entropy_with_logits(sl, m_targets,
                                                                 reduction='mean')[None]
    # offset branch, loft_roi_head.py:127-160
    off_feats = roi_extract(feats[:4], pos_rois, 7, cfg, roi_align_fn)
    off_pred = offset_head_forward(off_feats, p, cfg)
    o_targets = offset_targets(samples, gt_offsets, cfg)
    if off_pred.size(0) == 0:
        losses['loss_offset'] = off_pred.sum() * 0
    else:
        losses['loss_offset'] = cfg['loss_offset_weight'] * smooth_l1_mean(off_pred, o_targets)
    if aux is not None:
        aux.update(feats=feats, cls_scores=cls_scores, bbox_preds=bbox_preds, proposals=proposals,
                   samples=samples, rois=rois, bbox_feats=bbox_feats, cls_score=cls_score,
                   bbox_pred=bbox_pred, mask_pred=mask_pred, mask_targets=m_targets,
                   offset_pred=off_pred, offset_targets=o_targets, pos_rois=pos_rois)
    return losses


def parse_losses(losses):
    """BaseDetector._parse_losses (single process), detectors/base.py:175-208."""
    log_vars = {}
    for k, v in losses.items():
        if isinstance(v, torch.Tensor):
            log_vars[k] = v.mean()
        else:
            log_vars[k] = sum(x.mean() for x in v)
    loss = sum(v for k, v in log_vars.items() if 'loss' in k)
    log_vars['loss'] = loss
    return loss, log_vars


def sgd_step(params, grads, momentum_buf, lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0):
    """mmcv OptimizerHook(grad_clip=dict(max_norm=35, norm_type=2)) + torch.optim.SGD
    (configs/_base_/schedules/schedule_2x_bonai.py:2-3).  In-place on `params`."""
    keys = [k for k in params if grads.get(k) is not None]
    total = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in keys)).float()
    coef = max_norm / (total + 1e-6)
    for k in keys:
        g = grads[k] * coef if coef < 1 else grads[k]
        g = g + weight_decay * params[k]
        buf = momentum_buf.get(k)
        buf = g.clone() if buf is None else buf.mul_(momentum).add_(g)
        momentum_buf[k] = buf
        params[k].sub_(lr * buf)
    return total


# ----------------------------------------------------------------------------- synthetic data
def trainable_keys(p):
    """Parameters that receive gradients: everything except stem + layer1 (frozen_stages=1,
    resnet.py:573-589) and BN running statistics."""
    out = []
    for k, v in p.items():
        if 'running_' in k or 'num_batches' in k:
            continue
        if k.startswith('backbone.conv1') or k.startswith('backbone.bn1') or \
                k.startswith('backbone.layer1.'):
            continue
        out.append(k)
    return out


def make_inputs(seed, n_img, size, num_gt, device='cpu'):
    """Synthetic tile batch of SURVEY.md section 8(d): randn image, G boxes with log-uniform sides,
    inscribed-ellipse roof masks, U(-40,40) offsets.  `size` is an int (square) or (H, W);
    `num_gt` an int or one count per image (ragged batches)."""
    H, W = (size, size) if isinstance(size, int) else size
    counts = [num_gt] * n_img if isinstance(num_gt, int) else list(num_gt)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(n_img, 3, H, W, generator=g)
    gt_bboxes, gt_labels, gt_masks, gt_offsets = [], [], [], []
    lim = torch.tensor([W, H, W, H], dtype=torch.float32)
    smin, smax = 16.0, min(160.0, min(H, W) / 2.0)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                            torch.arange(W, dtype=torch.float32), indexing='ij')
    for i in range(n_img):
        n = counts[i]
        cxy = torch.rand(n, 2, generator=g) * lim[:2]
        wh = torch.exp(torch.rand(n, 2, generator=g) * math.log(smax / smin)) * smin
        b = torch.min(torch.cat([cxy - wh / 2, cxy + wh / 2], dim=1).clamp(min=0), lim)
        small = (b[:, 2:] - b[:, :2]) < 2
        b[:, 2:] = torch.where(small, torch.min(b[:, :2] + 2, lim[2:]), b[:, 2:])
        b[:, :2] = torch.min(b[:, :2], b[:, 2:] - 2)
        gt_bboxes.append(b)
        gt_labels.append(torch.zeros(n, dtype=torch.long))
        cx, cy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
        rx, ry = (b[:, 2] - b[:, 0]) / 2, (b[:, 3] - b[:, 1]) / 2
        m = (((xx[None] + 0.5 - cx[:, None, None]) / rx[:, None, None]) ** 2 +
             ((yy[None] + 0.5 - cy[:, None, None]) / ry[:, None, None]) ** 2) <= 1.0
        gt_masks.append(m.to(torch.uint8))
        gt_offsets.append(torch.rand(n, 2, generator=g) * 80 - 40)
    return img, gt_bboxes, gt_labels, gt_masks, gt_offsets


def randomize_bn(p, seed=0):
    """Well-conditioned BN re-randomisation for parity runs (default init has
    zero_init_residual => all residual branches dead)."""
    g = torch.Generator().manual_seed(seed)
    for k in list(p.keys()):
        if not (k.startswith('backbone') and ('bn' in k or 'downsample.1' in k)):
            continue
        if k.endswith('.weight'):
            p[k] = torch.rand(p[k].shape, generator=g) * 0.4 + 0.3
        elif k.endswith('.bias'):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
        elif k.endswith('running_mean'):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            p[k] = torch.rand(p[k].shape, generator=g) + 0.5
    return p


def init_params(seed=0):
    """Random-init LOFT R50-FPN state dict with the reference's names/shapes (App. D) and init
    distributions (kaiming / xavier / normal as in the reference's init_weights)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def kaiming(shape, mode='fan_out'):
        fan = shape[0] * shape[2] * shape[3] if mode == 'fan_out' else shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan)

    def bn(prefix, c, gamma=1.0):
        p[prefix + '.weight'] = torch.full((c,), gamma)
        p[prefix + '.bias'] = torch.zeros(c)
        p[prefix + '.running_mean'] = torch.zeros(c)
        p[prefix + '.running_var'] = torch.ones(c)
        p[prefix + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    p['backbone.conv1.weight'] = kaiming((64, 3, 7, 7))
    bn('backbone.bn1', 64)
    inpl = 64
    for li, (nb, planes) in enumerate(zip([3, 4, 6, 3], [64, 128, 256, 512])):
        for b in range(nb):
            pre = f'backbone.layer{li + 1}.{b}'
            p[pre + '.conv1.weight'] = kaiming((planes, inpl, 1, 1))
            bn(pre + '.bn1', planes)
            p[pre + '.conv2.weight'] = kaiming((planes, planes, 3, 3))
            bn(pre + '.bn2', planes)
            p[pre + '.conv3.weight'] = kaiming((planes * 4, planes, 1, 1))
            bn(pre + '.bn3', planes * 4, gamma=0.0)               # zero_init_residual
            if b == 0:
                p[pre + '.downsample.0.weight'] = kaiming((planes * 4, inpl, 1, 1))
                bn(pre + '.downsample.1', planes * 4)
            inpl = planes * 4

    def xavier(shape):
        fi = shape[1] * shape[2] * shape[3]
        fo = shape[0] * shape[2] * shape[3]
        b = math.sqrt(6.0 / (fi + fo))
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    for i, c in enumerate([256, 512, 1024, 2048]):
        p[f'neck.lateral_convs.{i}.conv.weight'] = xavier((256, c, 1, 1))
        p[f'neck.lateral_convs.{i}.conv.bias'] = torch.zeros(256)
        p[f'neck.fpn_convs.{i}.conv.weight'] = xavier((256, 256, 3, 3))
        p[f'neck.fpn_convs.{i}.conv.bias'] = torch.zeros(256)
    for name, co, k in [('rpn_conv', 256, 3), ('rpn_cls', 3, 1), ('rpn_reg', 12, 1)]:
        p[f'rpn_head.{name}.weight'] = torch.randn((co, 256, k, k), generator=g) * 0.01
        p[f'rpn_head.{name}.bias'] = torch.zeros(co)

    def xavier_lin(o, i):
        b = math.sqrt(6.0 / (i + o))
        return (torch.rand((o, i), generator=g) * 2 - 1) * b

    bh = 'roi_head.bbox_head'
    p[bh + '.fc_cls.weight'] = torch.randn((2, 1024), generator=g) * 0.01
    p[bh + '.fc_cls.bias'] = torch.zeros(2)
    p[bh + '.fc_reg.weight'] = torch.randn((4, 1024), generator=g) * 0.001
    p[bh + '.fc_reg.bias'] = torch.zeros(4)
    for i, fi in enumerate([12544, 1024]):
        p[f'{bh}.shared_fcs.{i}.weight'] = xavier_lin(1024, fi)
        p[f'{bh}.shared_fcs.{i}.bias'] = torch.zeros(1024)
    mh = 'roi_head.mask_head'
    for i in range(4):
        p[f'{mh}.convs.{i}.conv.weight'] = kaiming((256, 256, 3, 3))
        p[f'{mh}.convs.{i}.conv.bias'] = torch.zeros(256)
    p[mh + '.upsample.weight'] = torch.randn((256, 256, 2, 2), generator=g) * math.sqrt(2.0 / 1024)
    p[mh + '.upsample.bias'] = torch.zeros(256)
    p[mh + '.conv_logits.weight'] = kaiming((1, 256, 1, 1))
    p[mh + '.conv_logits.bias'] = torch.zeros(1)
    oh = 'roi_head.offset_head'
    for br in range(4):
        for c in range(10):
            p[f'{oh}.expand_convs.{br}.{c}.weight'] = kaiming((256, 256, 3, 3))
            p[f'{oh}.expand_convs.{br}.{c}.bias'] = torch.zeros(256)
    for i, fi in enumerate([12544, 1024]):
        b = math.sqrt(3.0 / fi)                                  # kaiming_uniform(a=1, fan_in)
        p[f'{oh}.fcs.{i}.weight'] = (torch.rand((1024, fi), generator=g) * 2 - 1) * b
        p[f'{oh}.fcs.{i}.bias'] = torch.zeros(1024)
    p[oh + '.fc_offset.weight'] = torch.randn((2, 1024), generator=g) * 0.01
    p[oh + '.fc_offset.bias'] = torch.zeros(2)
    return p


# ----------------------------------------------------------------------------- inference
TEST_CFG = dict(score_thr=0.05, nms_iou=0.5, max_per_img=2000, mask_thr_binary=0.5)   # :128-140


def multiclass_nms_soft(bboxes, scores, score_thr, iou_thr, max_num):
    """multiclass_nms with nms=dict(type='soft_nms'), 1 foreground class
    (core/post_processing/bbox_nms.py:5-69)."""
    sc = scores[:, :-1]
    valid = sc > score_thr
    b = bboxes.view(scores.size(0), -1, 4)[valid]
    s = sc[valid]
    labels = valid.nonzero(as_tuple=False)[:, 1]
    if b.numel() == 0:
        return bboxes.new_zeros((0, 5)), labels.new_zeros((0,))
    max_coordinate = b.max()
    offs = labels.to(b) * (max_coordinate + 1)
    dets, keep = ops_cpu.soft_nms_linear(b + offs[:, None], s, iou_thr, 1e-3)
    dets = torch.cat([b[keep], dets[:, -1:]], -1)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, labels[keep]


def paste_masks(mask_pred, det_bboxes, img_h, img_w, thr=0.5):
    """FCNMaskHead.get_seg_masks + _do_paste_mask on CPU (fcn_mask_head.py:151-308), rescale=False,
    scale_factor=1: returns a bool tensor [k, H, W]."""
    n = mask_pred.shape[0]
    out = torch.zeros(n, img_h, img_w, dtype=torch.bool)
    m = mask_pred.sigmoid()
    for i in range(n):                                   # CPU path: one chunk per detection
        boxes = det_bboxes[i:i + 1, :4]
        x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(torch.int32)
        x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(torch.int32)
        y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(torch.int32)
        x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
        img_y = torch.arange(y0_int, y1_int, dtype=torch.float32) + 0.5
        img_x = torch.arange(x0_int, x1_int, dtype=torch.float32) + 0.5
        img_y = (img_y - y0) / (y1 - y0) * 2 - 1
        img_x = (img_x - x0) / (x1 - x0) * 2 - 1
        img_x[torch.isinf(img_x)] = 0
        img_y[torch.isinf(img_y)] = 0
        gx = img_x[:, None, :].expand(1, img_y.size(1), img_x.size(1))
        gy = img_y[:, :, None].expand(1, img_y.size(1), img_x.size(1))
        grid = torch.stack([gx, gy], dim=3)
        pm = F.grid_sample(m[i:i + 1].float(), grid, align_corners=False)
        out[i, y0_int:y1_int, x0_int:x1_int] = pm[0, 0] >= thr
    return out


@torch.no_grad()
def simple_test(p, img, cfg=CFG, tcfg=TEST_CFG, proposals=None, aux=None):
    """TwoStageDetector.simple_test -> LoftRoIHead.simple_test (two_stage.py:187-199,
    loft_roi_head.py:196-227, test_mixins.py:53-72,152-177,211-241).  One image, rescale=False.
    Returns (dets[k,5], masks bool[k,H,W], offsets[k,2])."""
    assert img.size(0) == 1
    img_shape = tuple(img.shape[-2:])
    feats = fpn(resnet50(img, p), p)
    if proposals is None:
        cls_scores, bbox_preds = rpn_forward(feats, p)
        proposals = rpn_get_bboxes(cls_scores, bbox_preds, [img_shape], cfg)
    rois = bbox2roi([q[:, :4] for q in proposals])
    cls_score, bbox_pred = bbox_head_forward(roi_extract(feats[:4], rois, 7, cfg), p)
    scores = F.softmax(cls_score, dim=1)
    bboxes = delta2bbox(rois[:, 1:], bbox_pred, cfg['rcnn_stds'], max_shape=img_shape)
    dets, labels = multiclass_nms_soft(bboxes, scores, tcfg['score_thr'], tcfg['nms_iou'],
                                       tcfg['max_per_img'])
    if dets.shape[0] == 0:
        return dets, torch.zeros((0,) + img_shape, dtype=torch.bool), dets.new_zeros((0, 2))
    det_rois = bbox2roi([dets[:, :4]])
    mask_pred = mask_head_forward(roi_extract(feats[:4], det_rois, 14, cfg), p)
    masks = paste_masks(mask_pred[:, 0:1], dets, img_shape[0], img_shape[1], tcfg['mask_thr_binary'])
    off_pred = offset_head_forward(roi_extract(feats[:4], det_rois, 7, cfg), p, cfg)
    offsets = delta2offset(dets, offset_fusion_max(off_pred), cfg['offset_stds'],
                           max_shape=[1024, 1024])       # get_offsets default img_shape (App. C.7)
    if aux is not None:
        aux.update(proposals=proposals, mask_pred=mask_pred, offset_pred=off_pred)
    return dets, masks, offsets
