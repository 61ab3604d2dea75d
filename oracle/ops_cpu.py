"""ORACLE (test infrastructure, not product code) -- CPU restatement of the native ops the
reference's LOFT path takes from the un-vendored dependency mmcv-full==1.0.5
(pin: mmdet/__init__.py:18-26): RoIAlign, nms, batched_nms, soft_nms, sigmoid_focal_loss.

The algorithms are restated from their published definition (Detectron RoIAlign "aligned"
variant, greedy NMS, Bodla et al. soft-NMS, Lin et al. focal loss) as recorded in SURVEY.md
Appendix A, anchored on the reference's call sites:
  RoIAlign      roi_heads/roi_extractors/base_roi_extractor.py:49-55, single_level_roi_extractor.py:76
  roi_align     core/mask/structures.py:286-287
  batched_nms   dense_heads/rpn_head.py:166-167, core/post_processing/bbox_nms.py:63
  focal loss    models/losses/focal_loss.py:10-41 (py_sigmoid_focal_loss, pure-PyTorch twin)
Pinned in tests/test_oracle_ops.py against torchvision 0.26 (the implementation the reference
itself switches to with use_torchvision=True, mmdet/apis/inference.py:102-109) and against the
reference's own py_sigmoid_focal_loss via tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np
import torch


# --------------------------------------------------------------------------- RoIAlign
def _axis_samples(start, bin_size, grid, out, size):
    """Per-axis sample coordinates and bilinear taps for a group of RoIs sharing `grid`.

    start, bin_size: [k] tensors.  Returns (low, high, w_low, w_high, valid) each [k, out*grid].
    Follows the mmcv/Detectron bilinear_interpolate edge rules:
    out of range iff y < -1 or y > size; clamp y >= 0; if y_low >= size-1: low = high = size-1.
    """
    p = torch.arange(out, dtype=start.dtype).view(1, out, 1)
    i = torch.arange(grid, dtype=start.dtype).view(1, 1, grid)
    y = start.view(-1, 1, 1) + p * bin_size.view(-1, 1, 1) + \
        (i + 0.5) * bin_size.view(-1, 1, 1) / grid
    y = y.reshape(start.numel(), out * grid)
    valid = ~((y < -1.0) | (y > size))
    y = y.clamp(min=0)
    low = y.floor().long()
    top = low >= size - 1
    low = torch.where(top, torch.full_like(low, size - 1), low)
    high = torch.where(top, low, low + 1)
    y = torch.where(top, low.to(y.dtype), y)
    ly = y - low.to(y.dtype)
    hy = 1.0 - ly
    return low, high, hy, ly, valid


def roi_align(feat, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    """feat [N,C,H,W] (differentiable), rois [K,5] (batch_idx, x1,y1,x2,y2) -> [K,C,oh,ow].

    avg pooling; sampling_ratio 0 => adaptive ceil(roi/out) samples per bin.
    RoIs are processed in groups of equal (batch, grid_h, grid_w) so every group is a dense
    gather; gradients flow to `feat` through index ops.
    """
    oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
    N, C, H, W = feat.shape
    K = rois.shape[0]
    out = feat.new_zeros((K, C, oh, ow))
    if K == 0:
        return out
    rois = rois.detach()
    off = 0.5 if aligned else 0.0
    bidx = rois[:, 0].long()
    x1 = rois[:, 1] * spatial_scale - off
    y1 = rois[:, 2] * spatial_scale - off
    x2 = rois[:, 3] * spatial_scale - off
    y2 = rois[:, 4] * spatial_scale - off
    rw, rh = x2 - x1, y2 - y1
    if not aligned:
        rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
    bin_h, bin_w = rh / oh, rw / ow
    if sampling_ratio > 0:
        gh = torch.full((K,), sampling_ratio, dtype=torch.long)
        gw = gh.clone()
    else:
        gh = torch.ceil(rh / oh).long()
        gw = torch.ceil(rw / ow).long()
    count = (gh * gw).clamp(min=1).to(feat.dtype)
    key = (bidx * 100000 + gh.clamp(min=0)) * 100000 + gw.clamp(min=0)
    pieces = []
    order = []
    for kval in torch.unique(key):
        sel = torch.nonzero(key == kval, as_tuple=False).squeeze(1)
        b = int(bidx[sel[0]])
        g_h, g_w = int(gh[sel[0]]), int(gw[sel[0]])
        order.append(sel)
        if g_h <= 0 or g_w <= 0:
            pieces.append(feat.new_zeros((sel.numel(), C, oh, ow)) + feat[b].sum() * 0)
            continue
        yl, yh, wyl, wyh, vy = _axis_samples(y1[sel], bin_h[sel], g_h, oh, H)
        xl, xh, wxl, wxh, vx = _axis_samples(x1[sel], bin_w[sel], g_w, ow, W)
        fm = feat[b]  # [C,H,W]
        k = sel.numel()
        Y, X = oh * g_h, ow * g_w

        def tap(yi, xi):
            idx = (yi.view(k, Y, 1) * W + xi.view(k, 1, X)).reshape(-1)
            return fm.reshape(C, H * W).index_select(1, idx).view(C, k, Y, X)

        wy_l = (wyl * vy).view(1, k, Y, 1)
        wy_h = (wyh * vy).view(1, k, Y, 1)
        wx_l = (wxl * vx).view(1, k, 1, X)
        wx_h = (wxh * vx).view(1, k, 1, X)
        val = tap(yl, xl) * (wy_l * wx_l) + tap(yl, xh) * (wy_l * wx_h) + \
            tap(yh, xl) * (wy_h * wx_l) + tap(yh, xh) * (wy_h * wx_h)
        val = val.view(C, k, oh, g_h, ow, g_w).sum(dim=(3, 5)).permute(1, 0, 2, 3)
        pieces.append(val / count[sel].view(k, 1, 1, 1))
    order = torch.cat(order)
    vals = torch.cat(pieces, 0)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel())
    return vals.index_select(0, inv)


# --------------------------------------------------------------------------- NMS
def nms(boxes, scores, iou_threshold):
    """Greedy NMS: stable sort by score desc; suppress j iff inter/(a_i+a_j-inter) > thr (fp32,
    no +1).  Returns (dets[k,5], keep[k] int64) like mmcv.ops.nms."""
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu().numpy().astype(np.float32)
    n = b.shape[0]
    if n == 0:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long)
    order = np.argsort(-s, kind='stable')
    bs = b[order]
    x1, y1, x2, y2 = bs[:, 0], bs[:, 1], bs[:, 2], bs[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        iou = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= iou > thr
    keep = torch.from_numpy(np.asarray(keep, dtype=np.int64))
    dets = torch.cat([boxes[keep], scores[keep].reshape(-1, 1)], dim=1)
    return dets, keep


def batched_nms(boxes, scores, idxs, iou_threshold):
    """mmcv 1.0.5 batched_nms: offset boxes by idx*(max_coord+1) in fp32, then plain nms."""
    if boxes.numel() == 0:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + 1)
    boxes_for_nms = boxes + offsets[:, None]
    _, keep = nms(boxes_for_nms, scores, iou_threshold)
    return torch.cat([boxes[keep], scores[keep][:, None]], -1), keep


def soft_nms_linear(boxes, scores, iou_threshold=0.5, min_score=1e-3):
    """Linear soft-NMS (test-time, core/post_processing/bbox_nms.py:63 with
    bonai_loft_foa_r50_fpn_basic.py:138)."""
    b = boxes.detach().cpu().numpy().astype(np.float32).copy()
    s = scores.detach().cpu().numpy().astype(np.float32).copy()
    idx = np.arange(b.shape[0])
    n = b.shape[0]
    out_b, out_s, out_i = [], [], []
    while n > 0:
        m = int(np.argmax(s[:n]))
        for arr in (b, s, idx):
            tmp = arr[0].copy()
            arr[0] = arr[m]
            arr[m] = tmp
        out_b.append(b[0].copy())
        out_s.append(s[0])
        out_i.append(idx[0])
        bx = b[1:n]
        w = np.maximum(np.minimum(b[0, 2], bx[:, 2]) - np.maximum(b[0, 0], bx[:, 0]), 0)
        h = np.maximum(np.minimum(b[0, 3], bx[:, 3]) - np.maximum(b[0, 1], bx[:, 1]), 0)
        inter = w * h
        a0 = (b[0, 2] - b[0, 0]) * (b[0, 3] - b[0, 1])
        a = (bx[:, 2] - bx[:, 0]) * (bx[:, 3] - bx[:, 1])
        ovr = inter / (a0 + a - inter)
        s[1:n] = s[1:n] * np.where(ovr > iou_threshold, 1 - ovr, 1.0).astype(np.float32)
        km = s[1:n] >= min_score
        k = int(km.sum())
        b[:k], s[:k], idx[:k] = b[1:n][km], s[1:n][km], idx[1:n][km]
        n = k
    if not out_b:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long)
    dets = np.concatenate([np.stack(out_b), np.asarray(out_s, np.float32)[:, None]], 1)
    return torch.from_numpy(dets), torch.from_numpy(np.asarray(out_i, np.int64))


# --------------------------------------------------------------------------- focal loss
def sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25):
    """Element-wise sigmoid focal loss, target[N] int64 in [0,C] with C = background
    (mmcv CUDA op semantics; == py_sigmoid_focal_loss of focal_loss.py:10-41 on one-hot targets).
    Returns the un-reduced [N,C] loss."""
    C = pred.size(1)
    t = torch.nn.functional.one_hot(target.clamp(min=0), C + 1)[:, :C].to(pred.dtype)
    p = pred.sigmoid()
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return torch.nn.functional.binary_cross_entropy_with_logits(pred, t, reduction='none') * fw
