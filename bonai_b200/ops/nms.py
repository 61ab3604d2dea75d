"""NMS bindings with the mmcv.ops signatures used by the reference (`nms`, `batched_nms`,
dense_heads/rpn_head.py:166-167, core/post_processing/bbox_nms.py:63)."""
import ctypes

import torch

from .. import _lib as L

i32 = ctypes.c_int


def nms_sorted(boxes, idxs, iou_threshold, max_keep=-1):
    """boxes [B,n,4] sorted by score (desc) per image, idxs [B,n] int64 or None.
    Returns keep [B,n] (positions in the sorted order, first num_keep valid) and num_keep [B]."""
    B, n, _ = boxes.shape
    keep = torch.empty((B, n), device=boxes.device, dtype=torch.long)
    num = torch.zeros((B,), device=boxes.device, dtype=torch.int32)
    if n == 0 or B == 0:
        return keep, num
    ws_bytes = B * L.lib().loft_nms_workspace(i32(n))
    ws = torch.empty((ws_bytes,), device=boxes.device, dtype=torch.uint8)
    L.call('nms_sorted', L.ptr(boxes.contiguous()), L.ptr(idxs.contiguous()) if idxs is not None else None,
           i32(B), i32(n), L.f32(iou_threshold), i32(max_keep), L.ptr(keep), L.ptr(num), L.ptr(ws),
           ctypes.c_size_t(ws_bytes), L.stream())
    return keep, num


def nms_segmented(boxes, seg_sizes, order, iou_threshold, max_keep=-1):
    """Per-level form of batched NMS (the RPN's): boxes [B,n,4] segment-major, every segment sorted
    by score; seg_sizes the segment lengths (segment index = the batched_nms idx); order [B,n]
    int64 = segment-major indices in global score order.  Same (keep, num_keep) as
    `nms_sorted(boxes.gather(order), idxs.gather(order), ...)`."""
    B, n, _ = boxes.shape
    keep = torch.empty((B, n), device=boxes.device, dtype=torch.long)
    num = torch.zeros((B,), device=boxes.device, dtype=torch.int32)
    if n == 0 or B == 0:
        return keep, num
    nseg = len(seg_sizes)
    off = (ctypes.c_int * (nseg + 1))()
    for i, k in enumerate(seg_sizes):
        off[i + 1] = off[i] + int(k)
    assert off[nseg] == n
    fn = L.lib().loft_nms_segmented_workspace
    fn.restype = ctypes.c_size_t
    ws_bytes = B * fn(off, i32(nseg))
    ws = torch.empty((ws_bytes,), device=boxes.device, dtype=torch.uint8)
    L.call('nms_segmented', L.ptr(boxes.contiguous()), off, i32(nseg), L.ptr(order.contiguous()),
           i32(B), i32(n), L.f32(iou_threshold), i32(max_keep), L.ptr(keep), L.ptr(num), L.ptr(ws),
           ctypes.c_size_t(ws_bytes), L.stream())
    return keep, num


def nms(boxes, scores, iou_threshold, offset=0):
    """mmcv.ops.nms: returns (dets[k,5], inds[k]) with inds in score-descending order
    (ties keep input order)."""
    assert offset == 0
    if boxes.shape[0] == 0:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long, device=boxes.device)
    order = torch.sort(scores, descending=True, stable=True)[1]
    keep, num = nms_sorted(boxes[order].float()[None], None, float(iou_threshold))
    k = int(num[0])
    inds = order[keep[0, :k]]
    return torch.cat([boxes[inds], scores[inds, None]], dim=1), inds


def soft_nms(boxes, scores, iou_threshold=0.3, sigma=0.5, min_score=1e-3, method='linear',
             offset=0, idxs=None, max_keep=-1):
    """mmcv.ops.soft_nms (linear): returns (dets[k,5] with decayed scores, inds[k]) in selection
    order.  `idxs` (optional) applies batched_nms' per-class coordinate offset inside the kernel."""
    if method != 'linear' or offset != 0:
        raise NotImplementedError('LOFT test_cfg uses linear soft-NMS with offset 0')
    n = boxes.shape[0]
    dev = boxes.device
    dets = torch.empty((n, 5), device=dev, dtype=torch.float32)
    keep = torch.empty((n,), device=dev, dtype=torch.long)
    num = torch.zeros((1,), device=dev, dtype=torch.int32)
    if n == 0:
        return dets, keep
    L.call('soft_nms_linear', L.ptr(boxes.contiguous().float()), L.ptr(scores.contiguous().float()),
           L.ptr(idxs.contiguous().long()) if idxs is not None else None, i32(n),
           L.f32(iou_threshold), L.f32(min_score), i32(max_keep), L.ptr(dets), L.ptr(keep),
           L.ptr(num), L.stream())
    k = int(num[0])
    return dets[:k], keep[:k]


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.batched_nms (v1.0.5): boxes of different `idxs` never suppress each other (fp32
    coordinate-offset trick, applied inside the kernel)."""
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop('class_agnostic', class_agnostic)
    typ = cfg.pop('type', 'nms')
    if typ == 'soft_nms':
        dets, keep = soft_nms(boxes, scores, idxs=None if class_agnostic else idxs, **cfg)
        return torch.cat([boxes[keep], dets[:, 4:5]], dim=-1), keep
    if typ != 'nms':
        raise NotImplementedError(f'batched_nms type {typ!r}')
    thr = float(cfg.pop('iou_threshold'))
    if boxes.shape[0] == 0:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long, device=boxes.device)
    order = torch.sort(scores, descending=True, stable=True)[1]
    ids = None if class_agnostic else idxs[order].long()[None]
    keep, num = nms_sorted(boxes[order].float()[None], ids, thr)
    k = int(num[0])
    inds = order[keep[0, :k]]
    return torch.cat([boxes[inds], scores[inds, None]], dim=-1), inds
