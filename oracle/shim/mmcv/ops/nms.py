import numpy as np
import torch
import torchvision.ops as tvo


def nms(boxes, scores, iou_threshold, offset=0):
    assert offset == 0
    is_np = isinstance(boxes, np.ndarray)
    if is_np:
        boxes, scores = torch.from_numpy(boxes), torch.from_numpy(scores)
    inds = tvo.nms(boxes.float(), scores.float(), float(iou_threshold))
    dets = torch.cat((boxes[inds], scores[inds].reshape(-1, 1)), dim=1)
    if is_np:
        return dets.numpy(), inds.numpy()
    return dets, inds


def soft_nms(boxes, scores, iou_threshold=0.3, sigma=0.5, min_score=1e-3, method='linear',
             offset=0):
    """Linear / gaussian / naive soft-NMS (CPU loop), mmcv v1.0.5 semantics."""
    b = boxes.detach().cpu().numpy().astype(np.float32).copy()
    s = scores.detach().cpu().numpy().astype(np.float32).copy()
    n = b.shape[0]
    idx = np.arange(n)
    keep_d, keep_i = [], []
    while n > 0:
        m = int(np.argmax(s[:n]))
        b[[0, m]] = b[[m, 0]]
        s[[0, m]] = s[[m, 0]]
        idx[[0, m]] = idx[[m, 0]]
        keep_d.append(np.concatenate([b[0], s[:1]]))
        keep_i.append(idx[0])
        bx = b[1:n]
        xx1 = np.maximum(b[0, 0], bx[:, 0]); yy1 = np.maximum(b[0, 1], bx[:, 1])
        xx2 = np.minimum(b[0, 2], bx[:, 2]); yy2 = np.minimum(b[0, 3], bx[:, 3])
        w = np.maximum(xx2 - xx1 + offset, 0); h = np.maximum(yy2 - yy1 + offset, 0)
        inter = w * h
        a0 = (b[0, 2] - b[0, 0] + offset) * (b[0, 3] - b[0, 1] + offset)
        a = (bx[:, 2] - bx[:, 0] + offset) * (bx[:, 3] - bx[:, 1] + offset)
        ovr = inter / (a0 + a - inter)
        if method == 'linear':
            wgt = np.where(ovr > iou_threshold, 1 - ovr, 1.0)
        elif method == 'gaussian':
            wgt = np.exp(-(ovr * ovr) / sigma)
        else:
            wgt = np.where(ovr > iou_threshold, 0.0, 1.0)
        s[1:n] = s[1:n] * wgt
        keepmask = s[1:n] >= min_score
        k = int(keepmask.sum())
        b[:k] = b[1:n][keepmask]; s[:k] = s[1:n][keepmask]; idx[:k] = idx[1:n][keepmask]
        n = k
    dets = torch.from_numpy(np.stack(keep_d)) if keep_d else boxes.new_zeros((0, 5))
    inds = torch.tensor(keep_i, dtype=torch.long)
    return dets.to(boxes.device), inds.to(boxes.device)


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    nms_cfg_ = dict(nms_cfg)
    class_agnostic = nms_cfg_.pop('class_agnostic', class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + 1)
        boxes_for_nms = boxes + offsets[:, None]
    nms_type = nms_cfg_.pop('type', 'nms')
    nms_op = {'nms': nms, 'soft_nms': soft_nms}[nms_type]
    dets, keep = nms_op(boxes_for_nms, scores, **nms_cfg_)
    boxes = boxes[keep]
    scores = dets[:, -1]
    return torch.cat([boxes, scores[:, None]], -1), keep
