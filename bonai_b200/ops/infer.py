"""Test-time post-processing bindings (csrc/infer.cu): mask paste and FOA offset fusion + decode."""
import ctypes

import torch

from .. import _lib as L

i32 = ctypes.c_int


def paste_masks(mask_logits, boxes, img_h, img_w, threshold=0.5):
    """FCNMaskHead.get_seg_masks + _do_paste_mask (mmdet/models/roi_heads/mask_heads/
    fcn_mask_head.py:151-308) for one class channel.  mask_logits: [N, M, M] view (any strides over
    N and over the flattened M*M pixels, e.g. channel 0 of the fused NHWC head output) of the RAW
    mask logits; boxes [N, >=4].  Returns a bool (threshold >= 0) or uint8 tensor [N, H, W]."""
    N, M = mask_logits.shape[0], mask_logits.shape[-1]
    out = torch.zeros((N, int(img_h), int(img_w)), device=boxes.device, dtype=torch.uint8)
    if N == 0:
        return out.bool() if threshold >= 0 else out
    assert mask_logits.dim() == 3 and mask_logits.shape[1] == M
    sn, sy, sx = mask_logits.stride()
    assert sy == M * sx, 'mask pixels must be uniformly strided'
    boxes = boxes.float()
    if boxes.stride(1) != 1:
        boxes = boxes.contiguous()
    L.call('paste_masks', L.ptr(mask_logits), L.ll(sn), i32(sx), L.ptr(boxes), i32(boxes.stride(0)),
           i32(N), i32(M), i32(int(img_h)), i32(int(img_w)), L.f32(threshold), L.ptr(out),
           L.stream())
    return out.view(torch.bool) if threshold >= 0 else out


def offset_fusion_decode(offset_pred, boxes, stds=(0.5, 0.5), max_shape=None):
    """offset_fusion('max') + DeltaXYOffsetCoder.decode in one launch (offset_head_expand_feature.py:
    346-448, delta_xy_offset_coder.py:67-88).  offset_pred [4n, >=2] branch-major; boxes [n, >=4]."""
    n = boxes.shape[0]
    assert offset_pred.shape[0] == 4 * n and offset_pred.stride(1) == 1
    out = torch.empty((n, 2), device=boxes.device, dtype=torch.float32)
    if n == 0:
        return out
    boxes = boxes.float()
    if boxes.stride(1) != 1:
        boxes = boxes.contiguous()
    mx, my = (float(max_shape[1]), float(max_shape[0])) if max_shape is not None else (0.0, 0.0)
    L.call('offset_fusion_decode', L.ptr(offset_pred), i32(offset_pred.stride(0)), L.ll(n),
           L.ptr(boxes), i32(boxes.stride(0)), L.f32(stds[0]), L.f32(stds[1]), L.f32(mx), L.f32(my),
           L.ptr(out), L.stream())
    return out
