"""RoIAlign / mask-target bindings (mirror of mmcv.ops.RoIAlign / roi_align at the reference call
sites roi_extractors/base_roi_extractor.py:49-55, single_level_roi_extractor.py:53-80 and
core/mask/structures.py:286-287)."""
import ctypes

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from .. import _lib as L
from .dense import nhwc, new_nhwc

i32 = ctypes.c_int


def _pyramid_args(feats_nhwc, scales):
    n = len(feats_nhwc)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in feats_nhwc])
    Hs = (ctypes.c_int * n)(*[f.shape[1] for f in feats_nhwc])
    Ws = (ctypes.c_int * n)(*[f.shape[2] for f in feats_nhwc])
    sc = (ctypes.c_float * n)(*[float(s) for s in scales])
    return ptrs, Hs, Ws, sc


class _MultiLevelRoIAlign(Function):
    """All FPN levels in one launch; the level of each RoI is computed in-kernel."""

    @staticmethod
    def forward(ctx, rois, out_size, scales, finest_scale, sinks, *feats):
        fn = [nhwc(f) for f in feats]
        ctx.sinks = sinks
        K = rois.shape[0]
        C = fn[0].shape[-1]
        rois = rois.contiguous().float()
        out = new_nhwc(K, C, out_size, out_size, rois.device)
        ptrs, Hs, Ws, sc = _pyramid_args(fn, scales)
        L.call('roi_align_fwd', ptrs, Hs, Ws, sc, i32(len(fn)), L.ptr(rois), L.ll(K), i32(out_size),
               i32(C), L.f32(finest_scale), L.ptr(out.permute(0, 2, 3, 1)), None, L.stream())
        ctx.save_for_backward(rois)
        ctx.meta = (out_size, tuple(scales), finest_scale, [tuple(f.shape) for f in fn])
        return out

    @staticmethod
    def backward(ctx, dout):
        (rois,) = ctx.saved_tensors
        out_size, scales, finest, shapes = ctx.meta
        K = rois.shape[0]
        dn = nhwc(dout)
        C = dn.shape[-1]
        if ctx.sinks is not None:
            # the producer of the pyramid (bonai_b200.trunk) owns zeroed accumulation buffers:
            # every RoIAlign of the step adds into them, nothing is allocated, zero-filled or
            # summed by autograd, and no gradient tensor is returned for the feature maps
            ptrs, Hs, Ws, sc = _pyramid_args(ctx.sinks, scales)
            L.call('roi_align_bwd', ptrs, Hs, Ws, sc, i32(len(shapes)), L.ptr(rois), L.ll(K),
                   i32(out_size), i32(C), L.f32(finest), L.ptr(dn), L.stream())
            return (None,) * (5 + len(shapes))
        grads = [torch.zeros(s, device=dout.device, dtype=torch.float32) for s in shapes]
        ptrs, Hs, Ws, sc = _pyramid_args(grads, scales)
        L.call('roi_align_bwd', ptrs, Hs, Ws, sc, i32(len(grads)), L.ptr(rois), L.ll(K),
               i32(out_size), i32(C), L.f32(finest), L.ptr(dn), L.stream())
        return (None, None, None, None, None) + tuple(g.permute(0, 3, 1, 2) for g in grads)


def multilevel_roi_align(feats, rois, out_size, strides, finest_scale=56):
    C = feats[0].shape[1]
    if rois.shape[0] == 0:
        return feats[0].new_zeros((0, C, out_size, out_size))
    sinks = [getattr(f, '_loft_grad_sink', None) for f in feats]
    if any(s is None for s in sinks) or not torch.is_grad_enabled():
        sinks = None
    return _MultiLevelRoIAlign.apply(rois, int(out_size), [1.0 / s for s in strides],
                                     float(finest_scale), sinks, *feats)


class _TakeRows(Function):
    """feats [K,C,S,S] -> (feats, feats[rows]): a second consumer reads a subset of the rows (the
    FOA head reads the positives' rows of the bbox head's RoI features -- both extractors are the
    same 7x7 RoIAlign over the same pyramid, loft_roi_head.py:118-120).  The backward adds the
    subset's gradient into the same rows of the full gradient IN PLACE: no zero-filled [K,C,S,S]
    temporary, no full-size add, one RoIAlign backward for both heads."""

    @staticmethod
    def forward(ctx, feats, rows):
        ctx.set_materialize_grads(False)
        fn = nhwc(feats)
        K, S, S2, C = fn.shape
        P = int(rows.numel())
        out = new_nhwc(P, C, S, S2, feats.device)
        rows = rows.contiguous().long()
        L.call('gather_rows', L.ptr(fn), L.ptr(rows), L.ptr(out.permute(0, 2, 3, 1)), L.ll(P),
               L.ll(S * S2 * C), L.stream())
        ctx.save_for_backward(rows)
        ctx.full_shape = tuple(feats.shape)
        return feats, out

    @staticmethod
    def backward(ctx, g_full, g_rows):
        (rows,) = ctx.saved_tensors
        if g_rows is None:
            return g_full, None
        K, C, S, S2 = ctx.full_shape
        if g_full is None:
            g_full = new_nhwc(K, C, S, S2, g_rows.device).zero_()
        gf = g_full.permute(0, 2, 3, 1)
        if not gf.is_contiguous():
            gf = gf.contiguous()
            g_full = gf.permute(0, 3, 1, 2)
        L.call('scatter_add_rows', L.ptr(nhwc(g_rows)), L.ptr(rows), L.ptr(gf), L.ll(rows.numel()),
               L.ll(S * S2 * C), L.stream())
        return g_full, None


class _TakeRowsRot(Function):
    """feats [K,C,S,S] -> (feats, y [nb*P,C,S,S]) with y[b*P + p] = rot90(feats[rows[p]], ks[b]):
    `_TakeRows` fused with the FOA branch rotations and their concatenation (one launch instead of
    gather + one rot90 per branch + cat); the backward sums the branches' un-rotated gradients
    into the rows of the full gradient in place."""

    @staticmethod
    def forward(ctx, feats, rows, ks):
        ctx.set_materialize_grads(False)
        fn = nhwc(feats)
        K, S, S2, C = fn.shape
        assert S == S2
        P = int(rows.numel())
        nb = len(ks)
        out = new_nhwc(nb * P, C, S, S, feats.device)
        rows = rows.contiguous().long()
        karr = (ctypes.c_int * nb)(*[int(k) for k in ks])
        L.call('gather_rot', L.ptr(fn), L.ptr(rows), L.ptr(out.permute(0, 2, 3, 1)), L.ll(P), i32(S),
               i32(C), karr, i32(nb), L.stream())
        ctx.save_for_backward(rows)
        ctx.meta = (tuple(feats.shape), karr, nb)
        return feats, out

    @staticmethod
    def backward(ctx, g_full, g_y):
        (rows,) = ctx.saved_tensors
        (K, C, S, S2), karr, nb = ctx.meta
        if g_y is None:
            return g_full, None, None
        if g_full is None:
            g_full = new_nhwc(K, C, S, S2, g_y.device).zero_()
        gf = g_full.permute(0, 2, 3, 1)
        if not gf.is_contiguous():
            gf = gf.contiguous()
            g_full = gf.permute(0, 3, 1, 2)
        L.call('scatter_rot_add', L.ptr(nhwc(g_y)), L.ptr(rows), L.ptr(gf), L.ll(rows.numel()),
               i32(S), i32(C), karr, i32(nb), L.stream())
        return g_full, None, None


def take_rows_rot(feats, rows, ks):
    """(feats, cat_b rot90(feats[rows], ks[b])) with the shared-gradient backward; rows unique."""
    return _TakeRowsRot.apply(feats, rows, tuple(ks))


def take_rows(feats, rows):
    """(feats, feats[rows]) with the shared-gradient backward of `_TakeRows`; `rows` unique."""
    return _TakeRows.apply(feats, rows)


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
              aligned=True):
    """Functional single-level form with the mmcv signature (structures.py:286-287)."""
    if pool_mode != 'avg' or sampling_ratio != 0 or not aligned:
        raise NotImplementedError('LOFT path: RoIAlign is avg / sampling_ratio=0 / aligned=True '
                                  '(bonai_loft_foa_r50_fpn_basic.py:39,56,71)')
    oh, ow = _pair(output_size)
    assert oh == ow
    # a single level: finest_scale large enough that every RoI maps to level 0
    return _MultiLevelRoIAlign.apply(rois, int(oh), [float(spatial_scale)], 1e30, None, input)


class RoIAlign(nn.Module):
    """Same constructor / attributes as mmcv.ops.RoIAlign (output_size tuple, aligned,
    use_torchvision are consulted by the reference, single_level_roi_extractor.py:56,
    tests/test_config.py:276-291)."""

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio,
                         self.pool_mode, self.aligned)

    def __repr__(self):
        return (f'{self.__class__.__name__}(output_size={self.output_size}, '
                f'spatial_scale={self.spatial_scale}, sampling_ratio={self.sampling_ratio}, '
                f'pool_mode={self.pool_mode}, aligned={self.aligned}, '
                f'use_torchvision={self.use_torchvision})')


def mask_target_sample(masks_u8, boxes, gt_inds, size):
    """RoIAlign(size x size, scale 1, aligned) of each proposal over its assigned uint8 GT bitmap,
    thresholded at 0.5 (mask_target.py:31-62 + structures.py:261-291) -- reads the uint8 masks
    in place, no fp32 copy of the bitmaps."""
    P = boxes.shape[0]
    out = torch.empty((P, size, size), device=boxes.device, dtype=torch.float32)
    if P == 0:
        return out
    G, H, W = masks_u8.shape
    assert masks_u8.dtype == torch.uint8 and masks_u8.is_contiguous()
    L.call('mask_target', L.ptr(masks_u8), L.ptr(boxes.contiguous().float()),
           L.ptr(gt_inds.contiguous().long()), L.ll(P), i32(size), i32(H), i32(W), L.ptr(out),
           L.stream())
    return out
