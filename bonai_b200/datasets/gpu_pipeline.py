"""Device-side counterpart of the reference's post-load training pipeline for BONAI tiles
(configs/_base_/datasets/bonai_instance.py:5-17): Resize(keep_ratio) -> RandomFlip -> Normalize
-> Pad(32) -> DefaultFormatBundle -> Collect, applied to a uint8 tile and its annotations.

In the reference this runs in DataLoader worker processes on the CPU and ships a float32 image
(12 MB) plus uint8 bitmaps (1 MB per building) per tile; at >100 tiles/s per GPU those two worker
processes are the bottleneck.  Here the tile crosses PCIe as uint8 (3 MB) and two
kernels do the rest in one pass each (csrc/pipeline.cu); the building polygons are rasterised on
the device too (`polygons_to_bitmaps`).  For BONAI the Resize is the identity (1024^2 tiles at
img_scale=(1024,1024)); other tile sizes go through the restated cv2.resize kernels
(`resize_bilinear_u8`, `resize_nearest_u8`).  Image decoding stays on the host (datasets/bonai.py).
"""
import ctypes

import numpy as np
import torch

from .. import _lib as L
from ..core.mask import BitmapMasks

i32 = ctypes.c_int
_FLIP = {None: 0, 'horizontal': 1, 'vertical': 2}


def _ceil_to(v, d):
    return (v + d - 1) // d * d


def image_prep(img_u8, mean, std, to_rgb=True, flip=None, size_divisor=32):
    """uint8 [H,W,3] (BGR, as mmcv.imread loads it) device tensor -> float32 [3,Hp,Wp]:
    flip, BGR->RGB, (x-mean)*(1/std), zero pad to a multiple of `size_divisor`, HWC->CHW."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and \
        img_u8.shape[2] == 3
    img_u8 = img_u8.contiguous()
    H, W = int(img_u8.shape[0]), int(img_u8.shape[1])
    d = max(int(size_divisor or 1), 1)
    Hp, Wp = _ceil_to(H, d), _ceil_to(W, d)
    if Wp % 4:
        raise L.LoftError('image_prep: padded width must be a multiple of 4')
    out = torch.empty((3, Hp, Wp), device=img_u8.device, dtype=torch.float32)
    m = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(mean, dtype=np.float32)])
    s = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(std, dtype=np.float32)])
    L.call('image_prep', L.ptr(img_u8), L.ptr(out), i32(H), i32(W), i32(Hp), i32(Wp), m, s,
           i32(1 if to_rgb else 0), i32(_FLIP[flip]), L.stream())
    return out


def mask_flip_pad(masks_u8, flip=None, size_divisor=32):
    """uint8 [G,H,W] device bitmaps -> flipped, zero-padded [G,Hp,Wp] (BitmapMasks.flip + pad)."""
    assert masks_u8.is_cuda and masks_u8.dtype == torch.uint8 and masks_u8.dim() == 3
    masks_u8 = masks_u8.contiguous()
    G, H, W = (int(v) for v in masks_u8.shape)
    d = max(int(size_divisor or 1), 1)
    Hp, Wp = _ceil_to(H, d), _ceil_to(W, d)
    if flip is None and (Hp, Wp) == (H, W):
        return masks_u8
    if Wp % 16:
        raise L.LoftError('mask_flip_pad: padded width must be a multiple of 16')
    out = torch.empty((G, Hp, Wp), device=masks_u8.device, dtype=torch.uint8)
    L.call('mask_flip_pad', L.ptr(masks_u8), L.ptr(out), L.ll(G), i32(H), i32(W), i32(Hp), i32(Wp),
           i32(_FLIP[flip]), L.stream())
    return out


def resize_bilinear_u8(img_u8, new_h, new_w):
    """cv2.resize(INTER_LINEAR) of a uint8 [H,W,3] device tensor (`loft_resize_bilinear_u8`)."""
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 3 and \
        img_u8.shape[2] == 3
    img_u8 = img_u8.contiguous()
    out = torch.empty((new_h, new_w, 3), device=img_u8.device, dtype=torch.uint8)
    L.call('resize_bilinear_u8', L.ptr(img_u8), L.ptr(out), i32(img_u8.shape[0]),
           i32(img_u8.shape[1]), i32(new_h), i32(new_w), L.stream())
    return out


def resize_nearest_u8(masks_u8, new_h, new_w):
    """cv2.resize(INTER_NEAREST) of uint8 [G,H,W] device bitmaps (BitmapMasks.rescale)."""
    assert masks_u8.is_cuda and masks_u8.dtype == torch.uint8 and masks_u8.dim() == 3
    masks_u8 = masks_u8.contiguous()
    G, H, W = (int(v) for v in masks_u8.shape)
    out = torch.empty((G, new_h, new_w), device=masks_u8.device, dtype=torch.uint8)
    L.call('resize_nearest_u8', L.ptr(masks_u8), L.ptr(out), L.ll(G), i32(H), i32(W), i32(new_h),
           i32(new_w), L.stream())
    return out


def polygons_to_bitmaps(masks_ann, H, W, device='cuda'):
    """LoadAnnotations._load_masks with poly2mask=True (loading.py:301-326,345-368) on the device:
    `masks_ann` = per instance a list of polygon parts (flat [x0, y0, x1, y1, ...]); returns the
    uint8 [G,H,W] bitmaps pycocotools' frPyObjects -> merge -> decode would produce
    (`loft_poly_rasterize`).  As in frPyObjects, an instance whose first part has exactly 4
    numbers is a list of [x, y, w, h] boxes."""
    dev = torch.device(device)
    G = len(masks_ann)
    out = torch.empty((G, H, W), device=dev, dtype=torch.uint8)
    if G == 0:
        return out
    xy, part_off, inst_off = [], [0], [0]
    for parts in masks_ann:
        if isinstance(parts, dict):
            raise NotImplementedError('RLE mask annotations are not on the BONAI path')
        as_bbox = len(parts) > 0 and len(parts[0]) == 4
        for p in parts:
            p = [float(v) for v in p]
            if as_bbox:
                xs, ys, xe, ye = p[0], p[1], p[0] + p[2], p[1] + p[3]
                p = [xs, ys, xs, ye, xe, ye, xe, ys]
            if len(p) % 2:
                raise L.LoftError('polygon with an odd number of coordinates')
            xy.extend(p)
            part_off.append(part_off[-1] + len(p) // 2)
        inst_off.append(len(part_off) - 1)
    n_parts, n_vert = len(part_off) - 1, part_off[-1]
    xy_d = torch.tensor(xy if xy else [0.0], dtype=torch.float64).to(dev, non_blocking=True)
    po_d = torch.tensor(part_off, dtype=torch.int64).to(dev, non_blocking=True)
    io_d = torch.tensor(inst_off, dtype=torch.int32).to(dev, non_blocking=True)
    lib = L.lib()
    lib.loft_poly_scratch_bytes.restype = ctypes.c_longlong
    nbytes = int(lib.loft_poly_scratch_bytes(ctypes.c_longlong(n_vert), i32(n_parts)))
    scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    L.call('poly_rasterize', L.ptr(xy_d), L.ptr(po_d), L.ptr(io_d), i32(G), i32(n_parts),
           L.ll(n_vert), i32(H), i32(W), L.ptr(out), L.ptr(scratch), L.ptr(err), L.stream())
    e = int(err.item())            # loader path, not the training step: one read-back per tile
    if e:
        raise L.LoftError(f'polygon part {e - 1} has more than 8192 run boundaries')
    return out


class GpuTrainPipeline:
    """Same parameters as the reference's transform dicts.  `__call__` takes what
    LoadImageFromFile + LoadAnnotations produce for one tile and returns the keys `Collect`
    passes on (img, gt_bboxes, gt_labels, gt_masks, gt_offsets) plus the img_meta dict, all on
    the device, ready for `LOFT.forward_train`."""

    def __init__(self, img_scale=(1024, 1024), keep_ratio=True, flip_ratio=0.5,
                 direction=('horizontal', 'vertical'), mean=(123.675, 116.28, 103.53),
                 std=(58.395, 57.12, 57.375), to_rgb=True, size_divisor=32, device='cuda',
                 rng=None):
        self.img_scale, self.keep_ratio = tuple(img_scale), keep_ratio
        self.flip_ratio = flip_ratio
        self.rng = rng if rng is not None else np.random
        # the reference draws the direction ONCE, when the transform is constructed
        # (transforms.py:366-371), not per sample
        if isinstance(direction, str):
            self.direction = direction
        else:
            direction = list(direction)
            assert all(d in ('horizontal', 'vertical') for d in direction)
            self.direction = str(self.rng.choice(direction))
        self.mean = np.asarray(mean, dtype=np.float32)
        self.std = np.asarray(std, dtype=np.float32)
        self.to_rgb, self.size_divisor = to_rgb, size_divisor
        self.device = torch.device(device)

    @classmethod
    def from_cfg(cls, train_pipeline, **kw):
        """Build from the reference's `train_pipeline` list of dicts."""
        a = {}
        for t in train_pipeline:
            t = dict(t)
            typ = t.pop('type')
            if typ == 'Resize':
                a.update(img_scale=t.get('img_scale', (1024, 1024)),
                         keep_ratio=t.get('keep_ratio', True))
            elif typ == 'RandomFlip':
                a.update(flip_ratio=t.get('flip_ratio'), direction=t.get('direction', 'horizontal'))
            elif typ == 'Normalize':
                a.update(mean=t['mean'], std=t['std'], to_rgb=t.get('to_rgb', True))
            elif typ == 'Pad':
                if t.get('size') is not None:
                    raise NotImplementedError('Pad(size=...) is not on the BONAI path')
                a.update(size_divisor=t.get('size_divisor'))
            elif typ not in ('LoadImageFromFile', 'LoadAnnotations', 'DefaultFormatBundle',
                             'Collect'):
                raise NotImplementedError(f'transform {typ} is not on the BONAI path')
        a.update(kw)
        return cls(**a)

    def _scale_factor(self, H, W):
        if not self.keep_ratio:
            raise NotImplementedError('LOFT path: Resize(keep_ratio=True)')
        long_e, short_e = max(self.img_scale), min(self.img_scale)
        return min(long_e / max(H, W), short_e / min(H, W))

    def __call__(self, img, gt_bboxes, gt_labels, gt_masks, gt_offsets, flip=None):
        dev = self.device
        img = torch.as_tensor(img)
        H0, W0 = int(img.shape[0]), int(img.shape[1])
        sf = self._scale_factor(H0, W0)
        W, H = int(W0 * float(sf) + 0.5), int(H0 * float(sf) + 0.5)       # mmcv.rescale_size
        resample = (H, W) != (H0, W0)
        if flip is None:
            flip = bool(self.rng.rand() < self.flip_ratio) if self.flip_ratio is not None else False
        d = self.direction if flip else None
        img_d = img.to(dev, non_blocking=True)
        if resample:       # not the BONAI case (1024^2 tiles at img_scale 1024): cv2.resize restated
            img_d = resize_bilinear_u8(img_d, H, W)
        x = image_prep(img_d, self.mean, self.std, self.to_rgb, d, self.size_divisor)
        scale4 = np.array([W / W0, H / H0, W / W0, H / H0], dtype=np.float32)
        # Resize clips the boxes to the image even at scale 1 (transforms.py:222-229)
        b = torch.as_tensor(gt_bboxes, dtype=torch.float32).to(dev, non_blocking=True).clone()
        if resample:
            b = b * torch.from_numpy(scale4).to(dev)
        b[:, 0::2].clamp_(0, W)
        b[:, 1::2].clamp_(0, H)
        o = torch.as_tensor(gt_offsets, dtype=torch.float32).to(dev, non_blocking=True).clone()
        if flip:
            if d == 'horizontal':                       # bbox_flip / offset_flip, :378-404,458-466
                b = torch.stack([W - b[:, 2], b[:, 1], W - b[:, 0], b[:, 3]], 1)
                o[:, 0] = -o[:, 0]
            else:
                b = torch.stack([b[:, 0], H - b[:, 3], b[:, 2], H - b[:, 1]], 1)
                o[:, 1] = -o[:, 1]
        m = gt_masks.to_tensor(device=dev) if isinstance(gt_masks, BitmapMasks) else \
            torch.as_tensor(gt_masks).to(dev, non_blocking=True)
        m = m.to(torch.uint8)
        if resample:
            m = resize_nearest_u8(m, H, W)
        m = mask_flip_pad(m, d, self.size_divisor)
        Hp, Wp = int(x.shape[1]), int(x.shape[2])
        meta = dict(img_shape=(H, W, 3), ori_shape=(H0, W0, 3), pad_shape=(Hp, Wp, 3),
                    scale_factor=scale4, flip=bool(flip),
                    flip_direction=self.direction,
                    img_norm_cfg=dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb))
        labels = torch.as_tensor(gt_labels, dtype=torch.long).to(dev, non_blocking=True)
        return dict(img=x, img_metas=meta, gt_bboxes=b, gt_labels=labels,
                    gt_masks=BitmapMasks(m, Hp, Wp), gt_offsets=o)

    @staticmethod
    def collate(samples):
        """mmcv `collate` for this path: stack the (equally padded) images, keep the rest as
        per-image lists -- the argument layout of `forward_train`."""
        return dict(img=torch.stack([s['img'] for s in samples]),
                    img_metas=[s['img_metas'] for s in samples],
                    gt_bboxes=[s['gt_bboxes'] for s in samples],
                    gt_labels=[s['gt_labels'] for s in samples],
                    gt_masks=[s['gt_masks'] for s in samples],
                    gt_offsets=[s['gt_offsets'] for s in samples])
