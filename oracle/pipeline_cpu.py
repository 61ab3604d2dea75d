"""CPU restatement (numpy) of the reference's BONAI training pipeline after loading:
Resize(keep_ratio, scale factor 1) -> RandomFlip -> Normalize -> Pad(size_divisor) ->
DefaultFormatBundle (configs/_base_/datasets/bonai_instance.py:5-17).

TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/pipeline.npz, produced by
oracle/make_golden_pipeline.py from the UNMODIFIED reference transform classes run over the import
shim (whose mmcv.image functions are themselves a restatement of mmcv 1.0.5 -- that part of the
pin is "from memory", see oracle/shim/mmcv/image/__init__.py)."""
import numpy as np


def rescale_size(w, h, scale):
    """mmcv.rescale_size for a (long edge, short edge) tuple (mmcv/image/geometric.py, v1.0.5)."""
    max_long_edge, max_short_edge = max(scale), min(scale)
    sf = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    return int(w * float(sf) + 0.5), int(h * float(sf) + 0.5)


def _lin_taps(dsize, ssize, clamp_weights):
    """Source indices and 11-bit fixed-point weights of cv2.resize(INTER_LINEAR) along one axis
    (OpenCV imgproc/resize.cpp, resizeGeneric_ set-up): f = (float)((d + .5) * scale - .5).
    Along x the weight is reset to 0 where the tap leaves the image; along y only the ROWS are
    clipped.  PARITY UNPINNED: cv2 is not installed here, this is the published algorithm."""
    scale = 1.0 / (float(dsize) / float(ssize))
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo, hi = s < 0, s >= ssize - 1
        f = np.where(lo | hi, np.float32(0), f)
        s = np.where(lo, 0, np.where(hi, ssize - 1, s))
        s0, s1 = s, np.minimum(s + 1, ssize - 1)
    else:
        s0, s1 = np.clip(s, 0, ssize - 1), np.clip(s + 1, 0, ssize - 1)
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s0, s1, a0, a1


def resize_bilinear_u8(img, new_w, new_h):
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC."""
    h, w = img.shape[:2]
    x0, x1, ax0, ax1 = _lin_taps(new_w, w, True)
    y0, y1, by0, by1 = _lin_taps(new_h, h, False)
    src = img.astype(np.int64)
    hz = src[:, x0] * ax0[None, :, None] + src[:, x1] * ax1[None, :, None]      # [h, new_w, c]
    v = ((by0[:, None, None] * (hz[y0] >> 4)) >> 16) + ((by1[:, None, None] * (hz[y1] >> 4)) >> 16)
    return np.clip((v + 2) >> 2, 0, 255).astype(np.uint8)


def resize_nearest_u8(masks, new_w, new_h):
    """cv2.resize(mask, (new_w, new_h), interpolation=cv2.INTER_NEAREST) per [H,W] bitmap."""
    h, w = masks.shape[-2:]
    sx = np.minimum(np.floor(np.arange(new_w) * (1.0 / (new_w / w))).astype(np.int64), w - 1)
    sy = np.minimum(np.floor(np.arange(new_h) * (1.0 / (new_h / h))).astype(np.int64), h - 1)
    return np.ascontiguousarray(masks[..., sy[:, None], sx[None, :]])


def resize_keep_ratio(img, bboxes, masks, img_scale):
    """Resize(img_scale, keep_ratio=True) (transforms.py:186-253): image bilinear, bitmaps nearest,
    boxes scaled by the float32 (w_scale, h_scale) and clipped; gt_offsets are NOT rescaled by the
    reference.  Returns (img, bboxes, masks, scale_factor[4])."""
    h, w = img.shape[:2]
    nw, nh = rescale_size(w, h, img_scale)
    sf = np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32)
    if (nw, nh) != (w, h):
        img = resize_bilinear_u8(img, nw, nh)
        masks = resize_nearest_u8(masks, nw, nh)
    b = bboxes.astype(np.float32) * sf
    b[:, 0::2] = np.clip(b[:, 0::2], 0, nw)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, nh)
    return img, b, masks, sf


def resize_identity(bboxes, img_hw):
    """Resize with scale factor 1 (1024^2 tiles at img_scale=(1024,1024)): the image and masks are
    unchanged, boxes are still clipped to the image (transforms.py:222-229)."""
    b = bboxes.astype(np.float32) * np.ones(4, dtype=np.float32)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_hw[1])
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_hw[0])
    return b


def flip_image(img, direction):
    """mmcv.imflip (transforms.py:486-488)."""
    return img[:, ::-1] if direction == 'horizontal' else img[::-1]


def flip_bboxes(bboxes, img_hw, direction):
    """RandomFlip.bbox_flip (transforms.py:378-404)."""
    f = bboxes.copy()
    if direction == 'horizontal':
        w = img_hw[1]
        f[..., 0::4] = w - bboxes[..., 2::4]
        f[..., 2::4] = w - bboxes[..., 0::4]
    elif direction == 'vertical':
        h = img_hw[0]
        f[..., 1::4] = h - bboxes[..., 3::4]
        f[..., 3::4] = h - bboxes[..., 1::4]
    else:
        raise ValueError(direction)
    return f


def flip_masks(masks, direction):
    """BitmapMasks.flip (core/mask/structures.py:218-229)."""
    return masks[:, :, ::-1] if direction == 'horizontal' else masks[:, ::-1, :]


def flip_offsets(offsets, direction):
    """RandomFlip.offset_flip (transforms.py:458-466): the roof-to-footprint vector changes sign
    along the flipped axis."""
    f = offsets.astype(np.float32).copy()
    if direction == 'horizontal':
        f[:, 0] = -f[:, 0]
    elif direction == 'vertical':
        f[:, 1] = -f[:, 1]
    else:
        raise ValueError(direction)
    return f


def normalize(img_u8_bgr, mean, std, to_rgb=True):
    """Normalize (transforms.py:655-676) -> mmcv.imnormalize: BGR->RGB, cv2.subtract(img, mean),
    cv2.multiply(img, 1/std) on a float32 image with float64 scalars: the difference is rounded to
    float32, the product is formed in double and rounded once (bit-exact vs the golden vectors)."""
    x = img_u8_bgr.astype(np.float32)
    if to_rgb:
        x = x[..., ::-1]
    m = np.asarray(mean, dtype=np.float32)
    sinv = 1.0 / np.asarray(std, dtype=np.float32).astype(np.float64)
    return ((x - m).astype(np.float64) * sinv).astype(np.float32)


def pad_to_multiple(arr_hw_last2_or_hwc, divisor, channels_last):
    """Pad(size_divisor) with pad_val 0 (transforms.py:571-600, structures.py:231-240)."""
    if channels_last:
        H, W = arr_hw_last2_or_hwc.shape[:2]
    else:
        H, W = arr_hw_last2_or_hwc.shape[-2:]
    Hp, Wp = int(np.ceil(H / divisor)) * divisor, int(np.ceil(W / divisor)) * divisor
    if channels_last:
        out = np.zeros((Hp, Wp) + arr_hw_last2_or_hwc.shape[2:], dtype=arr_hw_last2_or_hwc.dtype)
        out[:H, :W] = arr_hw_last2_or_hwc
    else:
        out = np.zeros(arr_hw_last2_or_hwc.shape[:-2] + (Hp, Wp), dtype=arr_hw_last2_or_hwc.dtype)
        out[..., :H, :W] = arr_hw_last2_or_hwc
    return out


def train_pipeline(img_u8_bgr, gt_bboxes, gt_masks, gt_offsets, flip, direction, mean, std,
                   to_rgb=True, size_divisor=32, img_scale=None):
    """The whole post-load pipeline; returns (img [3,Hp,Wp] float32, bboxes, masks [G,Hp,Wp] uint8,
    offsets, meta) -- `img` in the CHW layout DefaultFormatBundle produces (formating.py:191-230)."""
    H, W = img_u8_bgr.shape[:2]
    sf = 1.0
    if img_scale is not None and rescale_size(W, H, img_scale) != (W, H):
        img_u8_bgr, bboxes, gt_masks, sf = resize_keep_ratio(img_u8_bgr, gt_bboxes, gt_masks,
                                                             img_scale)
        H, W = img_u8_bgr.shape[:2]
    else:
        bboxes = resize_identity(gt_bboxes, (H, W))
    img, masks, offsets = img_u8_bgr, gt_masks, gt_offsets.astype(np.float32)
    if flip:
        img = flip_image(img, direction)
        bboxes = flip_bboxes(bboxes, (H, W), direction)
        masks = flip_masks(masks, direction)
        offsets = flip_offsets(offsets, direction)
    x = normalize(np.ascontiguousarray(img), mean, std, to_rgb)
    x = pad_to_multiple(x, size_divisor, channels_last=True)
    masks = pad_to_multiple(np.ascontiguousarray(masks), size_divisor, channels_last=False)
    meta = dict(img_shape=(H, W, 3), pad_shape=x.shape, scale_factor=sf, flip=bool(flip),
                flip_direction=direction)
    return np.ascontiguousarray(x.transpose(2, 0, 1)), bboxes, masks, offsets, meta
