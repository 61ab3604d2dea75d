"""mmcv.image pieces the reference's training pipeline calls (mmdet/datasets/pipelines/
transforms.py:286-330,484-488,560-575,671-672; core/mask/structures.py:218-240), restated from
mmcv v1.0.5 [from memory -- the package is absent here]: cv2-backed flip / normalize / pad /
rescale.  TEST INFRASTRUCTURE ONLY (oracle/shim)."""
import cv2
import numpy as np


def imflip(img, direction='horizontal'):
    assert direction in ['horizontal', 'vertical']
    if direction == 'horizontal':
        return np.flip(img, axis=1)
    return np.flip(img, axis=0)


def imnormalize_(img, mean, std, to_rgb=True):
    assert img.dtype != np.uint8
    mean = np.float64(mean.reshape(1, -1))
    stdinv = 1 / np.float64(std.reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean, img)
    cv2.multiply(img, stdinv, img)
    return img


def imnormalize(img, mean, std, to_rgb=True):
    img = img.copy().astype(np.float32)
    return imnormalize_(img, mean, std, to_rgb)


def impad(img, shape, pad_val=0):
    if not isinstance(pad_val, (int, float)):
        assert len(pad_val) == img.shape[-1]
    if len(shape) < len(img.shape):
        shape = tuple(shape) + (img.shape[-1],)
    assert len(shape) == len(img.shape)
    for s, i in zip(shape, img.shape):
        assert s >= i
    pad = np.empty(shape, dtype=img.dtype)
    pad[...] = pad_val
    pad[:img.shape[0], :img.shape[1], ...] = img
    return pad


def impad_to_multiple(img, divisor, pad_val=0):
    pad_h = int(np.ceil(img.shape[0] / divisor)) * divisor
    pad_w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return impad(img, (pad_h, pad_w), pad_val)


def rescale_size(old_size, scale, return_scale=False):
    w, h = old_size
    if isinstance(scale, (float, int)):
        scale_factor = scale
    else:
        max_long_edge, max_short_edge = max(scale), min(scale)
        scale_factor = min(max_long_edge / max(h, w), max_short_edge / min(h, w))
    new_size = int(w * float(scale_factor) + 0.5), int(h * float(scale_factor) + 0.5)
    return (new_size, scale_factor) if return_scale else new_size


def imresize(img, size, return_scale=False, interpolation='bilinear', out=None, backend=None):
    h, w = img.shape[:2]
    codes = dict(nearest=cv2.INTER_NEAREST, bilinear=cv2.INTER_LINEAR, bicubic=cv2.INTER_CUBIC,
                 area=cv2.INTER_AREA, lanczos=cv2.INTER_LANCZOS4)
    out = cv2.resize(img, size, interpolation=codes[interpolation])
    if not return_scale:
        return out
    return out, size[0] / w, size[1] / h


def imrescale(img, scale, return_scale=False, interpolation='bilinear', backend=None):
    h, w = img.shape[:2]
    new_size, scale_factor = rescale_size((w, h), scale, return_scale=True)
    out = imresize(img, new_size, interpolation=interpolation)
    return (out, scale_factor) if return_scale else out
