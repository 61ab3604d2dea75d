"""Small HBM-bound ops of the path: stem (im2col + GEMM), 3x3/2 max-pool, FPN extra level,
FOA rot90 rotation."""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib as L
from .dense import nhwc, new_nhwc

i32 = ctypes.c_int


def maxpool3x3s2(x):
    """nn.MaxPool2d(3, 2, 1) forward (resnet.py:571); only used inside the frozen stem."""
    N, C, H, W = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = new_nhwc(N, C, Ho, Wo, x.device)
    L.call('maxpool3x3s2', L.ptr(nhwc(x)), L.ptr(y.permute(0, 2, 3, 1)), i32(N), i32(H), i32(W),
           i32(C), L.stream())
    return y


STEM_K = 7 * 8 * 4          # k extent of the packed stem weight: 7 kernel rows x (8 taps x 4 channels)


def pack_stem_weight(w_khwc, out=None):
    """[Cout,7,7,C<=4] (kernel row, tap, channel) -> [Cout, 224] = [Cout][7][8][4], zero pads: the
    K layout loft_stem_conv7x7 reads (one 32-float k-block per kernel row)."""
    Co, kh, kw, C = w_khwc.shape
    assert kh == 7 and kw == 7 and C <= 4
    if out is None:
        out = torch.zeros((Co, STEM_K), device=w_khwc.device, dtype=torch.float32)
    out.view(Co, 7, 8, 4)[:, :, :7, :C].copy_(w_khwc)
    return out


def stem_packed_shape(N, H, W):
    """Shape of the packed stem image loft_stem_pack writes (see include/loft_b200.h)."""
    return (N, 2, (H + 7) // 2, W + 8, 4)


def stem_conv(img, w_packed, scale, shift, xp=None):
    """7x7/2 pad-3 conv (3->64) + BN-eval + ReLU of the frozen stem (resnet.py:525-571), direct:
    the NCHW fp32 image is packed once (zero-padded NHWC4, TF32-rounded) and one tcgen05 GEMM
    reads its sliding windows through an overlapping-stride tensor map -- no im2col matrix."""
    N, C, H, W = img.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    img = img.contiguous().float()
    if xp is None:
        xp = torch.empty(stem_packed_shape(N, H, W), device=img.device, dtype=torch.float32)
    st = L.stream()
    L.call('stem_pack', L.ptr(img), L.ptr(xp), i32(N), i32(H), i32(W), i32(C), st)
    Cout = w_packed.shape[0]
    y = new_nhwc(N, Cout, Ho, Wo, img.device)
    e = L.make_epilogue(scale=scale, shift=shift, relu=True, round_out=True)
    L.call('stem_conv7x7', L.ptr(xp), L.ptr(w_packed), L.ptr(y.permute(0, 2, 3, 1)), i32(N), i32(H),
           i32(W), i32(Cout), ctypes.byref(e), st)
    return y


class _Subsample2(Function):
    """F.max_pool2d(x, 1, stride=2) == x[:, :, ::2, ::2] (FPN extra level, fpn.py:197-199)."""

    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        y = new_nhwc(N, C, Ho, Wo, x.device)
        L.call('subsample2', L.ptr(nhwc(x)), L.ptr(y.permute(0, 2, 3, 1)), i32(N), i32(H), i32(W),
               i32(C), L.stream())
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        dx = new_nhwc(N, C, H, W, dy.device)
        L.call('subsample2_bwd', L.ptr(nhwc(dy)), L.ptr(dx.permute(0, 2, 3, 1)), None, i32(N),
               i32(H), i32(W), i32(C), L.stream())
        return dx


def subsample2(x):
    return _Subsample2.apply(x)


class _Rot90(Function):
    """Rotation of RoI features by k*90 degrees (k = rotation/90).  The reference builds an affine
    grid and bilinearly resamples (offset_head_expand_feature.py:163-196); for 0/90/180/270 degrees
    that is the rot90 permutation up to 5e-7, which is what this kernel does."""

    @staticmethod
    def forward(ctx, x, k):
        K, C, S, S2 = x.shape
        assert S == S2
        y = new_nhwc(K, C, S, S, x.device)
        L.call('rot90', L.ptr(nhwc(x)), L.ptr(y.permute(0, 2, 3, 1)), L.ll(K), i32(S), i32(C), i32(k),
               L.stream())
        ctx.k = k
        return y

    @staticmethod
    def backward(ctx, dy):
        K, C, S, _ = dy.shape
        dx = new_nhwc(K, C, S, S, dy.device)
        L.call('rot90', L.ptr(nhwc(dy)), L.ptr(dx.permute(0, 2, 3, 1)), L.ll(K), i32(S), i32(C),
               i32(-ctx.k), L.stream())
        return dx, None


def rot90(x, k):
    if x.shape[0] == 0 or k % 4 == 0:
        return x
    return _Rot90.apply(x, int(k))
