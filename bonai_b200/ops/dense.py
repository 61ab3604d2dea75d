"""autograd bindings of the tcgen05 dense kernels (conv / linear / deconv) with fused epilogues.

Activations are logically NCHW (the reference's module protocol) but physically NHWC
(torch.channels_last); every Function allocates its outputs that way so no layout conversion
happens between layers.  Weight / bias / BN-beta gradients are accumulated by the kernels
directly into the ParamStore's flat gradient buffer (BN-gamma gradients are derived from the
weight gradients afterwards, engine._bn_finalize); the Functions return ``None`` for those inputs
(they are passed as tensors only so autograd schedules the node).  The static trunk of the LOFT
model does not go through these Functions in training -- see bonai_b200/trunk.py.
"""
import ctypes
import os

import torch
from torch.autograd import Function

from .. import _lib as L

i32 = ctypes.c_int


def nhwc(t):
    """[N,C,H,W] tensor -> contiguous [N,H,W,C] view (copies only if not already channels_last)."""
    if t.dim() != 4:
        return t.contiguous()
    p = t.permute(0, 2, 3, 1)
    if not p.is_contiguous():
        p = p.contiguous()
    return p


def new_nhwc(n, c, h, w, device):
    """Uninitialised logical-NCHW tensor with NHWC storage."""
    return torch.empty((n, h, w, c), device=device, dtype=torch.float32).permute(0, 3, 1, 2)


def _queue_finalize(store):
    if store is not None:
        store.queue_finalize()


class _maybe_side_stream:
    """Context: run the enclosed wgrad launch on the store's side stream when the launch is too
    small to fill the GPU (work items < SM count), else stay on the current stream."""

    def __init__(self, store, small, *tensors):
        self.ws = None
        if small and store is not None and hasattr(store, 'side_stream_for_wgrad'):
            self.ws = store.side_stream_for_wgrad(*tensors)
        self.ctx = torch.cuda.stream(self.ws) if self.ws is not None else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


class ConvSpec:
    """Static description of one conv / linear layer instance (weights + fused epilogue)."""

    def __init__(self, wref, ksize=1, stride=1, padding=0, relu=False, bias=None, bias_grad=None,
                 bn=None, bn_trainable=False, round_out=True, res_upsample=False, store=None,
                 cout=None, kpad=None, premask_in=False, grad_premasked=False):
        # bias gradient by the consumer: a layer whose (single) consumer has in_colsum set gets
        # its bias gradient as the column sum the consumer's dgrad epilogue accumulates, and
        # skips its own read-only reduction pass (see link_bias_colsum)
        self.in_colsum, self.in_colsum_gstride = None, 0
        self.bias_by_consumer = False
        self.wref = wref
        self.ksize, self.stride, self.padding = ksize, stride, padding
        self.relu = relu
        self.bias, self.bias_grad = bias, bias_grad
        # eval-mode BN is folded into the weights the tensor cores read (engine.fold_bn): the
        # epilogue only adds bn.shift, beta's gradient is the layer's "bias gradient" and gamma's
        # comes from the weight gradient (loft_bn_finalize) -- no pass over the activations
        self.bn, self.bn_trainable = bn, bn_trainable
        if bn is not None:
            self.bias = bn.shift
            self.bias_grad = bn.dbeta if bn_trainable else None
        self.round_out = round_out
        self.res_upsample = res_upsample
        self.store = store
        self.cout = cout
        self.kpad = kpad
        # ReLU-backward fusion along single-consumer chains: a layer with premask_in applies the
        # ReLU mask of its INPUT (x > 0) in its dgrad epilogue, so the producer of x (which sets
        # grad_premasked) receives an already-masked gradient and skips its own masking pass.
        self.premask_in = premask_in
        self.n_out = None          # real rows of a zero-padded narrow head weight (narrow_head)
        self.grad_premasked = grad_premasked


def link_bias_colsum(producer, consumer, gstride=0):
    """Wire producer -> consumer (a single-consumer chain): the consumer's dgrad epilogue writes
    dZ of the producer (it applies the producer's ReLU mask, or the producer has no ReLU), so the
    per-channel sum of what it writes IS the producer's bias gradient."""
    import os
    if os.environ.get('LOFT_NO_COLSUM', '0') != '0':
        return False
    pg = producer.bias_grads[0] if isinstance(producer, GroupedConvSpec) else producer.bias_grad
    if pg is None:
        return False
    if producer.relu and not (consumer.premask_in and producer.grad_premasked):
        return False
    if isinstance(consumer, ConvSpec) and not (
            (consumer.ksize == 3 and consumer.stride == 1 and consumer.padding == 1) or
            (consumer.ksize == 1 and consumer.stride == 1)):
        return False
    consumer.in_colsum = pg
    consumer.in_colsum_gstride = gstride
    producer.bias_by_consumer = True
    return True


def link_chain(specs):
    for a, b in zip(specs[:-1], specs[1:]):
        gs = a.b_grad_gstride if isinstance(a, GroupedConvSpec) else 0
        link_bias_colsum(a, b, gs)


def _fprop(spec, x, residual):
    """Returns (y, saved_input_for_wgrad)."""
    w = spec.wref.w
    dev = x.device
    N, Cin, H, W = x.shape
    Cout = spec.cout if spec.cout is not None else w.shape[0]
    k, s, pad = spec.ksize, spec.stride, spec.padding
    Ho = (H + 2 * pad - k) // s + 1
    Wo = (W + 2 * pad - k) // s + 1
    xn = nhwc(x)
    y = new_nhwc(N, Cout, Ho, Wo, dev)
    yn = y.permute(0, 2, 3, 1)
    rn = nhwc(residual) if residual is not None else None
    e = L.make_epilogue(shift=spec.bias, residual=rn,
                        ldr=(rn.shape[-1] if rn is not None else 0),
                        res_upsample2x=spec.res_upsample, relu=spec.relu,
                        round_out=spec.round_out)
    st = L.stream()
    if k == 3 and s == 1 and pad == 1:
        L.call('conv3x3_fprop', L.ptr(xn), L.ptr(w), L.ptr(yn), i32(N), i32(H), i32(W), i32(Cin),
               i32(Cout), ctypes.byref(e), st)
        saved = xn
    elif k == 1 and pad == 0:
        if s == 2:
            xs = torch.empty((N, Ho, Wo, Cin), device=dev, dtype=torch.float32)
            L.call('subsample2', L.ptr(xn), L.ptr(xs), i32(N), i32(H), i32(W), i32(Cin), st)
        else:
            assert s == 1
            xs = xn
        P = N * Ho * Wo
        L.call('gemm_fprop', L.ptr(xs), L.ptr(w), L.ptr(yn), L.ll(P), i32(Cin), i32(Cout),
               L.ll(Cin), L.ll(Cin), L.ll(Cout), i32(Ho), i32(Wo), ctypes.byref(e), st)
        saved = xs
    else:
        K = k * k * Cin
        Kp = spec.kpad or K
        col = torch.empty((N * Ho * Wo, Kp), device=dev, dtype=torch.float32)
        L.call('im2col', L.ptr(xn), L.ptr(col), i32(N), i32(H), i32(W), i32(Cin), i32(k), i32(k),
               i32(s), i32(pad), i32(Kp), i32(0), st)
        L.call('gemm_fprop', L.ptr(col), L.ptr(w), L.ptr(yn), L.ll(N * Ho * Wo), i32(Kp), i32(Cout),
               L.ll(Kp), L.ll(Kp), L.ll(Cout), i32(Ho), i32(Wo), ctypes.byref(e), st)
        saved = col
    return y, saved


class _ConvFn(Function):
    """y = act(affine(conv(x, W)) + residual).  Inputs after `x` are only autograd triggers."""

    @staticmethod
    def forward(ctx, x, residual, spec, *triggers):
        y, saved = _fprop(spec, x, residual)
        ctx.spec = spec
        ctx.x_shape = tuple(x.shape)
        ctx.has_res = residual is not None
        ctx.res_shape = tuple(residual.shape) if residual is not None else None
        relu_eff = spec.relu and not spec.grad_premasked
        xin = nhwc(x) if spec.premask_in else None
        ctx.save_for_backward(saved, y if relu_eff else None, xin)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        saved, y, xin = ctx.saved_tensors
        relu_eff = spec.relu and not spec.grad_premasked
        _queue_finalize(spec.store)
        st = L.stream()
        N, Cin, H, W = ctx.x_shape
        dyn = nhwc(dy)
        _, Ho, Wo, Cout = dyn.shape
        P = N * Ho * Wo
        dev = dy.device
        # 1. ReLU backward (+ the bias / BN-beta gradient = per-channel sum of dz)
        need_res = ctx.has_res and ctx.needs_input_grad[1]
        dbeta = spec.bias_grad
        dres = None
        if not relu_eff:
            # no ReLU, or its mask was already applied by the consumer's dgrad epilogue:
            # dz == dy, only the bias gradient needs a (read-only) pass unless the consumer's
            # epilogue produced it too.  dy is fed to the tensor cores as is; unrounded operands
            # only bias *gradients* by ~5e-4.
            dz = dyn
            if dbeta is not None and not spec.bias_by_consumer:
                L.call('act_bwd', L.ptr(dyn), None, None, None, None, None, None, None, None,
                       L.ptr(dbeta), L.ll(P), i32(Cout), i32(0), st)
        else:
            dz = torch.empty_like(dyn)
            L.call('act_bwd', L.ptr(dyn), L.ptr(nhwc(y)), None, None, None, None, L.ptr(dz), None,
                   None, L.ptr(dbeta) if dbeta is not None else None, L.ll(P), i32(Cout), i32(1),
                   st)
            dres = dz
        grad_res = None
        if need_res:
            if spec.res_upsample:
                # residual was nearest-upsampled x2: its gradient is the 2x2 sum of g
                g = dz if not relu_eff else dres
                rn, rc, rh, rw = ctx.res_shape
                out = torch.empty((rn, rh, rw, rc), device=dev, dtype=torch.float32)
                L.call('sum2x2_add', L.ptr(g), None, L.ptr(out), i32(rn), i32(rh), i32(rw), i32(rc),
                       st)
                grad_res = out.permute(0, 3, 1, 2)
            elif relu_eff:
                grad_res = dres.permute(0, 3, 1, 2)
            else:
                grad_res = dy
        # 2. weight gradient
        k, s, pad = spec.ksize, spec.stride, spec.padding
        w = spec.wref.w
        gw = spec.wref.grad
        conv3 = (k == 3 and s == 1 and pad == 1)
        if gw is not None:
            # small layers: overlap the weight gradient with the data-gradient chain
            small = P <= 16384 and ctx.needs_input_grad[0]
            with _maybe_side_stream(spec.store, small, dz, saved):
                st2 = L.stream()
                if conv3:
                    L.call('conv3x3_wgrad', L.ptr(dz), L.ptr(saved), L.ptr(gw), i32(N), i32(H),
                           i32(W), i32(Cin), i32(Cout), st2)
                elif Cout % 32 != 0:
                    _narrow_wgrad(dz.view(P, Cout), saved.view(P, -1), gw, st2)
                else:
                    K = saved.shape[-1]
                    L.call('gemm_wgrad', L.ptr(dz), L.ptr(saved), L.ptr(gw), L.ll(P), i32(K),
                           i32(Cout), L.ll(Cout), L.ll(K), L.ll(K), st2)
        # 3. data gradient
        dx = None
        if ctx.needs_input_grad[0]:
            e = L.make_epilogue(round_out=True, mask=xin if (xin is not None and s == 1) else None,
                                colsum=spec.in_colsum)
            if conv3:
                dx = new_nhwc(N, Cin, H, W, dev)
                L.call('conv3x3_dgrad', L.ptr(dz), L.ptr(w), L.ptr(dx.permute(0, 2, 3, 1)), i32(N),
                       i32(H), i32(W), i32(Cin), i32(Cout), ctypes.byref(e), st)
            elif k == 1:
                dxs = torch.empty((N, Ho, Wo, Cin), device=dev, dtype=torch.float32)
                L.call('gemm_dgrad', L.ptr(dz), L.ptr(w), L.ptr(dxs), L.ll(P), i32(Cin), i32(Cout),
                       L.ll(Cout), L.ll(Cin), L.ll(Cin), ctypes.byref(e), st)
                if s == 2:
                    dx = new_nhwc(N, Cin, H, W, dev)
                    L.call('subsample2_bwd', L.ptr(dxs), L.ptr(dx.permute(0, 2, 3, 1)), L.ptr(xin),
                           i32(N), i32(H), i32(W), i32(Cin), st)
                else:
                    dx = dxs.permute(0, 3, 1, 2)
            else:
                K = saved.shape[-1]
                dcol = torch.empty((P, K), device=dev, dtype=torch.float32)
                L.call('gemm_dgrad', L.ptr(dz), L.ptr(w), L.ptr(dcol), L.ll(P), i32(K), i32(Cout),
                       L.ll(Cout), L.ll(K), L.ll(K), None, st)
                dx = new_nhwc(N, Cin, H, W, dev)
                L.call('col2im', L.ptr(dcol), L.ptr(dx.permute(0, 2, 3, 1)), L.ptr(xin), i32(N),
                       i32(H), i32(W), i32(Cin), i32(k), i32(k), i32(s), i32(pad), i32(K), st)
        return (dx, grad_res, None) + (None,) * (len(ctx.needs_input_grad) - 3)


def conv(x, spec, residual=None, triggers=()):
    """Run one fused conv layer.  `triggers`: parameters whose presence makes autograd visit the
    node even if `x` carries no gradient (first trainable layer after the frozen stages)."""
    if x.shape[0] == 0:
        N, _, H, W = x.shape
        k, s, pad = spec.ksize, spec.stride, spec.padding
        return x.new_zeros((0, spec.cout or spec.wref.w.shape[0], (H + 2 * pad - k) // s + 1,
                            (W + 2 * pad - k) // s + 1))
    return _ConvFn.apply(x, residual, spec, *triggers)


class _NarrowHeadFn(Function):
    """1x1 conv with <= 4 (padded) output channels and no activation as HBM-bound warp-per-pixel
    kernels (`loft_narrow_head_fwd/bwd`) instead of a 128-row tensor-core tile that is almost
    entirely padding: same ConvSpec protocol as `_ConvFn` (premask_in, in_colsum, bias_grad)."""

    @staticmethod
    def forward(ctx, x, spec, *triggers):
        N, C, H, W = x.shape
        xn = nhwc(x)
        w = spec.wref.w
        y = new_nhwc(N, w.shape[0], H, W, x.device)
        L.call('narrow_head_fwd', L.ptr(xn), L.ptr(w), L.ptr(spec.bias), L.ptr(y.permute(0, 2, 3, 1)),
               L.ll(N * H * W), i32(C), i32(spec.n_out or 4), i32(0), i32(0), L.stream())
        ctx.spec = spec
        ctx.save_for_backward(xn)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        (xn,) = ctx.saved_tensors
        _queue_finalize(spec.store)
        N, H, W, C = xn.shape
        dyn = nhwc(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = new_nhwc(N, C, H, W, dy.device)
        L.call('narrow_head_bwd', L.ptr(dyn), L.ptr(xn), L.ptr(spec.wref.w),
               L.ptr(dx.permute(0, 2, 3, 1)) if dx is not None else None, L.ptr(spec.wref.grad),
               L.ptr(spec.bias_grad) if not spec.bias_by_consumer else None,
               L.ptr(spec.in_colsum) if dx is not None else None, L.ll(N * H * W), i32(C),
               i32(spec.n_out or 4), i32(1 if spec.premask_in else 0), i32(0), i32(0), L.stream())
        return (dx, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def narrow_head_ok(spec, x):
    """`conv(x, spec)` can run as `_NarrowHeadFn`: 1x1/s1, <= 4 padded outputs, no activation,
    no output rounding, 128 / 256 / 512 input channels."""
    return (os.environ.get('LOFT_NARROW_HEAD', '1') != '0' and spec.ksize == 1 and
            spec.stride == 1 and spec.padding == 0 and not spec.relu and spec.bn is None and
            not spec.round_out and spec.wref.w.dim() == 2 and spec.wref.w.shape[0] == 4 and
            x.dim() == 4 and x.shape[1] in (128, 256, 512) and spec.wref.w.shape[1] == x.shape[1])


def narrow_head(x, spec, triggers=()):
    if x.shape[0] == 0:
        N, _, H, W = x.shape
        return x.new_zeros((0, spec.wref.w.shape[0], H, W))
    return _NarrowHeadFn.apply(x, spec, *triggers)


def conv_nograd(x, spec, residual=None):
    """Forward only (frozen stages / inference)."""
    return _fprop(spec, x, residual)[0]


# ----------------------------------------------------------------------------------- linear
class _LinearFn(Function):
    """y[P, Cout_pad] = act(x[P,K] W^T + b); output may be wider than the logical Cout (padding to
    a TMA-legal pitch) -- the caller slices."""

    @staticmethod
    def forward(ctx, x, spec, *triggers):
        w = spec.wref.w
        P, K = x.shape
        Cout = w.shape[0]
        x = x.contiguous()
        y = torch.empty((P, Cout), device=x.device, dtype=torch.float32)
        e = L.make_epilogue(shift=spec.bias, relu=spec.relu, round_out=spec.round_out)
        L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Cout), L.ll(K),
               L.ll(K), L.ll(Cout), i32(1), i32(P), ctypes.byref(e), L.stream())
        ctx.spec = spec
        ctx.save_for_backward(x, y if (spec.relu and not spec.grad_premasked) else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        x, y = ctx.saved_tensors
        _queue_finalize(spec.store)
        st = L.stream()
        P, K = x.shape
        w = spec.wref.w
        Cout = w.shape[0]
        dy = dy.contiguous()
        if not (spec.relu and not spec.grad_premasked):
            dz = dy
            if spec.bias_grad is not None and not spec.bias_by_consumer:
                L.call('act_bwd', L.ptr(dy), None, None, None, None, None, None, None, None,
                       L.ptr(spec.bias_grad), L.ll(P), i32(Cout), i32(0), st)
        else:
            dz = torch.empty_like(dy)
            L.call('act_bwd', L.ptr(dy), L.ptr(y), None, None, None, None, L.ptr(dz), None, None,
                   L.ptr(spec.bias_grad) if spec.bias_grad is not None else None, L.ll(P),
                   i32(Cout), i32(spec.relu), st)
        if spec.wref.grad is not None:
            if Cout % 32 == 0:
                with _maybe_side_stream(spec.store, ctx.needs_input_grad[0], dz, x):
                    L.call('gemm_wgrad', L.ptr(dz), L.ptr(x), L.ptr(spec.wref.grad), L.ll(P), i32(K),
                           i32(Cout), L.ll(Cout), L.ll(K), L.ll(K), L.stream())
            else:
                # narrow heads (Cout < 32): dW^T[K, Cout] = x^T dz via the same kernel with the
                # roles swapped would need Cout%32==0 as well; use the transposed problem
                # dW[Cout,K] = dz^T x computed as a "dgrad": out[k, c] = sum_p x[p,k] dz[p,c]
                _narrow_wgrad(dz, x, spec.wref.grad, st)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((P, K), device=dy.device, dtype=torch.float32)
            e = L.make_epilogue(round_out=True, mask=x if spec.premask_in else None,
                                colsum=spec.in_colsum)
            L.call('gemm_dgrad', L.ptr(dz), L.ptr(w), L.ptr(dx), L.ll(P), i32(K), i32(Cout),
                   L.ll(Cout), L.ll(K), L.ll(K), ctypes.byref(e), st)
        return (dx, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def _narrow_wgrad(dz, x, gw, st):
    """dW[Cout,K] += dz[P,Cout]^T x[P,K] for Cout < 32: pad dz's columns to 32 in a scratch
    buffer (the MN-major TMA view needs 32-wide channel chunks)."""
    P, Cout = dz.shape
    K = x.shape[1]
    pad = torch.zeros((P, 32), device=dz.device, dtype=torch.float32)
    L.call('copy2d', L.ptr(dz), L.ll(Cout), L.ptr(pad), L.ll(32), L.ll(P), i32(Cout), i32(0), i32(0),
           st)
    tmp = torch.zeros((32, K), device=dz.device, dtype=torch.float32)
    L.call('gemm_wgrad', L.ptr(pad), L.ptr(x), L.ptr(tmp), L.ll(P), i32(K), i32(32), L.ll(32),
           L.ll(K), L.ll(K), st)
    L.call('copy2d', L.ptr(tmp), L.ll(K), L.ptr(gw), L.ll(K), L.ll(Cout), i32(K), i32(1), i32(0), st)


def linear(x, spec, triggers=()):
    if x.shape[0] == 0:
        return x.new_zeros((0, spec.wref.w.shape[0]))
    return _LinearFn.apply(x, spec, *triggers)


# ----------------------------------------------------------------------------------- deconv
class _Deconv2x2Fn(Function):
    """ConvTranspose2d(k=2, s=2) + bias + ReLU as one GEMM [P,Cin]x[Cin,4*Co] whose epilogue
    scatters (i,j) sub-pixels (fcn_mask_head.py:77-83,121-124).  spec.wref.w is the packed
    [(i,j,co), ci] weight, spec.bias the bias tiled 4x."""

    @staticmethod
    def forward(ctx, x, spec, *triggers):
        w = spec.wref.w
        N, Cin, H, W = x.shape
        Co = w.shape[0] // 4
        xn = nhwc(x)
        y = new_nhwc(N, Co, 2 * H, 2 * W, x.device)
        e = L.make_epilogue(shift=spec.bias, relu=spec.relu, deconv_shuffle=True,
                            round_out=spec.round_out)
        L.call('gemm_fprop', L.ptr(xn), L.ptr(w), L.ptr(y.permute(0, 2, 3, 1)), L.ll(N * H * W),
               i32(Cin), i32(4 * Co), L.ll(Cin), L.ll(Cin), L.ll(Co), i32(H), i32(W),
               ctypes.byref(e), L.stream())
        ctx.spec = spec
        ctx.save_for_backward(xn, y if (spec.relu and not spec.grad_premasked) else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        xn, y = ctx.saved_tensors
        _queue_finalize(spec.store)
        st = L.stream()
        N, H, W, Cin = xn.shape
        w = spec.wref.w
        Co = w.shape[0] // 4
        dyn = nhwc(dy)
        if y is None:
            dz = dyn
            if spec.bias_grad is not None and not spec.bias_by_consumer:
                L.call('act_bwd', L.ptr(dyn), None, None, None, None, None, None, None, None,
                       L.ptr(spec.bias_grad), L.ll(N * 4 * H * W), i32(Co), i32(0), st)
        else:
            dz = torch.empty_like(dyn)
            L.call('act_bwd', L.ptr(dyn), L.ptr(nhwc(y)), None, None, None, None, L.ptr(dz), None,
                   None, L.ptr(spec.bias_grad) if spec.bias_grad is not None else None,
                   L.ll(N * 4 * H * W), i32(Co), i32(1), st)
        # space-to-depth: dz[n,2h+i,2w+j,co] -> dzp[(n,h,w), (i,j,co)]
        dzp = dz.view(N, H, 2, W, 2, Co).permute(0, 1, 3, 2, 4, 5).contiguous().view(N * H * W,
                                                                                     4 * Co)
        P = N * H * W
        if spec.wref.grad is not None:
            L.call('gemm_wgrad', L.ptr(dzp), L.ptr(xn), L.ptr(spec.wref.grad), L.ll(P), i32(Cin),
                   i32(4 * Co), L.ll(4 * Co), L.ll(Cin), L.ll(Cin), st)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = new_nhwc(N, Cin, H, W, dy.device)
            e = L.make_epilogue(round_out=True, mask=xn if spec.premask_in else None,
                                colsum=spec.in_colsum)
            L.call('gemm_dgrad', L.ptr(dzp), L.ptr(w), L.ptr(dx.permute(0, 2, 3, 1)), L.ll(P),
                   i32(Cin), i32(4 * Co), L.ll(4 * Co), L.ll(Cin), L.ll(Cin), ctypes.byref(e), st)
        return (dx, None) + (None,) * (len(ctx.needs_input_grad) - 2)


class _DeconvLogitsFn(Function):
    """ConvTranspose2d(k=2, s=2) + bias + ReLU followed by a narrow 1x1 head (the mask head's
    upsample + conv_logits, fcn_mask_head.py:77-83,118-126) with the deconv output kept in the
    UN-shuffled GEMM layout [(n, h, w), (i, j, co)]: the head is per pixel, so it reads those rows
    as they are and only its 4-wide output / gradient rows are addressed in image order
    (`s2d_*` of loft_narrow_head_*).  The backward's dz then already has the layout the deconv's
    weight / data gradient GEMMs take -- the 160 MB space-to-depth copy of `_Deconv2x2Fn.backward`
    (0.11 ms) and the pixel-shuffle store of its forward are gone."""

    @staticmethod
    def forward(ctx, x, up, lg, *triggers):
        w = up.wref.w
        N, Cin, H, W = x.shape
        Co = w.shape[0] // 4
        P = N * H * W
        xn = nhwc(x)
        yp = torch.empty((P, 4 * Co), device=x.device, dtype=torch.float32)
        e = L.make_epilogue(shift=up.bias, relu=up.relu, round_out=up.round_out)
        L.call('gemm_fprop', L.ptr(xn), L.ptr(w), L.ptr(yp), L.ll(P), i32(Cin), i32(4 * Co),
               L.ll(Cin), L.ll(Cin), L.ll(4 * Co), i32(1), i32(P), ctypes.byref(e), L.stream())
        out = new_nhwc(N, lg.wref.w.shape[0], 2 * H, 2 * W, x.device)
        L.call('narrow_head_fwd', L.ptr(yp), L.ptr(lg.wref.w), L.ptr(lg.bias),
               L.ptr(out.permute(0, 2, 3, 1)), L.ll(4 * P), i32(Co), i32(lg.n_out or 4), i32(H),
               i32(W), L.stream())
        ctx.specs = (up, lg)
        ctx.save_for_backward(xn, yp)
        return out

    @staticmethod
    def backward(ctx, dy):
        up, lg = ctx.specs
        xn, yp = ctx.saved_tensors
        _queue_finalize(up.store)
        st = L.stream()
        N, H, W, Cin = xn.shape
        P = N * H * W
        w = up.wref.w
        Co = w.shape[0] // 4
        dzp = torch.empty_like(yp)
        # dz of the deconv (ReLU mask of its output applied, rounded), the head's weight / bias
        # gradients and the deconv's bias gradient (per-channel sum of dz) in one pass over yp
        L.call('narrow_head_bwd', L.ptr(nhwc(dy)), L.ptr(yp), L.ptr(lg.wref.w), L.ptr(dzp),
               L.ptr(lg.wref.grad), L.ptr(lg.bias_grad), L.ptr(up.bias_grad), L.ll(4 * P), i32(Co),
               i32(lg.n_out or 4), i32(1 if up.relu else 0), i32(H), i32(W), st)
        if up.wref.grad is not None:
            L.call('gemm_wgrad', L.ptr(dzp), L.ptr(xn), L.ptr(up.wref.grad), L.ll(P), i32(Cin),
                   i32(4 * Co), L.ll(4 * Co), L.ll(Cin), L.ll(Cin), st)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = new_nhwc(N, Cin, H, W, dy.device)
            e = L.make_epilogue(round_out=True, mask=xn if up.premask_in else None,
                                colsum=up.in_colsum)
            L.call('gemm_dgrad', L.ptr(dzp), L.ptr(w), L.ptr(dx.permute(0, 2, 3, 1)), L.ll(P),
                   i32(Cin), i32(4 * Co), L.ll(4 * Co), L.ll(Cin), L.ll(Cin), ctypes.byref(e), st)
        return (dx, None, None) + (None,) * (len(ctx.needs_input_grad) - 3)


def deconv_logits_ok(up, lg, x):
    """`deconv2x2(x, up)` followed by `conv(., lg)` can run as `_DeconvLogitsFn`."""
    Co = up.wref.w.shape[0] // 4
    return (os.environ.get('LOFT_NARROW_HEAD', '1') != '0' and
            os.environ.get('LOFT_DECONV_LOGITS', '1') != '0' and lg.ksize == 1 and lg.stride == 1 and
            lg.padding == 0 and not lg.relu and lg.bn is None and not lg.round_out and
            lg.wref.w.dim() == 2 and lg.wref.w.shape[0] == 4 and lg.wref.w.shape[1] == Co and
            Co in (128, 256, 512) and up.bn is None and lg.bias_grad is not None and
            up.bias_grad is not None and x.dim() == 4)


def deconv_logits(x, up, lg, triggers=()):
    if x.shape[0] == 0:
        N, _, H, W = x.shape
        return x.new_zeros((0, lg.wref.w.shape[0], 2 * H, 2 * W))
    return _DeconvLogitsFn.apply(x, up, lg, *triggers)


def deconv2x2(x, spec, triggers=()):
    if x.shape[0] == 0:
        N, _, H, W = x.shape
        return x.new_zeros((0, spec.wref.w.shape[0] // 4, 2 * H, 2 * W))
    return _Deconv2x2Fn.apply(x, spec, *triggers)


# ----------------------------------------------------------------------------------- grouped 3x3
class GroupedConvSpec:
    """G convs (3x3/s1/p1 + bias + ReLU) with separate weights applied to G consecutive groups of
    images in ONE launch.  The per-group weights / biases / gradients must be equally strided in
    memory (true for the FOA branches in the ParamStore's flat buffers)."""

    def __init__(self, wrefs, biases, bias_grads, relu=True, store=None, premask_in=False,
                 grad_premasked=False):
        self.premask_in, self.grad_premasked = premask_in, grad_premasked
        self.in_colsum, self.in_colsum_gstride, self.bias_by_consumer = None, 0, False
        self.G = len(wrefs)
        self.w0 = wrefs[0].w
        self.gw0 = wrefs[0].grad
        self.b0 = biases[0]
        self.bias_grads = bias_grads
        self.relu = relu
        self.store = store
        self.cout, self.cin = wrefs[0].w.shape[0], wrefs[0].w.shape[1]

        def stride(ts):
            d = [(ts[i + 1].data_ptr() - ts[i].data_ptr()) for i in range(len(ts) - 1)]
            ok = len(set(d)) <= 1 and (not d or (d[0] > 0 and d[0] % 16 == 0))
            return (d[0] // 4 if d else 0), ok
        self.w_gstride, ok1 = stride([r.w for r in wrefs])
        self.b_gstride, ok2 = stride(list(biases))
        ok3 = True
        self.gw_gstride = 0
        if self.gw0 is not None:
            self.gw_gstride, ok3 = stride([r.grad for r in wrefs])
        self.b_grad_gstride, ok4 = (0, True)
        if all(g is not None for g in bias_grads):
            self.b_grad_gstride, ok4 = stride(list(bias_grads))
        self.uniform = ok1 and ok2 and ok3 and ok4


class _GroupedConv3x3Fn(Function):
    @staticmethod
    def forward(ctx, x, spec, *triggers):
        N, Cin, H, W = x.shape
        xn = nhwc(x)
        y = new_nhwc(N, spec.cout, H, W, x.device)
        e = L.make_epilogue(shift=spec.b0, relu=spec.relu, round_out=True)
        L.call('conv3x3_fprop_grouped', L.ptr(xn), L.ptr(spec.w0), L.ptr(y.permute(0, 2, 3, 1)),
               i32(N), i32(H), i32(W), i32(Cin), i32(spec.cout), i32(spec.G), L.ll(spec.w_gstride),
               L.ll(spec.b_gstride), ctypes.byref(e), L.stream())
        ctx.spec = spec
        ctx.save_for_backward(xn, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        xn, y = ctx.saved_tensors
        _queue_finalize(spec.store)
        st = L.stream()
        N, H, W, Cin = xn.shape
        Cout, G = spec.cout, spec.G
        dyn = nhwc(dy)
        yn = nhwc(y)
        relu_eff = spec.relu and not spec.grad_premasked
        dz = torch.empty_like(dyn) if relu_eff else dyn
        rows_g = (N // G) * H * W
        for g in range(G):
            o = g * rows_g * Cout * 4
            if relu_eff:
                L.call('act_bwd', ctypes.c_void_p(dyn.data_ptr() + o),
                       ctypes.c_void_p(yn.data_ptr() + o), None, None, None, None,
                       ctypes.c_void_p(dz.data_ptr() + o), None, None,
                       L.ptr(spec.bias_grads[g]) if spec.bias_grads[g] is not None else None,
                       L.ll(rows_g), i32(Cout), i32(1), st)
            elif spec.bias_grads[g] is not None and not spec.bias_by_consumer:
                L.call('act_bwd', ctypes.c_void_p(dyn.data_ptr() + o), None, None, None, None, None,
                       None, None, None, L.ptr(spec.bias_grads[g]), L.ll(rows_g), i32(Cout), i32(0),
                       st)
        if spec.gw0 is not None:
            with _maybe_side_stream(spec.store, ctx.needs_input_grad[0], dz, xn):
                L.call('conv3x3_wgrad_grouped', L.ptr(dz), L.ptr(xn), L.ptr(spec.gw0), i32(N),
                       i32(H), i32(W), i32(Cin), i32(Cout), i32(G), L.ll(spec.gw_gstride),
                       L.stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = new_nhwc(N, Cin, H, W, dy.device)
            e = L.make_epilogue(round_out=True, mask=xn if spec.premask_in else None,
                                colsum=spec.in_colsum, colsum_gstride=spec.in_colsum_gstride)
            L.call('conv3x3_dgrad_grouped', L.ptr(dz), L.ptr(spec.w0), L.ptr(dx.permute(0, 2, 3, 1)),
                   i32(N), i32(H), i32(W), i32(Cin), i32(Cout), i32(G), L.ll(spec.w_gstride),
                   ctypes.byref(e), st)
        return (dx, None) + (None,) * (len(ctx.needs_input_grad) - 2)


def grouped_conv3x3(x, spec, triggers=()):
    return _GroupedConv3x3Fn.apply(x, spec, *triggers)
