"""Device-resident parameter store of the LOFT training path.

All floating-point parameters of a model are moved into ONE flat fp32 buffer (trainable ones
first) with twin flat buffers for gradients, SGD momentum and the TF32-rounded copy the tensor
cores read.  ``nn.Parameter.data`` / ``.grad`` become views, so ``state_dict()`` keeps the
reference's names and shapes while

* weight-gradient kernels accumulate straight into the flat gradient buffer (no autograd
  accumulation pass, nothing to flatten before the NCCL all-reduce -- replaces the torch DDP
  reducer wrapped at mmdet/apis/train.py:71-79),
* grad-norm clipping + SGD(momentum, weight decay) is two launches over the flat buffers
  (replaces mmcv OptimizerHook + torch.optim.SGD, configs/_base_/schedules/schedule_2x_bonai.py:2-3),
* 4-D conv weights live in channels_last order ([Cout][kh][kw][Cin]) which is exactly the K-major
  operand layout of the implicit-GEMM kernels.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib as L


class WeightRef:
    """Kernel-facing handles of one weight tensor: TF32 copy + gradient accumulation target."""

    def __init__(self, w, grad):
        self.w = w          # tensor the GEMM reads (TF32-rounded, kernel layout)
        self.grad = grad    # tensor wgrad accumulates into (None if frozen)


class BNRef:
    def __init__(self, scale, shift, rstd, mean, dgamma, dbeta):
        self.scale, self.shift, self.rstd, self.mean = scale, shift, rstd, mean
        self.dgamma, self.dbeta = dgamma, dbeta


class Packed:
    """A kernel weight that is a re-layout / fusion of one or more parameters (fc weights whose K
    axis must follow NHWC order, the 2x2 deconv, heads fused or padded to a TMA-legal width).
    ``build`` refreshes ``w``/``b`` from the master parameters; ``scatter`` adds the temp
    gradients back into the parameters' gradient views at the end of backward."""

    def __init__(self, w, b, gw, gb, build, scatter):
        self.w, self.b, self.gw, self.gb = w, b, gw, gb
        self.build, self.scatter = build, scatter


def _is_bn(m):
    return isinstance(m, nn.modules.batchnorm._BatchNorm)


class ParamStore:
    def __init__(self, model, device):
        self.device = torch.device(device)
        self.model = model
        bns = [m for m in model.modules() if _is_bn(m)]
        bn_param_ids = set()
        for m in bns:
            bn_param_ids.add(id(m.weight))
            bn_param_ids.add(id(m.bias))
        tr_bn = [m for m in bns if m.weight.requires_grad]
        fz_bn = [m for m in bns if not m.weight.requires_grad]
        others = [p for p in model.parameters() if id(p) not in bn_param_ids]
        tr_other = [p for p in others if p.requires_grad]
        fz_other = [p for p in others if not p.requires_grad]
        order = ([m.weight for m in tr_bn] + [m.bias for m in tr_bn] + tr_other +
                 [m.weight for m in fz_bn] + [m.bias for m in fz_bn] + fz_other)
        # 16-byte aligned offsets (TMA base addresses)
        offs, off = [], 0
        n_train_params = 2 * len(tr_bn) + len(tr_other)
        self.n_train = None
        for i, p in enumerate(order):
            if i == n_train_params:
                self.n_train = off
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        if self.n_train is None:
            self.n_train = off
        self.total = off
        dev = self.device
        self.P = torch.zeros(off, device=dev)
        self.T = torch.zeros(off, device=dev)
        self.G = torch.zeros(self.n_train, device=dev)
        self.M = torch.zeros(self.n_train, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float64, device=dev)
        self._grad_views = []
        for p, o in zip(order, offs):
            n = p.numel()
            src = p.data
            if p.dim() == 4 and (p.shape[2] > 1 or p.shape[3] > 1):
                a, b, c, d = p.shape
                view = self.P[o:o + n].view(a, c, d, b).permute(0, 3, 1, 2)
                tview = self.T[o:o + n].view(a, c, d, b).permute(0, 3, 1, 2)
                gview = self.G[o:o + n].view(a, c, d, b).permute(0, 3, 1, 2) \
                    if o < self.n_train else None
            else:
                view = self.P[o:o + n].view(p.shape)
                tview = self.T[o:o + n].view(p.shape)
                gview = self.G[o:o + n].view(p.shape) if o < self.n_train else None
            view.copy_(src)
            p.data = view
            p.grad = gview
            p._loft = WeightRef(tview, gview)
            self._grad_views.append((p, gview))
            if o < self.n_train:
                p._loft_momentum = self.M[o:o + n]
        # BN statistics in two contiguous runs (trainable BNs, frozen BNs)
        self._bn_groups = []
        for group in (tr_bn, fz_bn):
            C = sum(m.num_features for m in group)
            if C == 0:
                self._bn_groups.append(None)
                continue
            mean = torch.zeros(C, device=dev)
            var = torch.ones(C, device=dev)
            scale = torch.zeros(C, device=dev)
            shift = torch.zeros(C, device=dev)
            rstd = torch.zeros(C, device=dev)
            g0 = group[0].weight._loft
            c0 = 0
            # gammas of the group are contiguous in P (then betas): find their flat offsets
            goff = group[0].weight.data.data_ptr()
            boff = group[0].bias.data.data_ptr()
            for m in group:
                c = m.num_features
                assert c % 4 == 0, 'BN channel counts must be multiples of 4'
                mean[c0:c0 + c].copy_(m.running_mean)
                var[c0:c0 + c].copy_(m.running_var)
                m.running_mean = mean[c0:c0 + c]
                m.running_var = var[c0:c0 + c]
                if m.num_batches_tracked is not None:
                    m.num_batches_tracked = m.num_batches_tracked.to(dev)
                m._loft_bn = BNRef(scale[c0:c0 + c], shift[c0:c0 + c], rstd[c0:c0 + c],
                                   mean[c0:c0 + c], m.weight._loft.grad, m.bias._loft.grad)
                m._loft_eps = m.eps
                c0 += c
            eps = group[0].eps
            assert all(m.eps == eps for m in group)
            self._bn_groups.append(dict(C=C, gamma_ptr=goff, beta_ptr=boff, mean=mean, var=var,
                                        scale=scale, shift=shift, rstd=rstd, eps=eps))
        # move remaining (non-BN) buffers
        for m in model.modules():
            if _is_bn(m):
                continue
            for k, b in list(m._buffers.items()):
                if b is not None:
                    m._buffers[k] = b.to(dev)
        self.packed = []
        self.pre_finalize = []           # callables run at the start of finalize_grads
        self._fold_pairs = []
        self._owner = ''
        for name, m in model.named_modules():
            if hasattr(m, 'loft_prepare'):
                self._owner = name            # add_packed tags what the module registers
                m.loft_prepare(self)
        # The RoI heads' parameters are the tail of the trainable range (model.parameters()
        # order): their gradients are final when autograd reaches the trunk's backward, so a
        # data-parallel trainer can exchange them under it (Trainer, trunk._TrunkFn.backward).
        self.head_start = self.n_train
        names = {id(p): n for n, p in model.named_parameters()}
        tail = True
        for p, o in reversed(list(zip(order, offs))):
            if o >= self.n_train:
                continue
            if tail and names.get(id(p), '').startswith('roi_head.') and id(p) not in bn_param_ids:
                self.head_start = o
            else:
                tail = False
        self.heads_done_hooks = []           # called at the start of the trunk's backward
        # Second cut for an overlapped gradient exchange: the parameters of backbone.layer3 /
        # layer4, the neck and the RPN head are a contiguous range [mid_start, head_start) of the
        # flat buffer whose gradients are final once the trunk's backward has passed layer3 -- the
        # trunk then calls upper_done() between its two backward programs.  None if the layout
        # does not have that shape (other backbones).
        upper = ('backbone.layer3.', 'backbone.layer4.', 'neck.', 'rpn_head.')
        self.mid_start = None
        lo = None
        ok = True
        for p, o in zip(order, offs):
            if o >= self.head_start:
                continue
            nm = names.get(id(p), '')
            if nm.startswith(upper):
                lo = o if lo is None else min(lo, o)
        if lo is not None:
            for p, o in zip(order, offs):
                nm = names.get(id(p), '')
                if lo <= o < self.head_start and not nm.startswith(upper):
                    ok = False
                if o < lo and nm.startswith(upper):
                    ok = False
            if ok and 0 < lo < self.head_start:
                self.mid_start = lo
        self.upper_done_hooks = []
        self._build_fold_tables()
        # everything whose in-place modification (load_state_dict, init_weights, an external torch
        # optimizer) must trigger a rebuild of the derived copies: the parameters themselves --
        # `p.data = view` keeps each parameter's OWN version counter, self.P._version never moves
        # when a parameter is written through -- and the BN running statistics
        self._versioned = [p for p, _ in self._grad_views]
        for m in bns:
            self._versioned += [m.running_mean, m.running_var]
        self._dirty = False
        self._seen_version = None
        self._frozen_ready = False
        self._callback_queued = False
        # optional (LOFT_WSTREAM=1): weight-gradient kernels of small layers on a side stream so
        # they overlap the data-gradient chain.  Measured on B200: 23.97 ms/step with vs 23.56
        # without -- the 200 KB-smem GEMM CTAs cannot co-reside, so it stays off by default.
        import os
        self.wstream = torch.cuda.Stream(device=dev) if os.environ.get('LOFT_WSTREAM', '0') != '0' \
            else None
        self._w_pending = False
        model._loft_store = self

    # ------------------------------------------------------------------ BN folded into weights
    def fold_bn(self, conv_weight, bn):
        """Declare that eval-mode `bn` directly follows the conv owning `conv_weight`: the TF32
        copy the tensor cores read becomes scale*W (loft_bn_fold_weights) and gamma's gradient is
        recovered from the weight gradient after backward (loft_bn_finalize)."""
        self._fold_pairs.append((conv_weight, bn))

    def _build_fold_tables(self):
        """Row tables (flat offset, K, BN channel) per BN group: [0] trainable, [1] frozen."""
        dev = self.device
        self._fold_tables = [None, None]
        base = self.P.data_ptr()
        rows = [([], [], []), ([], [], [])]
        for w, bn in self._fold_pairs:
            gi = 0 if bn.weight.requires_grad else 1
            g = self._bn_groups[gi]
            assert w.requires_grad == bn.weight.requires_grad, \
                'a conv and the BN folded into it must be frozen / trainable together'
            cout = w.shape[0]
            K = w.numel() // cout
            off = (w.data.data_ptr() - base) // 4
            ch0 = (bn._loft_bn.scale.data_ptr() - g['scale'].data_ptr()) // 4
            ro, rk, rc = rows[gi]
            ro.extend(off + c * K for c in range(cout))
            rk.extend([K] * cout)
            rc.extend(ch0 + c for c in range(cout))
        for gi, (ro, rk, rc) in enumerate(rows):
            if ro:
                self._fold_tables[gi] = dict(
                    off=torch.tensor(ro, dtype=torch.int64, device=dev),
                    k=torch.tensor(rk, dtype=torch.int32, device=dev),
                    ch=torch.tensor(rc, dtype=torch.int32, device=dev), rows=len(ro))

    def _fold_weights(self, gi):
        t, g = self._fold_tables[gi], self._bn_groups[gi]
        if t is None:
            return
        L.call('bn_fold_weights', L.ptr(self.P), L.ptr(self.T), L.ptr(t['off']), L.ptr(t['k']),
               L.ptr(t['ch']), L.ptr(g['scale']), L.ll(t['rows']), L.stream())

    def _bn_finalize(self):
        t, g = self._fold_tables[0], self._bn_groups[0]
        if t is None:
            return
        C = g['C']
        L.call('bn_finalize', L.ptr(self.P), L.ptr(self.G), L.ptr(t['off']), L.ptr(t['k']),
               L.ptr(t['ch']), L.ptr(g['scale']), L.ptr(g['rstd']), L.ptr(g['mean']),
               L.ptr(self.G[C:2 * C]), L.ptr(self.G[0:C]), L.ll(t['rows']), L.stream())

    # ------------------------------------------------------------------ packed weights
    def add_packed(self, packed):
        packed.owner = self._owner
        packed.scattered = False
        self.packed.append(packed)
        return packed

    def heads_done(self):
        """The RoI heads' backward has been issued (called by the trunk's backward node): fold
        their packed-weight gradients into the parameter gradients now and tell the listeners."""
        if not self.heads_done_hooks:
            return
        with L.batched_copies():
            for pk in self.packed:
                if pk.owner.startswith('roi_head') and not pk.scattered:
                    pk.scatter()
                    pk.scattered = True
        for fn in self.heads_done_hooks:
            fn()

    def upper_done(self):
        """The trunk's backward has produced every gradient of [mid_start, head_start)."""
        for fn in self.upper_done_hooks:
            fn()

    # ------------------------------------------------------------------ per-step protocol
    def _version_sig(self):
        return self.P._version + sum(t._version for t in self._versioned)

    def mark_dirty(self):
        """Tell the store that master weights / BN statistics were modified behind autograd's
        back (through `.data`, a raw pointer, a custom kernel): the next step rebuilds the TF32
        copy, the folded-BN weights and the packed weights.  In-place tensor ops on the
        parameters (load_state_dict, optimizer.step of a torch optimizer, init_weights) are
        detected without this."""
        self._dirty = True

    def refresh_weights(self, force=False):
        """TF32-round the master weights into T, fold BN, rebuild packed weights (only when a
        parameter or a BN statistic changed since the last call)."""
        ver = self._version_sig()
        if not force and not self._dirty and self._seen_version == ver:
            return
        self._dirty = False
        end = self.total
        L.call('copy2d', L.ptr(self.P), L.ll(end), L.ptr(self.T), L.ll(end), L.ll(1),
               ctypes.c_int(end), ctypes.c_int(0), ctypes.c_int(1), L.stream())
        self._frozen_ready = False          # the full copy overwrote the folded frozen weights
        self._after_weight_update()
        self._seen_version = ver

    def _after_weight_update(self):
        for gi, g in enumerate(self._bn_groups):
            if g is None or (gi == 1 and self._frozen_ready):
                continue
            L.call('bn_fold', ctypes.c_void_p(g['gamma_ptr']), ctypes.c_void_p(g['beta_ptr']),
                   L.ptr(g['mean']), L.ptr(g['var']), L.f32(g['eps']), L.ptr(g['scale']),
                   L.ptr(g['shift']), L.ptr(g['rstd']), ctypes.c_int(g['C']), L.stream())
            self._fold_weights(gi)
        self._frozen_ready = True
        with L.batched_copies():            # the builds' small copies: one launch
            for pk in self.packed:
                pk.build()

    def begin_step(self):
        """Zero the gradient buffers (weight-gradient kernels accumulate atomically)."""
        self.refresh_weights()
        self.G.zero_()
        self.sqnorm.zero_()
        if not hasattr(self, '_packed_grads'):
            self._packed_grads = [t for pk in self.packed for t in (pk.gw, pk.gb) if t is not None]
        if self._packed_grads:
            torch._foreach_zero_(self._packed_grads)
        for pk in self.packed:
            pk.scattered = False
        self._callback_queued = False

    def queue_finalize(self):
        """Called from inside a backward node: run finalize_grads once autograd is done."""
        if self._callback_queued:
            return
        self._callback_queued = True
        torch.autograd.Variable._execution_engine.queue_callback(self.finalize_grads)

    def side_stream_for_wgrad(self, *tensors):
        """Returns the side stream (after making it wait for the current stream and registering
        the operand tensors with it) or None when disabled."""
        ws = self.wstream
        if ws is None:
            return None
        ws.wait_stream(torch.cuda.current_stream(self.device))
        for t in tensors:
            if t is not None:
                t.record_stream(ws)
        self._w_pending = True
        return ws

    def join_wgrad_stream(self):
        if self._w_pending and self.wstream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.wstream)
            self._w_pending = False

    def finalize_grads(self):
        for fn in self.pre_finalize:
            fn()
        self.join_wgrad_stream()
        # packed gradients first: a packed weight may belong to a BN-folded conv (HRNet stem),
        # whose gamma gradient bn_finalize derives from the un-packed weight gradient
        with L.batched_copies():
            for pk in self.packed:
                if not pk.scattered:
                    pk.scatter()
                    pk.scattered = True
        self._bn_finalize()
        for p, g in self._grad_views:
            if g is not None and p.grad is not g:
                p.grad = g

    # ------------------------------------------------------------------ optimizer
    def sgd_step(self, lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0, grad_scale=1.0):
        """clip_grad_norm_(max_norm) + SGD on the flat buffers; refreshes T in the same pass."""
        n = self.n_train
        use_clip = max_norm is not None and max_norm > 0
        if use_clip:
            L.call('grad_sqnorm', L.ptr(self.G), L.ll(n), L.ptr(self.sqnorm), L.stream())
        L.call('sgd_clip_step', L.ptr(self.P), L.ptr(self.G), L.ptr(self.M), L.ptr(self.T), L.ll(n),
               L.f32(lr), L.f32(momentum), L.f32(weight_decay), L.f32(max_norm if use_clip else 0.0),
               L.f32(grad_scale), L.ptr(self.sqnorm) if use_clip else None, L.stream())
        self._after_weight_update()
        # the kernels write P / T through raw pointers: no version counter moved, the signature
        # taken in begin_step still describes what T was derived from

    def momentum_state(self):
        """{parameter name: momentum buffer (reference layout)} for checkpoints."""
        out = {}
        for name, p in self.model.named_parameters():
            m = getattr(p, '_loft_momentum', None)
            if m is not None:
                phys = m.view(p.permute(0, 2, 3, 1).shape).permute(0, 3, 1, 2) \
                    if (p.dim() == 4 and (p.shape[2] > 1 or p.shape[3] > 1)) else m.view(p.shape)
                out[name] = phys.detach().clone().contiguous()
        return out

    def load_momentum_state(self, state):
        for name, p in self.model.named_parameters():
            m = getattr(p, '_loft_momentum', None)
            if m is not None and name in state:
                src = state[name].to(self.device)
                dst = m.view(p.permute(0, 2, 3, 1).shape).permute(0, 3, 1, 2) \
                    if (p.dim() == 4 and (p.shape[2] > 1 or p.shape[3] > 1)) else m.view(p.shape)
                dst.copy_(src)

    def grad_norm(self):
        """Global L2 norm of the last step's gradients (device tensor, no sync)."""
        return self.sqnorm.sqrt().float()


def get_store(model, device=None):
    st = getattr(model, '_loft_store', None)
    if st is None:
        if not torch.cuda.is_available():
            raise L.LoftError('the LOFT hot path needs a CUDA device (sm_100a); there is no CPU '
                              'fallback')
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        st = ParamStore(model, device)
    return st
