"""The static part of the LOFT training step as two recorded launch programs.

Everything between the input tile and the RoI heads has shapes that depend only on (N, H, W):
the frozen stem + layer1, the trainable ResNet stages (mmdet/models/backbones/resnet.py:260-300,
623-638), FPN (necks/fpn.py:164-216) and the RPN convolutions (dense_heads/rpn_head.py:38-44).
Instead of ~150 autograd nodes that each allocate, wrap and launch, this module writes that part
as ONE autograd.Function whose forward and backward are fixed sequences of C-ABI launches over
buffers allocated once: the sequence is *recorded* the first time it runs (`Tape`), then replayed
as a CUDA graph -- one graph launch for the forward, two for the backward (RPN part, the rest).

Because the whole backward is one hand-ordered program, ReLU / BN / residual backward never
touches HBM on its own: every data-gradient launch applies, in its epilogue, the ReLU mask of
the tensor it differentiates, adds the gradient arriving over the other branch (identity path,
FPN lateral, RPN), and accumulates the per-channel sum that is the producer's BN-beta / bias
gradient; BN-gamma gradients come from the weight gradients (engine._bn_finalize).
"""
import ctypes
import os

import torch
from torch.autograd import Function

from . import _lib as L
from .ops.misc import stem_packed_shape

i32 = ctypes.c_int


class Tape:
    """Records C-ABI calls made while active; `launch()` re-issues them (same pointers, current
    stream) -- as a captured CUDA graph after `capture()`."""

    def __init__(self, name='tape'):
        self.name = name
        self.calls = []
        self.graph = None

    def __enter__(self):
        assert L.RECORD is None, 'nested tape recording'
        L.RECORD = self.calls
        return self

    def __exit__(self, *a):
        L.RECORD = None
        return False

    def replay(self):
        st = L.stream()
        for name, fn, args in self.calls:
            rc = fn(*args[:-1], st)
            if rc != 0:
                msg = L.lib().loft_last_error()
                raise L.LoftError(f'loft_{name} failed ({rc}) in tape replay: '
                                  f'{msg.decode() if msg else ""}')

    def capture(self):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode='thread_local'):
            self.replay()
        self.graph = g

    def launch(self):
        L.LAUNCHES[0] += len(self.calls)
        if L.TRACE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.graph.replay() if self.graph is not None else self.replay()
            e1.record()
            L.TRACE.append((self.name, (len(self.calls),), e0, e1))
            return
        if self.graph is not None:
            self.graph.replay()
        else:
            self.replay()


def _empty(*shape, dev):
    return torch.empty(shape, device=dev, dtype=torch.float32)


def _epi(**kw):
    return ctypes.byref(L.make_epilogue(**kw))


class _Program:
    """Builds (by executing once under a Tape) the forward / backward programs for one input
    shape.  All tensors are NHWC; `self.keep` pins every buffer the recorded pointers refer to."""

    def __init__(self, model, N, H, W, dev, train):
        self.model, self.N, self.H, self.W, self.dev, self.train = model, N, H, W, dev, train
        self.img = torch.empty((N, 3, H, W), device=dev, dtype=torch.float32)
        self.fwd, self.bwd = Tape('trunk_forward_graph'), Tape('trunk_backward_graph')
        # the tail of the main backward (the stages below `split_stage`), recorded separately when
        # a data-parallel trainer wants to exchange the upper stages' gradients under it
        self.bwd_tail = Tape('trunk_backward_tail_graph')
        self.split_stage = None
        self.bwd_rpn = Tape('trunk_rpn_backward_graph')
        self.fwd_ready = self.bwd_ready = self.bwd_rpn_ready = False
        self.rpn_done, self.rpn_event = False, None
        self.use_graph = os.environ.get('LOFT_GRAPH', '1') != '0'
        self.keep = []

    # ---------------------------------------------------------------- primitive launches
    def gemm_f(self, x, w, y, P, K, Co, Ho, Wo, **epi):
        L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co), L.ll(K),
               L.ll(K), L.ll(y.shape[-1]), i32(Ho), i32(Wo), _epi(**epi), L.stream())

    def conv3_f(self, x, w, y, **epi):
        N, H, W, Ci = x.shape
        L.call('conv3x3_fprop', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H), i32(W), i32(Ci),
               i32(y.shape[-1]), _epi(**epi), L.stream())

    def gemm_d(self, dy, w, dx, P, Cin, Cout, H=0, W=0, **epi):
        L.call('gemm_dgrad_hw', L.ptr(dy), L.ptr(w), L.ptr(dx), L.ll(P), i32(Cin), i32(Cout),
               L.ll(Cout), L.ll(Cin), L.ll(Cin), i32(H), i32(W), _epi(**epi), L.stream())

    def gemm_w(self, dy, x, gw, P, Cin, Cout):
        L.call('gemm_wgrad', L.ptr(dy), L.ptr(x), L.ptr(gw), L.ll(P), i32(Cin), i32(Cout),
               L.ll(Cout), L.ll(Cin), L.ll(Cin), L.stream())

    def conv3_d(self, dy, w, dx, **epi):
        N, H, W, Ci = dx.shape
        L.call('conv3x3_dgrad', L.ptr(dy), L.ptr(w), L.ptr(dx), i32(N), i32(H), i32(W), i32(Ci),
               i32(dy.shape[-1]), _epi(**epi), L.stream())

    def conv3_w(self, dy, x, gw):
        N, H, W, Ci = x.shape
        L.call('conv3x3_wgrad', L.ptr(dy), L.ptr(x), L.ptr(gw), i32(N), i32(H), i32(W), i32(Ci),
               i32(dy.shape[-1]), L.stream())

    def colsum(self, g, P, C, out):
        """out[c] += sum_p g[p,c] (read-only pass; used where no dgrad epilogue can do it)."""
        L.call('act_bwd', L.ptr(g), None, None, None, None, None, None, None, None, L.ptr(out),
               L.ll(P), i32(C), i32(0), L.stream())

    def buf(self, *shape):
        t = _empty(*shape, dev=self.dev)
        self.keep.append(t)
        return t

    # ---------------------------------------------------------------- forward program
    def _block_fwd(self, x, blk):
        s1, s2, s3, sd = blk._s1, blk._s2, blk._s3, blk._sd
        N, H, W, Ci = x.shape
        p = s1.wref.w.shape[0]
        P1 = N * H * W
        h1 = self.buf(N, H, W, p)
        self.gemm_f(x, s1.wref.w, h1, P1, Ci, p, H, W, shift=s1.bias, relu=True, round_out=True)
        st = s2.stride
        col = None
        if st == 1:
            Ho, Wo = H, W
            h2 = self.buf(N, H, W, p)
            self.conv3_f(h1, s2.wref.w, h2, shift=s2.bias, relu=True, round_out=True)
        else:
            Ho, Wo = (H + 2 - 3) // st + 1, (W + 2 - 3) // st + 1
            col = self.buf(N * Ho * Wo, 9 * p)
            L.call('im2col', L.ptr(h1), L.ptr(col), i32(N), i32(H), i32(W), i32(p), i32(3), i32(3),
                   i32(st), i32(1), i32(9 * p), i32(0), L.stream())
            h2 = self.buf(N, Ho, Wo, p)
            self.gemm_f(col, s2.wref.w, h2, N * Ho * Wo, 9 * p, p, Ho, Wo, shift=s2.bias,
                        relu=True, round_out=True)
        P3 = N * Ho * Wo
        xs = x
        if sd is not None:
            if sd.stride == 2:
                xs = self.buf(N, Ho, Wo, Ci)
                L.call('subsample2', L.ptr(x), L.ptr(xs), i32(N), i32(H), i32(W), i32(Ci),
                       L.stream())
            else:
                assert sd.stride == 1
            idn = self.buf(N, Ho, Wo, 4 * p)
            self.gemm_f(xs, sd.wref.w, idn, P3, Ci, 4 * p, Ho, Wo, shift=sd.bias, round_out=True)
        else:
            idn = x
        out = self.buf(N, Ho, Wo, 4 * p)
        self.gemm_f(h2, s3.wref.w, out, P3, p, 4 * p, Ho, Wo, shift=s3.bias, residual=idn,
                    ldr=4 * p, relu=True, round_out=True)
        return out, dict(blk=blk, x=x, xs=xs, h1=h1, h2=h2, col=col, out=out, stride=st)

    def build_forward(self):
        m = self.model
        bb, neck, rpn = m.backbone, m.neck, m.rpn_head
        N, H, W = self.N, self.H, self.W
        with self.fwd:
            # frozen stem: pack the NCHW image (padded NHWC4) + direct 7x7/2 conv GEMM (+folded BN,
            # ReLU) over its sliding windows + 3x3/2 max-pool
            Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            xp = self.buf(*stem_packed_shape(N, H, W))
            L.call('stem_pack', L.ptr(self.img), L.ptr(xp), i32(N), i32(H), i32(W), i32(3), L.stream())
            c1 = self.buf(N, Ho, Wo, bb._stem_w.shape[0])
            L.call('stem_conv7x7', L.ptr(xp), L.ptr(bb._stem_w), L.ptr(c1), i32(N), i32(H), i32(W),
                   i32(c1.shape[-1]),
                   _epi(shift=bb.bn1._loft_bn.shift, relu=True, round_out=True), L.stream())
            Hp, Wp = (Ho - 1) // 2 + 1, (Wo - 1) // 2 + 1
            x = self.buf(N, Hp, Wp, c1.shape[-1])
            L.call('maxpool3x3s2', L.ptr(c1), L.ptr(x), i32(N), i32(Ho), i32(Wo),
                   i32(c1.shape[-1]), L.stream())
            self.stages = []
            C = []
            for name in bb.res_layers:
                recs = []
                for blk in getattr(bb, name):
                    x, rec = self._block_fwd(x, blk)
                    recs.append(rec)
                self.stages.append(recs)
                C.append(x)
            self.C = C
            # FPN: laterals top-down (the upsample-add rides in the lateral's epilogue), 3x3 outs
            n = len(neck._lat)
            lat = [None] * n
            for i in range(n - 1, -1, -1):
                sp = neck._lat[i]
                Ni, Hi, Wi, Ci = C[i].shape
                lat[i] = self.buf(Ni, Hi, Wi, neck.out_channels)
                res = lat[i + 1] if i < n - 1 else None
                if res is not None:
                    assert Hi == 2 * res.shape[1] and Wi == 2 * res.shape[2], \
                        'FPN levels must differ by exactly 2x (pad inputs to a multiple of 32)'
                self.gemm_f(C[i], sp.wref.w, lat[i], Ni * Hi * Wi, Ci, neck.out_channels, Hi, Wi,
                            shift=sp.bias, residual=res, ldr=neck.out_channels,
                            res_upsample2x=1 if res is not None else 0, round_out=True)
            outs = []
            for i in range(n):
                sp = neck._out[i]
                o = self.buf(*lat[i].shape)
                self.conv3_f(lat[i], sp.wref.w, o, shift=sp.bias, round_out=True)
                outs.append(o)
            while len(outs) < neck.num_outs:
                src = outs[-1]
                Ns, Hs, Ws, Cs = src.shape
                o = self.buf(Ns, (Hs + 1) // 2, (Ws + 1) // 2, Cs)
                L.call('subsample2', L.ptr(src), L.ptr(o), i32(Ns), i32(Hs), i32(Ws), i32(Cs),
                       L.stream())
                outs.append(o)
            self.lat, self.P = lat, outs
            # RPN: shared 3x3 conv + ReLU, fused cls/reg 1x1 head, per level
            cs, hs = rpn._conv_spec, rpn._head_spec
            self.rpn_t, self.rpn_out = [], []
            Wd = hs.wref.w.shape[0]
            for f in outs:
                Nf, Hf, Wf, Cf = f.shape
                t = self.buf(Nf, Hf, Wf, cs.wref.w.shape[0])
                self.conv3_f(f, cs.wref.w, t, shift=cs.bias, relu=True, round_out=True)
                o = self.buf(Nf, Hf, Wf, Wd)
                self.gemm_f(t, hs.wref.w, o, Nf * Hf * Wf, t.shape[-1], Wd, Hf, Wf, shift=hs.bias,
                            round_out=False)
                self.rpn_t.append(t)
                self.rpn_out.append(o)
            # gradient sinks of the FPN maps: RoIAlign backward (3 per step) accumulates straight
            # into them; zeroed here, inside the forward graph
            self.S = []
            if self.train:
                for f in outs:
                    sbuf = self.buf(*f.shape)
                    L.call('fill', L.ptr(sbuf), L.ll(sbuf.numel()), L.f32(0.0), L.stream())
                    self.S.append(sbuf)
        self.fwd_ready = True

    # ---------------------------------------------------------------- backward program
    def _block_bwd(self, rec, g, prev=None):
        """g = dL/d(pre-ReLU block output), already masked by (out > 0).  Launches the block's
        weight gradients and the data gradients down to conv1's input.  With `prev` (the block
        producing this block's input, same stage, so no downsample here) it also returns
        dL/d(input) = (x > 0) * (conv1 dgrad + g over the identity path), accumulating prev's
        beta gradients on the way; otherwise returns None and leaves rec['_dz1'] for the caller."""
        blk = rec['blk']
        s1, s2, s3, sd = blk._s1, blk._s2, blk._s3, blk._sd
        x, xs, h1, h2, col = rec['x'], rec['xs'], rec['h1'], rec['h2'], rec['col']
        N, H, W, Ci = x.shape
        _, Ho, Wo, p = h2.shape
        P1, P3 = N * H * W, N * Ho * Wo
        # conv3 (1x1)
        self.gemm_w(g, h2, s3.wref.grad, P3, p, 4 * p)
        dz2 = self.buf(N, Ho, Wo, p)
        self.gemm_d(g, s3.wref.w, dz2, P3, p, 4 * p, mask=h2, colsum=s2.bias_grad, round_out=True)
        # conv2 (3x3, stride 1 direct / stride 2 through im2col)
        dz1 = self.buf(N, H, W, p)
        if rec['stride'] == 1:
            self.conv3_w(dz2, h1, s2.wref.grad)
            self.conv3_d(dz2, s2.wref.w, dz1, mask=h1, colsum=s1.bias_grad, round_out=True)
        else:
            self.gemm_w(dz2, col, s2.wref.grad, P3, 9 * p, p)
            dcol = self.buf(P3, 9 * p)
            self.gemm_d(dz2, s2.wref.w, dcol, P3, 9 * p, p)
            L.call('col2im', L.ptr(dcol), L.ptr(dz1), L.ptr(h1), i32(N), i32(H), i32(W), i32(p),
                   i32(3), i32(3), i32(rec['stride']), i32(1), i32(9 * p), L.stream())
            self.colsum(dz1, P1, p, s1.bias_grad)
        # conv1 (1x1) and the downsample branch
        self.gemm_w(dz1, x, s1.wref.grad, P1, Ci, p)
        if sd is not None:
            self.gemm_w(g, xs, sd.wref.grad, P3, Ci, 4 * p)
        rec['_dz1'] = dz1
        if prev is None:
            return None
        assert sd is None, 'only the first block of a stage may have a downsample branch'
        dx = self.buf(N, H, W, Ci)
        self.gemm_d(dz1, s1.wref.w, dx, P1, Ci, p, residual=g, ldr=Ci, mask=x,
                    colsum=prev._s3.bias_grad,
                    colsum2=prev._sd.bias_grad if prev._sd is not None else None, round_out=True)
        return dx

    def build_backward_rpn(self):
        """RPN part of the backward (needs only the RPN losses' gradient gR): fused-head and shared
        3x3 conv weight / bias gradients over all levels, and the gradient the RPN path sends into
        each FPN map, written to self.B[l].  Its own program so that it can also be launched ahead
        of loss.backward() (Trunk.early_rpn_backward); by default it runs first thing inside
        _TrunkFn.backward."""
        rpn = self.model.rpn_head
        dev = self.dev
        with self.bwd_rpn:
            cs, hs = rpn._conv_spec, rpn._head_spec
            Wd = hs.wref.w.shape[0]
            Cr = cs.wref.w.shape[0]
            # narrow-head weight gradient: dW[Wd,Cr] = gR^T t through 32-column padded copies
            wtmp = self.buf(32, Cr)
            L.call('fill', L.ptr(wtmp), L.ll(wtmp.numel()), L.f32(0.0), L.stream())
            for l in range(len(self.P)):
                f, t, gR = self.P[l], self.rpn_t[l], self.gR[l]
                Nf, Hf, Wf, Cf = f.shape
                Pl = Nf * Hf * Wf
                self.colsum(gR, Pl, Wd, hs.bias_grad)
                pad = torch.zeros((Pl, 32), device=dev)
                self.keep.append(pad)
                L.call('copy2d', L.ptr(gR), L.ll(Wd), L.ptr(pad), L.ll(32), L.ll(Pl), i32(Wd),
                       i32(0), i32(0), L.stream())
                self.gemm_w(pad, t, wtmp, Pl, Cr, 32)
                dt = self.buf(Nf, Hf, Wf, Cr)
                self.gemm_d(gR, hs.wref.w, dt, Pl, Cr, Wd, mask=t, colsum=cs.bias_grad,
                            round_out=True)
                self.conv3_w(dt, f, cs.wref.grad)
                self.conv3_d(dt, cs.wref.w, self.B[l])
            L.call('copy2d', L.ptr(wtmp), L.ll(Cr), L.ptr(hs.wref.grad), L.ll(Cr), L.ll(Wd), i32(Cr),
                   i32(1), i32(0), L.stream())
        self.bwd_rpn_ready = True

    def build_backward(self):
        """Everything below the FPN maps.  self.B[l] holds the TOTAL gradient of FPN map l (RPN
        path from build_backward_rpn + what the RoI heads sent, added by _TrunkFn.backward)."""
        m = self.model
        bb, neck, rpn = m.backbone, m.neck, m.rpn_head
        nlev = len(self.P)
        n = len(self.lat)
        with self.bwd:
            gtot = self.B
            # The bias gradient of FPN output conv l is the per-channel sum of gtot[l]: it rides on
            # the LAST add that completes gtot[l] (add_colsum) instead of a read-only pass of its own.
            last_fold = {l - 1 for l in range(nlev - 1, n - 1, -1)}      # levels an extra level folds into

            def add_into(l, other, final):
                Nf, Hf, Wf, Cf = self.P[l].shape
                bg = neck._out[l].bias_grad if l < n else None
                if final and bg is not None and Cf % 4 == 0 and 256 % (Cf // 4) == 0:
                    L.call('add_colsum', L.ptr(gtot[l]), L.ptr(other), L.ptr(gtot[l]),
                           L.ll(Nf * Hf * Wf), i32(Cf), L.ptr(bg), i32(1), L.stream())
                    return True
                L.call('add', L.ptr(gtot[l]), L.ptr(other), L.ptr(gtot[l]), L.ll(gtot[l].numel()),
                       i32(1), L.stream())
                return False

            summed = set()
            for l in range(len(self.S)):       # what the RoI heads sent (RoIAlign backward sinks)
                if add_into(l, self.S[l], final=l not in last_fold):
                    summed.add(l)
            # extra levels (P6 = P5-out subsampled): fold their gradient into the level below
            for l in range(nlev - 1, n - 1, -1):
                Nf, Hf, Wf, Cf = self.P[l - 1].shape
                u = self.buf(Nf, Hf, Wf, Cf)
                L.call('subsample2_bwd', L.ptr(gtot[l]), L.ptr(u), None, i32(Nf), i32(Hf), i32(Wf),
                       i32(Cf), L.stream())
                if add_into(l - 1, u, final=True):      # a level is folded into once, after its sink
                    summed.add(l - 1)
            for l in range(n):
                if l in summed:
                    continue
                Nf, Hf, Wf, Cf = self.P[l].shape
                self.colsum(gtot[l], Nf * Hf * Wf, Cf, neck._out[l].bias_grad)
            # FPN 3x3 output convs, fine -> coarse (the top-down path's gradient flows that way)
            dlat = [None] * n
            for l in range(n):
                sp = neck._out[l]
                self.conv3_w(gtot[l], self.lat[l], sp.wref.grad)
                d = self.buf(*self.lat[l].shape)
                self.conv3_d(gtot[l], sp.wref.w, d,
                             colsum=(neck._lat[0].bias_grad if l == 0 else None), round_out=True)
                if l == 0:
                    dlat[0] = d
                else:
                    Nl, Hl, Wl, Cl = d.shape
                    dl = self.buf(Nl, Hl, Wl, Cl)
                    L.call('sum2x2_add', L.ptr(dlat[l - 1]), L.ptr(d), L.ptr(dl), i32(Nl), i32(Hl),
                           i32(Wl), i32(Cl), L.stream())
                    self.colsum(dl, Nl * Hl * Wl, Cl, neck._lat[l].bias_grad)
                    dlat[l] = dl
            for l in range(n):
                Nl, Hl, Wl, Cc = self.C[l].shape
                self.gemm_w(dlat[l], self.C[l], neck._lat[l].wref.grad, Nl * Hl * Wl, Cc,
                            neck.out_channels)
            # ResNet stages, top down.  The gradient of stage output C[s] is
            #   (C[s] > 0) * (lateral dgrad + next stage's conv1 dgrad + its strided-downsample dgrad)
            # assembled by chaining epilogue residuals, never by a separate add / mask pass.
            g = None
            for s in range(len(self.stages) - 1, -1, -1):
                recs = self.stages[s]
                if not recs[0]['blk']._trainable:
                    break
                if self.split_stage is not None and s == self.split_stage - 1:
                    # everything from here on goes to the tail program: the gradients of stages
                    # >= split_stage, the FPN and the RPN are final at this point
                    self.bwd.__exit__(None, None, None)
                    self.bwd_tail.__enter__()
                last = recs[-1]['blk']
                ldb = last._s3.bias_grad
                ldb2 = last._sd.bias_grad if last._sd is not None else None
                Ns, Hs, Ws, Cc = self.C[s].shape
                Ps = Ns * Hs * Ws
                if s == len(self.stages) - 1:
                    g = self.buf(Ns, Hs, Ws, Cc)
                    self.gemm_d(dlat[s], neck._lat[s].wref.w, g, Ps, Cc, neck.out_channels,
                                mask=self.C[s], colsum=ldb, colsum2=ldb2, round_out=True)
                # else: g was produced by the block-0 backward of stage s+1 (below)
                for bi in range(len(recs) - 1, -1, -1):
                    rec = recs[bi]
                    if bi > 0:
                        g = self._block_bwd(rec, g, prev=recs[bi - 1]['blk'])
                    else:
                        self._block_bwd(rec, g)
                        below = s > 0 and self.stages[s - 1][0]['blk']._trainable
                        g = self._stage_boundary_bwd(rec, g, s - 1, dlat[s - 1]) if below else None
            if L.RECORD is self.bwd_tail.calls:
                # leave the recording state the enclosing `with self.bwd` expects to close
                self.bwd_tail.__exit__(None, None, None)
                L.RECORD = self.bwd.calls
        self.bwd_ready = True

    def _stage_boundary_bwd(self, rec, g, s_in, dlat_in):
        """Block 0 of a stage (after _block_bwd): its input C[s_in] also feeds FPN lateral s_in.
        Returns the masked gradient of C[s_in] (= the `g` of stage s_in's last block) and
        accumulates that block's beta gradient:
            t  = dgrad of the (strided) downsample 1x1            [half resolution]
            u  = lateral dgrad + zero-stuffed t                   [epilogue residual]
            dx = (x > 0) * (conv1 dgrad + u)                      [epilogue residual + mask]"""
        neck = self.model.neck
        blk = rec['blk']
        s1, sd = blk._s1, blk._sd
        x, h2, dz1 = rec['x'], rec['h2'], rec['_dz1']
        N, H, W, Ci = x.shape
        _, Ho, Wo, p = h2.shape
        P1, P3 = N * H * W, N * Ho * Wo
        prev = self.stages[s_in][-1]['blk']
        assert sd is not None and sd.stride in (1, 2)
        t = self.buf(N, Ho, Wo, Ci)
        self.gemm_d(g, sd.wref.w, t, P3, Ci, 4 * p)
        u = self.buf(N, H, W, Ci)
        self.gemm_d(dlat_in, neck._lat[s_in].wref.w, u, P1, Ci, neck.out_channels, H, W,
                    residual=t, ldr=Ci, res_upsample2x=(2 if sd.stride == 2 else 0))
        dx = self.buf(N, H, W, Ci)
        self.gemm_d(dz1, s1.wref.w, dx, P1, Ci, p, residual=u, ldr=Ci, mask=x,
                    colsum=prev._s3.bias_grad,
                    colsum2=prev._sd.bias_grad if prev._sd is not None else None, round_out=True)
        return dx


class _TrunkFn(Function):
    """(img, trainable parameters as autograd triggers) -> (FPN maps..., fused RPN maps...)."""

    @staticmethod
    def forward(ctx, img, trunk, *params):
        prog = trunk.run_forward(img)
        ctx.prog, ctx.trunk = prog, trunk
        # Outputs that receive no gradient through autograd (the FPN maps: RoIAlign's backward
        # accumulates straight into the gradient sinks; the fused RPN maps when the RPN loss wrote
        # its gradient itself) must arrive as None, not as materialised zero tensors -- per step
        # those were five zero-fills, five layout copies and five adds over up to 134 MB each
        # (~0.35 ms, profiles/r02_launch_summary.txt).
        ctx.set_materialize_grads(False)
        outs = [t.permute(0, 3, 1, 2) for t in prog.P] + \
               [t.permute(0, 3, 1, 2) for t in prog.rpn_out]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        prog, trunk = ctx.prog, ctx.trunk
        trunk.store.queue_finalize()
        # every gradient that reaches the FPN maps has passed through the RoI heads: their
        # parameter gradients are final -> a data-parallel trainer starts exchanging them now,
        # under the ~4.6 ms of trunk backward that follow
        trunk.store.heads_done()
        n = len(prog.P)
        if not prog.rpn_done:
            # the RPN losses were not back-propagated early: their gradient arrives here
            trunk.run_rpn_backward(prog, grads[n:], side=False)
        trunk.run_main_backward(prog, grads[:n])
        return (None,) * len(ctx.needs_input_grad)


class Trunk:
    """Per-model manager: one recorded program set per input shape (at most four shapes cached,
    oldest evicted).  A program owns its activations, so only ONE step per program may be in
    flight: forward, then backward, then the next forward."""

    def __init__(self, model, store):
        self.model, self.store = model, store
        self.progs = {}
        self.params = [p for mod in (model.backbone, model.neck, model.rpn_head)
                       for p in mod.parameters() if p.requires_grad]
        self.side = torch.cuda.Stream(device=store.device)
        # When the RPN losses are back-propagated ahead of loss.backward():
        #   'end'  (default) at the end of forward_train, on the main stream: its 1.3 ms of GEMMs
        #          keep the GPU busy while the launch thread sums the losses and starts the
        #          autograd engine (measured 18.4 -> 18.2 ms/step);
        #   'side' right after the RPN loss, on a side stream under the proposal / sampling phase
        #          (measured: no gain -- that phase's bubbles are many and short);
        #   '0'    inside _TrunkFn.backward, like any other node.
        self.early_rpn = os.environ.get('LOFT_EARLY_RPN', 'end')
        self.bwd_sm_reserve = 0        # set by a distributed Trainer that overlaps the exchange
        # main backward in two programs: stages >= split_stage (+ FPN, RPN) first, the rest after
        # `store.upper_done()` -- only when something listens (store.upper_done_hooks)
        self.split_stage = int(os.environ.get('LOFT_SPLIT_BWD', '2'))
        self.current = None
        self.step_ctx = (None, None)
        store.pre_finalize.append(self.flush)

    @staticmethod
    def eligible(model):
        from .models.backbones.resnet import Bottleneck, ResNet
        from .models.dense_heads.rpn_head import RPNHead
        from .models.necks.fpn import FPN
        if os.environ.get('LOFT_TRUNK', '1') == '0':
            return False
        bb, neck, rpn = getattr(model, 'backbone', None), getattr(model, 'neck', None), \
            getattr(model, 'rpn_head', None)
        if type(bb) is not ResNet or type(neck) is not FPN or type(rpn) is not RPNHead:
            return False
        if bb.style != 'pytorch' or neck.start_level != 0 or \
                len(neck.lateral_convs) != len(bb.res_layers) or \
                tuple(bb.out_indices) != tuple(range(len(bb.res_layers))):
            return False
        return all(isinstance(b, Bottleneck) for n in bb.res_layers for b in getattr(bb, n))

    def run_forward(self, img):
        N, _, H, W = img.shape
        metas, pcfg = self.step_ctx
        shapes = tuple(tuple(m['img_shape'][:2]) for m in metas) if metas else None
        key = (N, H, W, shapes)
        prog = self.progs.get(key)
        if prog is None:
            if len(self.progs) >= 4:          # ~3 GB of static buffers per 2x1024^2 program
                self.progs.pop(next(iter(self.progs)))
            prog = self.progs[key] = _Program(self.model, N, H, W, self.store.device, True)
        prog.img.copy_(img)
        if not prog.fwd_ready:
            prog.build_forward()
            prog.B = [torch.zeros_like(f) for f in prog.P]          # total gradient of FPN map l
            prog.gR = [torch.zeros_like(o) for o in prog.rpn_out]   # gradient of fused RPN map l
            prog.proposals, prog.proposals_fresh = None, False
            if prog.use_graph:
                self._capture_forward(prog, metas, pcfg)
        else:
            prog.fwd.launch()
            prog.proposals_fresh = prog.proposals is not None and prog.fwd.graph is not None
        prog.rpn_done = False
        self.current = prog
        return prog

    def _capture_forward(self, prog, metas, pcfg):
        """Forward graph = the recorded launches + (when the caller gave the proposal config) the
        whole RPN proposal generation: per-level sigmoid / sort / top-k decode, the global sort,
        per-level NMS and the fixed-size [N, nms_post, 5] proposal block all have static shapes,
        so ~80 small kernels of RPNHead.get_bboxes ride in the same graph launch instead of
        being issued one by one by the (forward-phase-bound) launch thread."""
        rpn = self.model.rpn_head
        want = pcfg is not None and metas is not None and \
            os.environ.get('LOFT_GRAPH_PROPOSALS', '1') != '0'
        for with_props in ([True, False] if want else [False]):
            g = torch.cuda.CUDAGraph()
            try:
                with torch.no_grad(), torch.cuda.graph(g, capture_error_mode='thread_local'):
                    prog.fwd.replay()
                    if with_props:
                        outs = rpn.outs_from_fused([t.permute(0, 3, 1, 2) for t in prog.rpn_out])
                        prog.proposals = rpn.get_bboxes(*outs, metas, cfg=pcfg, fixed_size=True)
                prog.fwd.graph = g
                return
            except Exception as e:      # an op in get_bboxes that cannot be captured
                prog.proposals = None
                if not with_props:
                    raise
                import warnings
                warnings.warn(f'proposal generation not captured into the forward graph: {e}')

    def proposals(self):
        """The proposal block the forward graph produced for the current step (None on the
        recording step or when not captured: LOFT_GRAPH_PROPOSALS=0)."""
        prog = self.current
        if prog is None or not prog.proposals_fresh:
            return None
        return prog.proposals

    def direct_rpn_grads(self):
        """True when the RPN loss may write its gradient straight into the RPN backward program's
        input buffers (training step in flight, default scheduling of the RPN backward)."""
        return self.current is not None and torch.is_grad_enabled() and \
            self.early_rpn in ('end', 'side') and os.environ.get('LOFT_RPN_DIRECT', '1') != '0'

    def rpn_backward_direct(self):
        """Launch the RPN part of the backward; prog.gR was filled by the fused RPN loss."""
        if torch.is_grad_enabled():
            self.run_rpn_backward(self.current, None, side=False)

    def run_rpn_backward(self, prog, grads, side):
        for buf, g in zip(prog.gR, grads if grads is not None else ()):
            if g is None:
                buf.zero_()
            else:
                buf.copy_(g.permute(0, 2, 3, 1))
        main = torch.cuda.current_stream(self.store.device)
        if side:
            self.side.wait_stream(main)
            ctx = torch.cuda.stream(self.side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            if not prog.bwd_rpn_ready:
                prog.build_backward_rpn()          # executes the program once while recording it
                if prog.use_graph:
                    prog.bwd_rpn.capture()
            else:
                prog.bwd_rpn.launch()
            if side:
                prog.rpn_event = self.side.record_event()
        prog.rpn_done = True

    def run_main_backward(self, prog, grads):
        main = torch.cuda.current_stream(self.store.device)
        if prog.rpn_event is not None:
            main.wait_event(prog.rpn_event)
            prog.rpn_event = None
        for buf, g in zip(prog.B, grads):
            if g is not None:
                gn = g.permute(0, 2, 3, 1)
                if not gn.is_contiguous():
                    gn = gn.contiguous()
                L.call('add', L.ptr(buf), L.ptr(gn), L.ptr(buf), L.ll(buf.numel()), i32(1),
                       L.stream())
        if not prog.bwd_ready:
            # with a concurrent gradient exchange, leave its CTAs a few SMs (baked into the graph)
            prev = L.lib().loft_reserve_sms(self.bwd_sm_reserve) if self.bwd_sm_reserve else 0
            prog.split_stage = self.split_stage if self.store.upper_done_hooks else None
            prog.build_backward()
            if prog.use_graph:
                prog.bwd.capture()
                if prog.bwd_tail.calls:
                    prog.bwd_tail.capture()
            if self.bwd_sm_reserve:
                L.lib().loft_reserve_sms(prev)
            # (the recording pass executed both parts back to back: this step's upper-stage
            # gradients are exchanged with the rest, after the backward)
            # The tapes pin thousands of small Python objects; move them (and everything else
            # built so far) out of the cyclic GC's reach so a generation-2 sweep never stalls the
            # launch thread in the middle of a step.
            import gc
            gc.collect()
            gc.freeze()
        else:
            prog.bwd.launch()
            if prog.bwd_tail.calls:
                self.store.upper_done()
                prog.bwd_tail.launch()
        prog.rpn_done = False

    def early_rpn_backward(self, losses, when):
        """Back-propagate the RPN losses ahead of the outer loss.backward() (see __init__).
        Returns the same dict with detached values so the outer backward does not visit them
        again."""
        prog = self.current
        if prog is None or self.early_rpn != when or not torch.is_grad_enabled():
            return losses
        terms = [t for v in losses.values() for t in (v if isinstance(v, (list, tuple)) else [v])]
        if not any(t.requires_grad for t in terms):
            return losses
        total = terms[0]
        for t in terms[1:]:
            total = total + t
        grads = torch.autograd.grad(total, self.rpn_outs, allow_unused=True)
        self.run_rpn_backward(prog, grads, side=(when == 'side'))
        return {k: ([t.detach() for t in v] if isinstance(v, (list, tuple)) else v.detach())
                for k, v in losses.items()}

    def flush(self):
        """Called before gradient finalisation: if the RPN part ran early but nothing reached the
        FPN maps from the RoI heads, the main backward still has to run."""
        prog = self.current
        if prog is not None and prog.rpn_done:
            self.run_main_backward(prog, [None] * len(prog.P))

    def __call__(self, img, img_metas=None, proposal_cfg=None):
        self.step_ctx = (img_metas, proposal_cfg)
        outs = _TrunkFn.apply(img, self, *self.params)
        n = len(outs) // 2
        self.rpn_outs = list(outs[n:])
        feats = tuple(outs[:n])
        if self.current is not None and self.current.S:
            for f, sbuf in zip(feats, self.current.S):
                f._loft_grad_sink = sbuf
        return feats, self.rpn_outs
