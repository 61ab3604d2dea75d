"""GPU parity tests, kernel by kernel: every C-ABI entry point of libloft_b200.so is compared with
the CPU oracle (oracle/) or -- for floating-point contractions -- with a plain torch fp32
reference (TF32 disabled).  Index outputs (NMS keep, assignment, mask targets) must be bit-exact;
TF32 contractions must stay within 2e-3 relative L2 (operands rounded to 10-bit mantissas)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# Operands of these single-layer tests are pre-rounded to the TF32 grid, so every product is exact
# in fp32 and the pre-activations differ from the fp32 reference by accumulation order only
# (~1e-6 relative): a ReLU mask practically never flips, and what remains is the RNA rounding of
# the stored output (`round_out`, 2^-11 / sqrt(3) ~ 2.8e-4 relative per value).  A wrong filter
# tap, halo row or tile edge shows up at 1e-2 .. 1, far above these bounds.
TF32_TOL = 5e-4          # forward outputs and data gradients (both stored TF32-rounded)
GRAD_TOL = 2e-3          # weight / bias gradients (fp32 atomics in varying order over <= 1e5 terms)


@pytest.fixture(scope='module', autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def rnd(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()


def tf32_round(t):
    """round-to-nearest-away to a 10-bit mantissa, like cvt.rna.tf32.f32"""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def relu_like(z, y_gpu):
    """Reference ReLU whose BACKWARD uses the mask of the GPU's own output: a pre-activation that
    is zero to 1e-7 may land on either side, and ONE flipped unit with a large upstream gradient
    moves the data gradient of its 3x3 neighbourhood by percents of the whole tensor's norm
    (measured: tools/diag_dgrad.py -- 1 flip in 650 k elements = 1.0e-2 relative L2, the other
    three groups of the same launch 2.1e-4).  Forward values are unaffected (|z| ~ 1e-7 there)."""
    return z * (y_gpu.detach() > 0).to(z.dtype)


class _Store:
    device = torch.device('cuda')

    def queue_finalize(self):
        pass


def _wref(w, grad=True):
    from bonai_b200.engine import WeightRef
    return WeightRef(w, torch.zeros_like(w) if grad else None)


# ----------------------------------------------------------------------------- dense (tcgen05)
@pytest.mark.parametrize('shape', [(2, 32, 32, 64, 128), (11, 7, 7, 256, 256), (3, 14, 14, 256, 256),
                                   (2, 20, 25, 32, 64), (1, 64, 64, 256, 256)])
def test_conv3x3_fwd_bwd(shape):
    from bonai_b200.ops import dense as D
    N, H, W, Ci, Co = shape
    x = tf32_round(rnd(N, Ci, H, W, seed=1)).contiguous(memory_format=torch.channels_last)
    w = tf32_round(rnd(Co, Ci, 3, 3, seed=2, scale=0.05)).contiguous(memory_format=torch.channels_last)
    b = rnd(Co, seed=3)
    gb = torch.zeros_like(b)
    wref = _wref(w)
    spec = D.ConvSpec(wref, ksize=3, padding=1, relu=True, bias=b, bias_grad=gb, store=_Store())
    xg = x.clone().requires_grad_(True)
    trig = torch.zeros(1, device='cuda', requires_grad=True)
    y = D.conv(xg, spec, triggers=(trig,))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    zr = F.conv2d(xr, wr, br, padding=1)
    yr = relu_like(zr, y)
    assert rel(y, F.relu(zr)) < TF32_TOL
    dy = tf32_round(rnd(*y.shape, seed=4))
    y.backward(dy)
    yr.backward(dy)
    assert rel(xg.grad, xr.grad) < GRAD_TOL
    assert rel(wref.grad, wr.grad) < GRAD_TOL
    assert rel(gb, br.grad) < GRAD_TOL


@pytest.mark.parametrize('cfg', [dict(k=1, s=1, ci=256, co=64), dict(k=1, s=2, ci=256, co=512),
                                 dict(k=3, s=2, ci=128, co=128)])
def test_conv_bn_residual_paths(cfg):
    """1x1, strided 1x1 (subsample) and strided 3x3 (im2col) with BN-eval affine + residual + ReLU,
    including dgamma / dbeta."""
    from bonai_b200.engine import BNRef
    from bonai_b200.ops import dense as D
    k, s, Ci, Co = cfg['k'], cfg['s'], cfg['ci'], cfg['co']
    N, H, W = 2, 16, 16
    pad = 1 if k == 3 else 0
    x = tf32_round(rnd(N, Ci, H, W, seed=1)).contiguous(memory_format=torch.channels_last)
    w = tf32_round(rnd(Co, Ci, k, k, seed=2, scale=0.05))
    if k > 1:
        w = w.contiguous(memory_format=torch.channels_last)
    gamma, beta = rnd(Co, seed=3).abs() + 0.3, rnd(Co, seed=4) * 0.1
    mean, var = rnd(Co, seed=5) * 0.1, rnd(Co, seed=6).abs() + 0.5
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale, shift = gamma * rstd, beta - mean * gamma * rstd
    dgamma, dbeta = torch.zeros(Co, device='cuda'), torch.zeros(Co, device='cuda')
    bn = BNRef(scale, shift, rstd, mean, dgamma, dbeta)
    # eval-mode BN is folded into the weights the tensor cores read: T = tf32(scale * W)
    import ctypes
    from bonai_b200 import _lib as L
    Kw = Ci * k * k
    wflat = w.permute(0, 2, 3, 1).contiguous().view(-1) if k > 1 else w.contiguous().view(-1)
    tflat = torch.empty_like(wflat)
    gflat = torch.zeros_like(wflat)
    off = (torch.arange(Co, dtype=torch.int64) * Kw).cuda()
    rk = torch.full((Co,), Kw, dtype=torch.int32).cuda()
    ch = torch.arange(Co, dtype=torch.int32).cuda()
    L.call('bn_fold_weights', L.ptr(wflat), L.ptr(tflat), L.ptr(off), L.ptr(rk), L.ptr(ch),
           L.ptr(scale), L.ll(Co), L.stream())
    shape4 = (Co, k, k, Ci)
    from bonai_b200.engine import WeightRef
    wref = WeightRef(tflat.view(shape4).permute(0, 3, 1, 2), gflat.view(shape4).permute(0, 3, 1, 2))
    spec = D.ConvSpec(wref, ksize=k, stride=s, padding=pad, relu=True, bn=bn, bn_trainable=True,
                      store=_Store())
    Ho = (H + 2 * pad - k) // s + 1
    res = tf32_round(rnd(N, Co, Ho, Ho, seed=7)).contiguous(memory_format=torch.channels_last)
    xg, rg = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    y = D.conv(xg, spec, residual=rg)
    xr, rr, wr = x.clone().requires_grad_(True), res.clone().requires_grad_(True), \
        w.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.conv2d(xr, wr, stride=s, padding=pad)
    zr = F.batch_norm(z, mean, var, gr, br, False, 0.0, 1e-5) + rr
    yr = relu_like(zr, y)
    assert rel(y, F.relu(zr)) < TF32_TOL
    dy = tf32_round(rnd(*y.shape, seed=8))
    y.backward(dy)
    yr.backward(dy)
    # gamma's gradient and the un-folded weight gradient come from the accumulated dW' rows
    L.call('bn_finalize', L.ptr(wflat), L.ptr(gflat), L.ptr(off), L.ptr(rk), L.ptr(ch),
           L.ptr(scale), L.ptr(rstd), L.ptr(mean), L.ptr(dbeta), L.ptr(dgamma), L.ll(Co),
           L.stream())
    assert rel(xg.grad, xr.grad) < GRAD_TOL
    assert rel(rg.grad, rr.grad) < GRAD_TOL
    assert rel(wref.grad, wr.grad) < GRAD_TOL
    assert rel(dgamma, gr.grad) < GRAD_TOL
    assert rel(dbeta, br.grad) < GRAD_TOL


def test_fpn_lateral_upsample_add():
    from bonai_b200.ops import dense as D
    N, Ci, Co, H = 2, 512, 256, 16
    x = tf32_round(rnd(N, Ci, H, H, seed=1)).contiguous(memory_format=torch.channels_last)
    coarse = tf32_round(rnd(N, Co, H // 2, H // 2, seed=2)).contiguous(memory_format=torch.channels_last)
    w = tf32_round(rnd(Co, Ci, 1, 1, seed=3, scale=0.05))
    b, gb = rnd(Co, seed=4), torch.zeros(Co, device='cuda')
    wref = _wref(w)
    spec = D.ConvSpec(wref, ksize=1, bias=b, bias_grad=gb, res_upsample=True, store=_Store())
    xg, cg = x.clone().requires_grad_(True), coarse.clone().requires_grad_(True)
    y = D.conv(xg, spec, residual=cg)
    xr, cr, wr = x.clone().requires_grad_(True), coarse.clone().requires_grad_(True), \
        w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, b) + F.interpolate(cr, scale_factor=2, mode='nearest')
    assert rel(y, yr) < TF32_TOL
    dy = tf32_round(rnd(*y.shape, seed=5))
    y.backward(dy)
    yr.backward(dy)
    assert rel(xg.grad, xr.grad) < TF32_TOL
    assert rel(cg.grad, cr.grad) < 1e-3
    assert rel(wref.grad, wr.grad) < TF32_TOL


@pytest.mark.parametrize('shape', [(300, 12544, 1024, True), (77, 1024, 8, False), (50, 1024, 4, False)])
def test_linear(shape):
    from bonai_b200.ops import dense as D
    P, K, Co, relu = shape
    x = tf32_round(rnd(P, K, seed=1))
    w = tf32_round(rnd(Co, K, seed=2, scale=0.03))
    b, gb = rnd(Co, seed=3), torch.zeros(Co, device='cuda')
    wref = _wref(w)
    spec = D.ConvSpec(wref, relu=relu, bias=b, bias_grad=gb, store=_Store())
    xg = x.clone().requires_grad_(True)
    y = D.linear(xg, spec)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), \
        b.clone().requires_grad_(True)
    zr = F.linear(xr, wr, br)
    yr = relu_like(zr, y) if relu else zr
    assert rel(y, F.relu(zr) if relu else zr) < TF32_TOL
    dy = tf32_round(rnd(P, Co, seed=4))
    y.backward(dy)
    yr.backward(dy)
    assert rel(xg.grad, xr.grad) < GRAD_TOL
    assert rel(wref.grad, wr.grad) < GRAD_TOL
    assert rel(gb, br.grad) < GRAD_TOL


def test_deconv2x2():
    from bonai_b200.ops import dense as D
    N, Ci, Co, H = 5, 256, 256, 14
    x = tf32_round(rnd(N, Ci, H, H, seed=1)).contiguous(memory_format=torch.channels_last)
    wt = tf32_round(rnd(Ci, Co, 2, 2, seed=2, scale=0.05))       # ConvTranspose2d layout
    b = rnd(Co, seed=3)
    wp = wt.permute(2, 3, 1, 0).reshape(4 * Co, Ci).contiguous()  # [(i,j,co), ci]
    gb = torch.zeros(Co, device='cuda')
    wref = _wref(wp)
    spec = D.ConvSpec(wref, relu=True, bias=b.repeat(4).contiguous(), bias_grad=gb, store=_Store())
    xg = x.clone().requires_grad_(True)
    y = D.deconv2x2(xg, spec)
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), \
        b.clone().requires_grad_(True)
    zr = F.conv_transpose2d(xr, wr, br, stride=2)
    yr = relu_like(zr, y)
    assert rel(y, F.relu(zr)) < TF32_TOL
    dy = tf32_round(rnd(*y.shape, seed=4))
    y.backward(dy)
    yr.backward(dy)
    assert rel(xg.grad, xr.grad) < GRAD_TOL
    assert rel(wref.grad, wr.grad.permute(2, 3, 1, 0).reshape(4 * Co, Ci)) < GRAD_TOL
    assert rel(gb, br.grad) < GRAD_TOL


@pytest.mark.parametrize('hw', [(64, 64), (70, 90), (256, 320), (33, 1100)])
def test_stem_and_maxpool(hw):
    """Direct 7x7/2 stem (packed NHWC4 image read through an overlapping-stride tensor map) against
    torch fp32 conv2d; even / odd sizes, rows wider than one 256-column tile."""
    from bonai_b200.ops import misc as M
    img = rnd(2, 3, *hw, seed=1)
    w = rnd(64, 3, 7, 7, seed=2, scale=0.1)
    scale, shift = rnd(64, seed=3).abs() + 0.5, rnd(64, seed=4) * 0.1
    wp = M.pack_stem_weight(tf32_round(w.permute(0, 2, 3, 1).contiguous()))
    y = M.stem_conv(img, wp, scale, shift)
    yr = F.relu(F.conv2d(tf32_round(img), tf32_round(w), stride=2, padding=3) *
                scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert y.shape == yr.shape
    assert rel(y, yr) < TF32_TOL
    # the border rows / columns read the zero padding of the packed image
    assert rel(y[:, :, :2], yr[:, :, :2]) < TF32_TOL and rel(y[:, :, -2:], yr[:, :, -2:]) < TF32_TOL
    assert rel(y[..., :2], yr[..., :2]) < TF32_TOL and rel(y[..., -2:], yr[..., -2:]) < TF32_TOL
    p = M.maxpool3x3s2(y)
    assert torch.equal(p, F.max_pool2d(y, 3, 2, 1))


# ----------------------------------------------------------------------------- RoIAlign
def test_roi_align_golden(golden_units):
    from bonai_b200.ops import roi_align
    g = golden_units
    fm = torch.from_numpy(g['ra_feat']).cuda()
    fm4 = torch.cat([fm, fm[:, :1]], 1).contiguous(memory_format=torch.channels_last)  # C=4
    rois = torch.from_numpy(g['ra_rois']).cuda()
    for key, (s, sc) in dict(ra_out7_s4=(7, 0.25), ra_out14_s8=(14, 0.125),
                             ra_out28_s1=(28, 1.0)).items():
        out = roi_align(fm4, rois, s, sc, 0, 'avg', True)[:, :3]
        ref = torch.from_numpy(g[key]).cuda()
        assert (out - ref).abs().max() < 2e-3 * ref.abs().max(), key    # output is TF32-rounded


def test_multilevel_roi_align_fwd_bwd_vs_oracle():
    from bonai_b200.ops import multilevel_roi_align
    from oracle import loft_cpu as O
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(2, 8, 256 // s, 256 // s, generator=g) for s in (4, 8, 16, 32)]
    c = torch.rand(40, 2, generator=g) * 256
    wh = torch.exp(torch.rand(40, 2, generator=g) * 4.0) * 6
    b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 256)
    rois = torch.cat([torch.randint(0, 2, (40, 1), generator=g).float(), b], 1)
    fo = [f.clone().requires_grad_(True) for f in feats]
    ref = O.roi_extract(fo, rois, 7)
    fg = [f.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
          for f in feats]
    out = multilevel_roi_align(fg, rois.cuda(), 7, [4, 8, 16, 32], 56)
    assert (out.cpu() - ref).abs().max() < 2e-3 * ref.abs().max()
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    out.backward(dy.cuda())
    for a, bb in zip(fg, fo):
        if bb.grad is None:                       # level received no RoI
            assert float(a.grad.abs().max()) == 0.0
        else:
            assert rel(a.grad.cpu(), bb.grad) < 1e-4


def test_mask_target_bit_exact_vs_oracle():
    from bonai_b200.ops import mask_target_sample
    from oracle import loft_cpu as O
    _, gb, gl, gm, go = O.make_inputs(3, 1, 256, 12)
    g = torch.Generator().manual_seed(1)
    inds = torch.randint(0, 12, (60,), generator=g)
    props = gb[0][inds] + torch.randn(60, 4, generator=g) * 4
    props[:5] = gb[0][inds[:5]]
    sr = dict(pos_bboxes=props, pos_assigned_gt_inds=inds)
    ref = O.mask_target([sr], [gm[0]])
    out = mask_target_sample(gm[0].cuda().contiguous(), props.cuda(), inds.cuda(), 28)
    assert torch.equal(out.cpu(), ref)


# ----------------------------------------------------------------------------- box kernels
def test_iou_assign_bit_exact(golden_units):
    from bonai_b200.core import MaxIoUAssigner
    from oracle import loft_cpu as O
    g = golden_units
    b1, b2 = torch.from_numpy(g['iou_b1']).cuda(), torch.from_numpy(g['iou_b2']).cuda()
    for name, (pos, neg, mn) in dict(rpn=(0.7, 0.3, 0.3), rcnn=(0.5, 0.5, 0.5)).items():
        r = MaxIoUAssigner(pos, neg, mn).assign(b1, b2, gt_labels=torch.zeros(7, dtype=torch.long,
                                                                               device='cuda'))
        assert torch.equal(r.gt_inds.cpu(), torch.from_numpy(g[f'assign_{name}_gt_inds']))
        assert torch.equal(r.max_overlaps.cpu(), torch.from_numpy(g[f'assign_{name}_max_overlaps']))
        assert torch.equal(r.labels.cpu(), torch.from_numpy(g[f'assign_{name}_labels']))
    # anchors of a 256^2 tile against synthetic GTs, vs the oracle restatement
    _, gb, _, _, _ = O.make_inputs(1, 1, 256, 10)
    anchors = torch.cat(O.grid_anchors([(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]))
    gi, mo, _ = O.max_iou_assign(anchors, gb[0], 0.7, 0.3, 0.3)
    r = MaxIoUAssigner(0.7, 0.3, 0.3).assign(anchors.cuda(), gb[0].cuda())
    assert torch.equal(r.gt_inds.cpu(), gi)
    assert torch.equal(r.max_overlaps.cpu(), mo)
    # empty cases (tests/test_assigner.py:65-152 of the reference)
    r = MaxIoUAssigner(0.5, 0.5).assign(anchors[:10].cuda(), torch.empty(0, 4).cuda())
    assert r.gt_inds.tolist() == [0] * 10
    r = MaxIoUAssigner(0.5, 0.5).assign(torch.empty(0, 4).cuda(), gb[0].cuda())
    assert r.gt_inds.numel() == 0


def test_nms_bit_exact(golden_units):
    from bonai_b200.ops import nms, batched_nms
    g = golden_units
    boxes = torch.from_numpy(g['nms_boxes']).cuda()
    scores = torch.from_numpy(g['nms_scores']).cuda()
    ids = torch.from_numpy(g['nms_ids']).cuda()
    dets, keep = nms(boxes, scores, 0.7)
    assert torch.equal(keep.cpu(), torch.from_numpy(g['nms_keep']))
    dets, keep = batched_nms(boxes, scores, ids, dict(type='nms', iou_threshold=0.7))
    assert torch.equal(keep.cpu(), torch.from_numpy(g['bnms_keep']))
    assert torch.equal(dets.cpu(), torch.from_numpy(g['bnms_dets']))


@pytest.mark.parametrize('sizes', [(3000, 3000, 3000, 3000, 768), (700, 1, 0, 65, 64), (129,)])
def test_segmented_nms_equals_global_batched_nms(sizes, golden_units):
    """The per-level NMS of the RPN path (loft_nms_segmented) must give exactly the keep list of
    the global batched NMS (loft_nms_sorted with idxs = level, itself pinned to the reference's
    mmcv batched_nms golden above): clustered boxes, many exact score ties, two images."""
    from bonai_b200.ops import nms_sorted, nms_segmented
    g = torch.Generator().manual_seed(sum(sizes))
    B, n = 2, sum(sizes)
    ctr = torch.rand(B, 40, 2, generator=g) * 900 + 60
    boxes_l, scores_l, ids_l = [], [], []
    for l, k in enumerate(sizes):
        c = ctr[:, torch.randint(0, 40, (k,), generator=g)] + torch.randn(B, k, 2, generator=g) * 6
        wh = torch.rand(B, k, 2, generator=g) * 60 + 8
        bx = torch.cat([c - wh / 2, c + wh / 2], -1).clamp(0, 1024)
        sc = (torch.rand(B, k, generator=g) * 50).round() / 50          # ties
        sc, o = sc.sort(dim=1, descending=True, stable=True)
        boxes_l.append(torch.gather(bx, 1, o[:, :, None].expand(-1, -1, 4)))
        scores_l.append(sc)
        ids_l.append(torch.full((B, k), l, dtype=torch.long))
    bx, sc, ids = (torch.cat(t, 1).cuda() for t in (boxes_l, scores_l, ids_l))
    sc_s, order = sc.sort(dim=1, descending=True, stable=True)
    bx_s = torch.gather(bx, 1, order[:, :, None].expand(-1, -1, 4)).contiguous()
    ids_s = torch.gather(ids, 1, order).contiguous()
    for max_keep in (-1, 1000):
        k0, n0 = nms_sorted(bx_s, ids_s, 0.7, max_keep)
        k1, n1 = nms_segmented(bx, list(sizes), order, 0.7, max_keep)
        assert torch.equal(n0, n1), (n0, n1)
        for b in range(B):
            assert torch.equal(k0[b, :int(n0[b])], k1[b, :int(n1[b])])
        assert int(n0.min()) > 10


def test_nms_large_vs_oracle():
    from bonai_b200.ops import batched_nms
    from oracle import ops_cpu
    g = torch.Generator().manual_seed(5)
    n = 6000
    c = torch.rand(n, 2, generator=g) * 1024
    wh = torch.exp(torch.rand(n, 2, generator=g) * 3) * 8
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024)
    scores = (torch.rand(n, generator=g) * 200).round() / 200
    ids = torch.randint(0, 5, (n,), generator=g)
    d0, k0 = ops_cpu.batched_nms(boxes, scores, ids, 0.7)
    d1, k1 = batched_nms(boxes.cuda(), scores.cuda(), ids.cuda(), dict(type='nms', iou_threshold=0.7))
    assert torch.equal(k1.cpu(), k0)
    assert torch.equal(d1.cpu(), d0)


def test_coders_and_offset_targets(golden_units):
    from bonai_b200.core import bbox2delta
    g = golden_units
    props, gts = torch.from_numpy(g['coder_props']).cuda(), torch.from_numpy(g['coder_gts']).cuda()
    d = bbox2delta(props, gts, (0., 0., 0., 0.), (0.1, 0.1, 0.2, 0.2))
    assert torch.allclose(d.cpu(), torch.from_numpy(g['coder_deltas']), rtol=2e-6, atol=1e-6)
    from bonai_b200.models.roi_heads.attribute_heads import OffsetHeadExpandFeature

    class SR:
        pos_bboxes = props
        pos_assigned_gt_inds = torch.from_numpy(g['foa_pos_inds']).cuda()

    head = OffsetHeadExpandFeature(expand_feature_num=4, share_expand_fc=True, num_convs=1,
                                   loss_offset=dict(type='SmoothL1Loss', loss_weight=16.0))
    t = head.get_targets([SR()], [torch.from_numpy(g['offset_gt']).cuda()], None)
    assert torch.allclose(t.cpu(), torch.from_numpy(g['foa_targets']), rtol=1e-5, atol=1e-6)


def test_rot90_matches_reference_rotation(golden_units):
    from bonai_b200.ops.misc import rot90
    g = golden_units
    feat = torch.from_numpy(g['foa_feat']).cuda().contiguous(memory_format=torch.channels_last)
    for i in range(4):
        x = feat.clone().requires_grad_(True)
        y = rot90(x, i)
        assert (y.cpu() - torch.from_numpy(g[f'foa_rot{i}'])).abs().max() < 2e-6
        dy = rnd(*y.shape, seed=i)
        y.backward(dy)
        assert torch.equal(x.grad, torch.rot90(dy, -i, (2, 3)))


# ----------------------------------------------------------------------------- losses
def test_elem_losses_and_ce():
    from bonai_b200.ops import losses as K
    P = 1000
    pred = rnd(P, 16, seed=1)
    t3, w3 = (rnd(P * 3, seed=2) > 0).float(), (rnd(P * 3, seed=3) > 0.5).float()
    p1 = pred.clone().requires_grad_(True)
    l1 = K.elem_loss(p1, t3, w3, K.BCE_LOGITS, 1 / 77.0, col_off=0, ncols=3)
    p2 = pred.clone().requires_grad_(True)
    l2 = (F.binary_cross_entropy_with_logits(p2[:, :3].reshape(-1), t3, reduction='none') *
          w3).sum() / 77.0
    assert abs(float(l1) - float(l2)) < 1e-5 * abs(float(l2))
    (l1 * 2.0).sum().backward()
    (l2 * 2.0).backward()
    assert rel(p1.grad, p2.grad) < 1e-5
    t12, w12 = rnd(P * 12, seed=4), (rnd(P * 12, seed=5) > 0).float()
    for mode, fn in ((K.L1, lambda a, b: (a - b).abs()),
                     (K.SMOOTH_L1, lambda a, b: F.smooth_l1_loss(a, b, reduction='none', beta=1.0))):
        p1 = pred.clone().requires_grad_(True)
        l1 = K.elem_loss(p1, t12, w12, mode, 0.01, col_off=3, ncols=12)
        p2 = pred.clone().requires_grad_(True)
        l2 = (fn(p2[:, 3:15].reshape(-1), t12) * w12).sum() * 0.01
        assert abs(float(l1) - float(l2)) < 1e-5 * abs(float(l2))
        l1.sum().backward()
        l2.backward()
        assert rel(p1.grad, p2.grad) < 1e-6
    logits = rnd(512, 8, seed=6)
    labels = torch.randint(0, 2, (512,), generator=torch.Generator().manual_seed(7)).cuda()
    w = (rnd(512, seed=8) > -1).float()
    a = logits.clone().requires_grad_(True)
    out = K.softmax_ce(a, labels, w, 2, 1 / 512.0)
    b = logits.clone().requires_grad_(True)
    ref = (F.cross_entropy(b[:, :2], labels, reduction='none') * w).sum() / 512.0
    assert abs(float(out[0]) - float(ref)) < 1e-5
    assert int(out[1]) == int((b[:, :2].argmax(1) == labels).sum())
    out[0].backward()
    ref.backward()
    assert rel(a.grad, b.grad) < 1e-5


def test_sigmoid_focal_loss(golden_units):
    from bonai_b200.ops import sigmoid_focal_loss
    from oracle import ops_cpu
    g = golden_units
    x = torch.from_numpy(g['focal_logits']).cuda().requires_grad_(True)
    t = torch.from_numpy(g['focal_target']).cuda()
    out = sigmoid_focal_loss(x, t, 2.0, 0.25, None, 'none')
    assert torch.allclose(out.cpu(), torch.from_numpy(g['focal_loss_none']), rtol=1e-5, atol=1e-6)
    out.sum().backward()
    xc = torch.from_numpy(g['focal_logits']).requires_grad_(True)
    ops_cpu.sigmoid_focal_loss(xc, torch.from_numpy(g['focal_target'])).sum().backward()
    assert torch.allclose(x.grad.cpu(), xc.grad, rtol=1e-4, atol=1e-6)


def test_sgd_clip_step_vs_oracle():
    import bonai_b200._lib as L
    from oracle import loft_cpu as O
    n = 100003
    p, g = rnd(n, seed=1), rnd(n, seed=2) * 3
    m = rnd(n, seed=3) * 0.1
    params, grads, bufs = {'w': p.cpu().clone()}, {'w': g.cpu().clone()}, {'w': m.cpu().clone()}
    O.sgd_step(params, grads, bufs, lr=0.01, max_norm=35.0)
    sq = torch.zeros(1, dtype=torch.float64, device='cuda')
    t = torch.empty_like(p)
    L.call('grad_sqnorm', L.ptr(g), L.ll(n), L.ptr(sq), L.stream())
    L.call('sgd_clip_step', L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(t), L.ll(n), L.f32(0.01), L.f32(0.9),
           L.f32(1e-4), L.f32(35.0), L.f32(1.0), L.ptr(sq), L.stream())
    assert torch.allclose(p.cpu(), params['w'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(m.cpu(), bufs['w'], rtol=1e-5, atol=1e-6)
    assert torch.equal(t, tf32_round(p))


def test_grouped_conv3x3_matches_per_group():
    """4 FOA-style branches (own weights) in one launch == 4 separate convs, fwd / dgrad / wgrad."""
    from bonai_b200.engine import WeightRef
    from bonai_b200.ops import dense as D
    G, Pg, C = 4, 13, 256
    wall = tf32_round(rnd(G, C + 1, 3, 3, C, seed=1, scale=0.03))       # +1 row of slack = stride
    gwall = torch.zeros_like(wall)
    ball = rnd(G, 260, seed=2)
    gball = torch.zeros_like(ball)
    wrefs = [WeightRef(wall[g, :C].permute(0, 3, 1, 2), gwall[g, :C].permute(0, 3, 1, 2))
             for g in range(G)]
    biases = [ball[g, :C] for g in range(G)]
    bgrads = [gball[g, :C] for g in range(G)]
    spec = D.GroupedConvSpec(wrefs, biases, bgrads, relu=True, store=_Store())
    assert spec.uniform
    x = tf32_round(rnd(G * Pg, C, 7, 7, seed=3)).contiguous(memory_format=torch.channels_last)
    xg = x.clone().requires_grad_(True)
    y = D.grouped_conv3x3(xg, spec)
    dy = tf32_round(rnd(*y.shape, seed=4))
    y.backward(dy)
    for g in range(G):
        xr = x[g * Pg:(g + 1) * Pg].clone().requires_grad_(True)
        wr = wrefs[g].w.clone().requires_grad_(True)
        br = biases[g].clone().requires_grad_(True)
        zr = F.conv2d(xr, wr, br, padding=1)
        yr = relu_like(zr, y[g * Pg:(g + 1) * Pg])
        assert rel(y[g * Pg:(g + 1) * Pg], F.relu(zr)) < TF32_TOL
        yr.backward(dy[g * Pg:(g + 1) * Pg])
        assert rel(xg.grad[g * Pg:(g + 1) * Pg], xr.grad) < GRAD_TOL
        assert rel(wrefs[g].grad, wr.grad) < GRAD_TOL
        assert rel(bgrads[g], br.grad) < GRAD_TOL


def test_take_rows_shares_the_gradient():
    """_TakeRows: (feats, feats[rows]) with the subset's gradient added in place into the rows of
    the full gradient == autograd over plain indexing."""
    from bonai_b200.ops.roi import take_rows
    K, C, S = 37, 8, 7
    x = rnd(K, C, S, S, seed=1).contiguous(memory_format=torch.channels_last).requires_grad_()
    xr = x.detach().clone().requires_grad_()
    rows = torch.tensor([0, 3, 4, 17, 36], device='cuda')
    full, sub = take_rows(x, rows)
    assert torch.equal(full, x) and torch.equal(sub, x[rows])
    wf, ws = rnd(K, C, S, S, seed=2), rnd(5, C, S, S, seed=3)
    ((full * wf).sum() + (sub * ws).sum()).backward()
    ((xr * wf).sum() + (xr[rows] * ws).sum()).backward()
    assert torch.allclose(x.grad, xr.grad, rtol=0, atol=1e-6)
    # only the subset is differentiated
    x.grad = None
    full, sub = take_rows(x, rows)
    (sub * ws).sum().backward()
    ref = torch.zeros_like(xr)
    ref[rows] = ws
    assert torch.allclose(x.grad, ref, rtol=0, atol=1e-6)


def test_take_rows_rot_equals_gather_rot90_cat():
    """_TakeRowsRot == (feats, cat_b rot90(feats[rows], k_b)) forward and backward."""
    from bonai_b200.ops.roi import take_rows_rot
    K, C, S = 23, 8, 7
    x = rnd(K, C, S, S, seed=1).contiguous(memory_format=torch.channels_last).requires_grad_()
    xr = x.detach().clone().requires_grad_()
    rows = torch.tensor([1, 2, 9, 22], device='cuda')
    ks = (0, 1, 2, 3)
    full, y = take_rows_rot(x, rows, ks)
    yr = torch.cat([torch.rot90(xr[rows], k, dims=(2, 3)) for k in ks], 0)
    assert torch.equal(full, x) and torch.equal(y, yr)
    wf, wy = rnd(K, C, S, S, seed=2), rnd(4 * rows.numel(), C, S, S, seed=3)
    ((full * wf).sum() + (y * wy).sum()).backward()
    ((xr * wf).sum() + (yr * wy).sum()).backward()
    assert torch.allclose(x.grad, xr.grad, rtol=0, atol=1e-5)


@pytest.mark.parametrize('C,n_out', [(128, 4), (256, 1), (256, 2), (512, 3)])
def test_narrow_head_forward_backward(C, n_out):
    """loft_narrow_head_fwd/bwd (<= 4 output channels, warp per pixel) against fp32 autograd:
    outputs, data gradient with the ReLU mask of the input, weight / bias gradients and the
    per-channel sum of the data gradient (the producer's bias gradient), all from one pass."""
    from bonai_b200.engine import WeightRef
    from bonai_b200.ops import dense as D
    N, H, W = 3, 27, 27                               # odd row count: the 2-row loop tail
    x = tf32_round(rnd(N, C, H, W, seed=1)).contiguous(memory_format=torch.channels_last)
    x = x.requires_grad_()
    w = tf32_round(rnd(4, C, seed=2, scale=0.1))
    w[n_out:] = 0                                     # padded rows
    b = rnd(4, seed=3)
    gw, gb, cs = torch.zeros_like(w), torch.zeros_like(b), torch.zeros(C, device='cuda')
    spec = D.ConvSpec(WeightRef(w, gw), ksize=1, bias=b, bias_grad=gb, round_out=False,
                      premask_in=True)
    spec.in_colsum = cs
    spec.n_out = n_out
    assert D.narrow_head_ok(spec, x)
    y = D.narrow_head(x, spec)
    xr = x.detach().clone().requires_grad_()
    wr, br = w.clone().requires_grad_(), b.clone().requires_grad_()
    yr = F.conv2d(xr, wr[:, :, None, None], br)
    assert rel(y, yr) < 1e-5
    dy = rnd(N, 4, H, W, seed=4).contiguous(memory_format=torch.channels_last)
    dy[:, n_out:] = 0              # the loss only differentiates the real outputs
    y.backward(dy)
    yr.backward(dy)
    dx_ref = xr.grad * (xr > 0)
    assert rel(x.grad, dx_ref) < TF32_TOL              # dx is rounded to TF32
    assert rel(gw[:n_out], wr.grad[:n_out]) < 1e-4 and bool((gw[n_out:] == 0).all())
    assert rel(gb, br.grad) < 1e-4
    assert rel(cs, x.grad.sum((0, 2, 3))) < 1e-4


def test_deconv_logits_fused_equals_separate_layers():
    """_DeconvLogitsFn (deconv output kept in its GEMM layout, narrow head addressed through the
    space-to-depth row map) == deconv2x2 + narrow head / 1x1 conv as separate layers: logits and
    every gradient."""
    from bonai_b200.engine import WeightRef
    from bonai_b200.ops import dense as D
    N, Ci, Co, H, W = 5, 64, 128, 6, 7

    def build():
        wu = tf32_round(rnd(4 * Co, Ci, seed=2, scale=0.1))
        bu = rnd(Co, seed=3, scale=0.1)
        b4 = bu.repeat(4).contiguous()
        wl = torch.zeros(4, Co, device='cuda')
        wl[0] = tf32_round(rnd(Co, seed=4, scale=0.1))
        bl = torch.zeros(4, device='cuda')
        bl[0] = 0.3
        g = dict(wu=torch.zeros_like(wu), bu=torch.zeros_like(bu), wl=torch.zeros_like(wl),
                 bl=torch.zeros_like(bl))
        up = D.ConvSpec(WeightRef(wu, g['wu']), relu=True, bias=b4, bias_grad=g['bu'],
                        grad_premasked=True)
        lg = D.ConvSpec(WeightRef(wl, g['wl']), ksize=1, bias=bl, bias_grad=g['bl'], round_out=False,
                        premask_in=True)
        lg.n_out = 1
        D.link_chain([up, lg])
        return up, lg, g

    x0 = tf32_round(rnd(N, Ci, H, W, seed=1)).contiguous(memory_format=torch.channels_last)
    dy = torch.zeros(N, 4, 2 * H, 2 * W, device='cuda').contiguous(memory_format=torch.channels_last)
    dy[:, 0] = tf32_round(rnd(N, 2 * H, 2 * W, seed=5))     # on the TF32 grid: exact in both paths
    up, lg, ga = build()
    xa = x0.clone().requires_grad_()
    assert D.deconv_logits_ok(up, lg, xa)
    ya = D.deconv_logits(xa, up, lg)
    ya.backward(dy)
    # (a) the separate layers of the product
    up2, lg2, gb = build()
    xb = x0.clone().requires_grad_()
    yb = D.conv(D.deconv2x2(xb, up2), lg2)
    yb.backward(dy)
    assert rel(ya[:, :1], yb[:, :1]) < 1e-5
    assert rel(xa.grad, xb.grad) < TF32_TOL
    for k in ('wu', 'bu', 'wl', 'bl'):
        ref = gb[k][:1] if k in ('wl', 'bl') else gb[k]
        got = ga[k][:1] if k in ('wl', 'bl') else ga[k]
        assert rel(got, ref) < GRAD_TOL, k
    # (b) torch fp32: ConvTranspose2d(2, 2) + ReLU (output rounded to TF32, straight through) + 1x1
    wu, bu, wl, bl = up.wref.w, up.bias[:Co], lg.wref.w, lg.bias
    wt = wu.view(2, 2, Co, Ci).permute(3, 2, 0, 1).contiguous().requires_grad_()   # [Ci,Co,i,j]
    bt, wlt, blt = bu.clone().requires_grad_(), wl[:1].clone().requires_grad_(), \
        bl[:1].clone().requires_grad_()
    xr = x0.clone().requires_grad_()
    h = F.relu(F.conv_transpose2d(xr, wt, bt, stride=2))
    h = h + (tf32_round(h.detach()) - h.detach())
    yr = F.conv2d(h, wlt[:, :, None, None], blt)
    yr.backward(dy[:, :1])
    assert rel(ya[:, :1], yr) < TF32_TOL
    assert rel(xa.grad, xr.grad) < 2e-3               # dz is rounded to TF32 before the dgrad GEMM
    assert rel(ga['wu'].view(2, 2, Co, Ci).permute(3, 2, 0, 1), wt.grad) < GRAD_TOL
    assert rel(ga['bu'], bt.grad) < GRAD_TOL
    assert rel(ga['wl'][:1], wlt.grad) < GRAD_TOL and rel(ga['bl'][:1], blt.grad) < GRAD_TOL
