"""Pins the CPU oracle (oracle/ops_cpu.py, oracle/loft_cpu.py) against golden vectors produced by
the reference's own code (oracle/make_golden.py) and against the reference's docstring/tests
known answers."""
import numpy as np
import torch

from oracle import loft_cpu as O
from oracle import ops_cpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_anchor_generator_matches_reference(golden_units):
    sizes = [tuple(int(v) for v in s) for s in golden_units['anchor_sizes']]
    for i, a in enumerate(O.grid_anchors(sizes)):
        assert torch.equal(a, T(golden_units[f'anchors_l{i}']))


def test_anchor_known_answer():
    # tests/test_anchor.py:22-40 of the reference (strides [4], ratios [1.], scales [1.], 2x2 map)
    cfg = dict(O.CFG, anchor_strides=[4], anchor_ratios=[1.0], anchor_scales=[1.0])
    a = O.grid_anchors([(2, 2)], cfg)[0]
    exp = torch.tensor([[-2., -2., 2., 2.], [2., -2., 6., 2.], [-2., 2., 2., 6.], [2., 2., 6., 6.]])
    assert torch.equal(a, exp)


def test_bbox_overlaps_and_assigner(golden_units):
    b1, b2 = T(golden_units['iou_b1']), T(golden_units['iou_b2'])
    assert torch.equal(O.bbox_overlaps(b2, b1), T(golden_units['iou']))
    for name, (pos, neg, mn) in dict(rpn=(0.7, 0.3, 0.3), rcnn=(0.5, 0.5, 0.5)).items():
        gi, mo, lab = O.max_iou_assign(b1, b2, pos, neg, mn, torch.zeros(7, dtype=torch.long))
        assert torch.equal(gi, T(golden_units[f'assign_{name}_gt_inds']))
        assert torch.equal(mo, T(golden_units[f'assign_{name}_max_overlaps']))
        assert torch.equal(lab, T(golden_units[f'assign_{name}_labels']))


def test_assigner_known_answers():
    # tests/test_assigner.py:14-36 of the reference -> gt_inds [1, 0, 2, 0]
    bboxes = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gts = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    gi, _, lab = O.max_iou_assign(bboxes, gts, 0.5, 0.5, 0.0, torch.LongTensor([2, 3]))  # min_pos_iou default .0
    assert gi.tolist() == [1, 0, 2, 0]
    assert lab.tolist() == [2, -1, 3, -1]
    # empty gts (tests/test_assigner.py:65-93): everything background
    gi, _, _ = O.max_iou_assign(bboxes, torch.empty(0, 4), 0.5, 0.5, 0.5)
    assert gi.tolist() == [0, 0, 0, 0]
    # empty boxes
    gi, _, _ = O.max_iou_assign(torch.empty(0, 4), gts, 0.5, 0.5, 0.5)
    assert gi.numel() == 0


def test_delta2bbox_docstring():
    # delta_xywh_bbox_coder.py:149-162 of the reference
    rois = torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]])
    deltas = torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.],
                           [0.7, -1.9, -0.5, 0.3]])
    out = O.delta2bbox(rois, deltas, (1., 1., 1., 1.), max_shape=(32, 32))
    exp = torch.tensor([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                        [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]])
    assert torch.allclose(out, exp, atol=1e-4)


def test_coders_match_reference(golden_units):
    g = golden_units
    props, gts = T(g['coder_props']), T(g['coder_gts'])
    assert torch.equal(O.bbox2delta(props, gts, O.CFG['rcnn_stds']), T(g['coder_deltas']))
    dec = O.delta2bbox(props, T(g['coder_big_deltas']), O.CFG['rcnn_stds'], max_shape=(100, 120))
    assert torch.equal(dec, T(g['coder_decoded']))
    assert torch.equal(O.offset2delta(props, T(g['offset_gt'])), T(g['offset_encoded']))
    assert torch.equal(O.delta2offset(props, T(g['offset_deltas']), max_shape=[1024, 1024]),
                       T(g['offset_decoded']))


def test_foa_targets_rotation_fusion_loss(golden_units):
    g = golden_units
    sr = dict(pos_bboxes=T(g['coder_props']), pos_assigned_gt_inds=T(g['foa_pos_inds']))
    t = O.offset_targets([sr], [T(g['offset_gt'])])
    assert torch.equal(t, T(g['foa_targets']))
    feat = T(g['foa_feat'])
    for i, rot in enumerate(O.CFG['rotations']):
        r = O.rotate_feature(feat, rot)
        assert torch.equal(r, T(g[f'foa_rot{i}']))
        # the rotation is a rot90 permutation to fp noise
        assert (r - torch.rot90(feat, i, (2, 3))).abs().max() < 2e-6
    assert torch.equal(O.offset_fusion_max(T(g['foa_pred'])), T(g['foa_fused']))
    loss = 16.0 * O.smooth_l1_mean(T(g['foa_loss_pred']), T(g['foa_loss_target']))
    assert abs(float(loss) - float(g['foa_loss'])) < 1e-5


def test_foa_target_closed_form(golden_units):
    """Closed form used by the CUDA kernel: (x/pw, y/ph)*2, (y/ph,-x/pw)*2, (-x/pw,-y/ph)*2,
    (-y/ph, x/pw)*2 -- equals the reference's polar/float64 path to fp32 noise."""
    g = golden_units
    props, offs, inds = T(g['coder_props']), T(g['offset_gt']), T(g['foa_pos_inds'])
    o = offs[inds]
    pw, ph = props[:, 2] - props[:, 0], props[:, 3] - props[:, 1]
    x, y = o[:, 0] / pw * 2, o[:, 1] / ph * 2
    exp = torch.cat([torch.stack(v, 1) for v in [(x, y), (y, -x), (-x, -y), (-y, x)]])
    assert torch.allclose(exp, T(g['foa_targets']), rtol=1e-5, atol=1e-6)


def test_focal_loss(golden_units):
    g = golden_units
    out = ops_cpu.sigmoid_focal_loss(T(g['focal_logits']), T(g['focal_target']))
    assert torch.allclose(out, T(g['focal_loss_none']), rtol=1e-6, atol=1e-7)


def test_roi_align_matches_torchvision_golden(golden_units):
    g = golden_units
    fm, rois = T(g['ra_feat']), T(g['ra_rois'])
    for key, (s, scale) in dict(ra_out7_s4=(7, 0.25), ra_out14_s8=(14, 0.125),
                                ra_out28_s1=(28, 1.0)).items():
        out = ops_cpu.roi_align(fm, rois, s, scale, 0, True)
        assert torch.allclose(out, T(g[key]), rtol=1e-5, atol=1e-6), key


def test_roi_align_backward_matches_torchvision():
    import torchvision.ops as tvo
    g = torch.Generator().manual_seed(3)
    fm = torch.randn(2, 4, 16, 16, generator=g, dtype=torch.float64)
    rois = torch.tensor([[0, 1.3, 2.2, 30.7, 41.9], [1, -5.0, -3.0, 12.0, 9.0],
                         [1, 0.0, 0.0, 64.0, 64.0], [0, 10.2, 10.7, 11.1, 11.9]],
                        dtype=torch.float64)
    w = torch.randn(4, 4, 7, 7, generator=g, dtype=torch.float64)
    a = fm.clone().requires_grad_(True)
    b = fm.clone().requires_grad_(True)
    (ops_cpu.roi_align(a, rois, 7, 0.25, 0, True) * w).sum().backward()
    (tvo.roi_align(b, rois, (7, 7), 0.25, 0, True) * w).sum().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-10, atol=1e-12)


def test_nms_bit_exact(golden_units):
    g = golden_units
    boxes, scores, ids = T(g['nms_boxes']), T(g['nms_scores']), T(g['nms_ids'])
    _, keep = ops_cpu.nms(boxes, scores, 0.7)
    assert torch.equal(keep, T(g['nms_keep']))
    dets, keep = ops_cpu.batched_nms(boxes, scores, ids, 0.7)
    assert torch.equal(keep, T(g['bnms_keep']))
    assert torch.equal(dets, T(g['bnms_dets']))


def test_nms_edge_cases():
    _, keep = ops_cpu.nms(torch.zeros(0, 4), torch.zeros(0), 0.5)
    assert keep.numel() == 0
    b = torch.tensor([[0., 0., 10., 10.]] * 3)
    _, keep = ops_cpu.nms(b, torch.tensor([0.5, 0.5, 0.5]), 0.5)
    assert keep.tolist() == [0]          # ties keep input order (stable)


def test_cross_entropy_known_answer():
    # tests/test_models/test_losses.py:18-31 of the reference: CE([[0,100]],[1]) = 0 ... uses
    # class_weight; here the plain values the path relies on: CE([[100,-100]],[1]) = 200
    import torch.nn.functional as F
    assert float(F.cross_entropy(torch.Tensor([[100., -100.]]), torch.LongTensor([1]))) == 200.


def test_level_by_level_nms_equals_batched_nms():
    """The product runs the RPN's batched NMS level by level (loft_nms_segmented): boxes of
    different levels are pushed (max+1) apart by batched_nms, so they never suppress each other.
    Restated with the oracle on CPU: per-level NMS on the SAME offset coordinates, kept boxes merged
    back in global score order and cut to max_keep, equals the oracle's batched_nms (itself pinned
    to the reference's golden vectors above), ties included."""
    import torch
    from oracle import ops_cpu as O
    g = torch.Generator().manual_seed(11)
    sizes = (300, 300, 120, 33, 1)
    ctr = torch.rand(12, 2, generator=g) * 400 + 50
    boxes, scores, ids = [], [], []
    for l, k in enumerate(sizes):
        c = ctr[torch.randint(0, 12, (k,), generator=g)] + torch.randn(k, 2, generator=g) * 5
        wh = torch.rand(k, 2, generator=g) * 50 + 6
        boxes.append(torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 512))
        scores.append((torch.rand(k, generator=g) * 40).round() / 40)        # many exact ties
        ids.append(torch.full((k,), l, dtype=torch.long))
    boxes, scores, ids = torch.cat(boxes), torch.cat(scores), torch.cat(ids)
    _, keep_ref = O.batched_nms(boxes, scores, ids, 0.7)
    # level by level, with batched_nms's fp32 offsets
    off = ids.to(boxes) * (boxes.max() + 1)
    shifted = boxes + off[:, None]
    kept = torch.zeros(boxes.shape[0], dtype=torch.bool)
    for l in range(len(sizes)):
        sel = torch.nonzero(ids == l)[:, 0]
        _, k = O.nms(shifted[sel], scores[sel], 0.7)
        kept[sel[k]] = True
    order = torch.sort(scores, descending=True, stable=True)[1]
    merged = order[kept[order]]
    assert torch.equal(merged, keep_ref)
    assert torch.equal(merged[:100], keep_ref[:100])                          # max_keep cut


def test_bn_fold_identities():
    """The product folds eval-mode BN into the conv weights and recovers the BN gradients from the
    weight gradient (loft_bn_fold_weights / loft_bn_finalize).  The algebra, checked
    against torch autograd in float64: with W' = s*W, shift = beta - mean*s, s = gamma*rstd,
      dW = s * dW',   dbeta = sum_p dY,   dgamma = rstd * (sum_k W*dW' - mean*dbeta)."""
    import torch
    import torch.nn.functional as F
    torch.manual_seed(3)
    x = torch.randn(3, 8, 9, 9, dtype=torch.float64)
    w = torch.randn(12, 8, 3, 3, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(12, dtype=torch.float64) + 0.3).requires_grad_(True)
    gamma.data[5] = 0.0                                   # zero_init_residual corner
    beta = torch.randn(12, dtype=torch.float64, requires_grad=True)
    mean, var = torch.randn(12, dtype=torch.float64), torch.rand(12, dtype=torch.float64) + 0.5
    y = F.relu(F.batch_norm(F.conv2d(x, w, padding=1), mean, var, gamma, beta, False, 0.0, 1e-5))
    dy = torch.randn_like(y)
    y.backward(dy)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    s = (gamma * rstd).detach()
    wf = (w.detach() * s[:, None, None, None]).requires_grad_(True)       # folded weights
    shift = (beta - mean * gamma * rstd).detach().requires_grad_(True)
    y2 = F.relu(F.conv2d(x, wf, padding=1) + shift[None, :, None, None])
    assert torch.allclose(y2, y.detach(), atol=1e-12)
    y2.backward(dy)
    dwf, dbeta = wf.grad, shift.grad
    assert torch.allclose(dbeta, beta.grad, atol=1e-10)
    assert torch.allclose(dwf * s[:, None, None, None], w.grad, atol=1e-10)
    dgamma = rstd * ((w.detach() * dwf).sum((1, 2, 3)) - mean * dbeta)
    assert torch.allclose(dgamma, gamma.grad, atol=1e-9)
    assert gamma.grad[5].abs() > 0 and w.grad[5].abs().max() == 0       # gamma=0: dW=0, dgamma!=0
