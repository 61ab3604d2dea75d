"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) over the
import shim.  Runs only in the build container; the fixtures are committed so the GPU box (which
has no /root/reference) can still pin the oracle and the CUDA path against reference outputs.

    python -m oracle.make_golden            # from the repo root
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import loft_cpu as O  # noqa: E402
from oracle import ref_env  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
CONFIG = 'configs/loft_foa/loft_foa_r50_fpn_2x_bonai.py'


def build_reference_model(state=None):
    ref_env.activate()
    from mmcv import Config
    from mmdet.models import build_detector
    cfg = Config.fromfile(os.path.join(ref_env.REF_ROOT, CONFIG))
    cfg.model.pretrained = None
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    if state is not None:
        model.load_state_dict(state, strict=True)
    model.train()
    return model, cfg


def reference_step(model, img, gb, gl, gm, go, seed):
    from mmdet.core import BitmapMasks
    h, w = img.shape[-2:]
    metas = [dict(img_shape=(h, w, 3), ori_shape=(h, w, 3), pad_shape=(h, w, 3), scale_factor=1.0,
                  flip=False, filename='synthetic') for _ in range(img.size(0))]
    torch.manual_seed(seed)
    model.zero_grad()
    losses = model.forward_train(img, metas, gb, gl,
                                 gt_masks=[BitmapMasks(m.numpy(), h, w) for m in gm],
                                 gt_offsets=go)
    loss, log_vars = model._parse_losses(losses)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
             if p.grad is not None}
    return {k: float(v) for k, v in log_vars.items()}, grads


def gold_loft_step():
    """Whole-step fixture: reference losses + per-parameter gradient norms on the portable init."""
    p = O.randomize_bn(O.init_params(0), 0)
    model, _ = build_reference_model(p)
    img, gb, gl, gm, go = O.make_inputs(0, 1, 256, 10)
    log_vars, grads = reference_step(model, img, gb, gl, gm, go, seed=123)
    keys = sorted(grads.keys())
    np.savez_compressed(
        os.path.join(GOLD, 'loft_step_256.npz'),
        loss_names=np.array(list(log_vars.keys())),
        loss_values=np.array(list(log_vars.values()), dtype=np.float64),
        grad_names=np.array(keys),
        grad_norms=np.array([float(grads[k].double().norm()) for k in keys]),
        grad_sums=np.array([float(grads[k].double().sum()) for k in keys]),
        meta=np.array(['init_params(0)+randomize_bn(0); make_inputs(0,1,256,10); '
                       'torch.manual_seed(123); torch ' + torch.__version__]))
    print('loft_step_256:', log_vars)


def gold_units():
    """Known answers of the BONAI-specific / path functions, produced by the reference's own code."""
    ref_env.activate()
    from mmdet.core.anchor import AnchorGenerator
    from mmdet.core.bbox import MaxIoUAssigner, DeltaXYWHBBoxCoder
    from mmdet.core.bbox.coder import DeltaXYOffsetCoder
    from mmdet.core.bbox.iou_calculators import bbox_overlaps
    from mmdet.models.losses.focal_loss import py_sigmoid_focal_loss
    from mmdet.models.roi_heads.attribute_heads import OffsetHeadExpandFeature
    import torchvision.ops as tvo
    g = torch.Generator().manual_seed(7)
    out = {}
    # anchors (config of bonai_loft_foa_r50_fpn_basic.py:23-27) on a 2-level toy pyramid
    ag = AnchorGenerator(strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8])
    sizes = [(6, 5), (3, 3), (2, 2), (1, 1), (1, 1)]
    for i, a in enumerate(ag.grid_anchors(sizes, device='cpu')):
        out[f'anchors_l{i}'] = a.numpy()
    out['anchor_sizes'] = np.array(sizes)
    # IoU / assigner
    b1 = torch.rand(40, 4, generator=g) * 100
    b1[:, 2:] = b1[:, :2] + torch.rand(40, 2, generator=g) * 60 + 1
    b2 = torch.rand(7, 4, generator=g) * 100
    b2[:, 2:] = b2[:, :2] + torch.rand(7, 2, generator=g) * 60 + 1
    b1[5] = b2[2]
    b1[6] = b2[2]
    out['iou_b1'], out['iou_b2'] = b1.numpy(), b2.numpy()
    out['iou'] = bbox_overlaps(b2, b1).numpy()
    for name, (pos, neg, mn) in dict(rpn=(0.7, 0.3, 0.3), rcnn=(0.5, 0.5, 0.5)).items():
        asg = MaxIoUAssigner(pos_iou_thr=pos, neg_iou_thr=neg, min_pos_iou=mn,
                             match_low_quality=True, ignore_iof_thr=-1)
        r = asg.assign(b1, b2, gt_labels=torch.zeros(7, dtype=torch.long))
        out[f'assign_{name}_gt_inds'] = r.gt_inds.numpy()
        out[f'assign_{name}_max_overlaps'] = r.max_overlaps.numpy()
        out[f'assign_{name}_labels'] = r.labels.numpy()
    # coders
    coder = DeltaXYWHBBoxCoder(target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2])
    props, gts = b1[:7], b2
    d = coder.encode(props, gts)
    out['coder_props'], out['coder_gts'], out['coder_deltas'] = props.numpy(), gts.numpy(), d.numpy()
    big = torch.randn(7, 4, generator=g) * 3
    out['coder_big_deltas'] = big.numpy()
    out['coder_decoded'] = coder.decode(props, big, max_shape=(100, 120)).numpy()
    oc = DeltaXYOffsetCoder()
    offs = torch.rand(7, 2, generator=g) * 80 - 40
    out['offset_gt'] = offs.numpy()
    out['offset_encoded'] = oc.encode(props, offs).numpy()
    od = torch.randn(7, 2, generator=g)
    out['offset_deltas'] = od.numpy()
    out['offset_decoded'] = oc.decode(props, od, max_shape=[1024, 1024]).numpy()
    # FOA head: targets, rotation, fusion
    head = OffsetHeadExpandFeature(expand_feature_num=4, share_expand_fc=True,
                                   rotations=[0, 90, 180, 270], num_fcs=2, fc_out_channels=1024,
                                   num_convs=10,
                                   loss_offset=dict(type='SmoothL1Loss', loss_weight=16.0))

    class SR:
        pass

    sr = SR()
    sr.pos_bboxes = props
    sr.pos_assigned_gt_inds = torch.tensor([3, 1, 0, 6, 2, 2, 5])
    out['foa_pos_inds'] = sr.pos_assigned_gt_inds.numpy()
    out['foa_targets'] = head.get_targets([sr], [offs], None).numpy()
    feat = torch.randn(3, 4, 7, 7, generator=g)
    out['foa_feat'] = feat.numpy()
    for i in range(4):
        out[f'foa_rot{i}'] = head.expand_feature(feat, i).numpy()
    pred = torch.randn(4 * 6, 2, generator=g)
    out['foa_pred'] = pred.numpy()
    out['foa_fused'] = head.offset_fusion(pred).numpy()
    pt = torch.randn(24, 2, generator=g) * 2
    out['foa_loss_pred'], out['foa_loss_target'] = pred.numpy(), pt.numpy()
    out['foa_loss'] = np.array(float(head.loss(pred, pt)['loss_offset']))
    # focal loss (reference's pure-PyTorch twin of the mmcv CUDA op)
    logits = torch.randn(50, 3, generator=g) * 2
    tgt = torch.randint(0, 4, (50,), generator=g)
    onehot = torch.nn.functional.one_hot(tgt, 4)[:, :3].float()
    out['focal_logits'], out['focal_target'] = logits.numpy(), tgt.numpy()
    out['focal_loss_none'] = py_sigmoid_focal_loss(logits, onehot, reduction='none').numpy()
    # RoIAlign / NMS through the implementation the reference switches to on CPU (torchvision)
    fm = torch.randn(2, 3, 20, 24, generator=g)
    rois = torch.tensor([[0, 1.3, 2.2, 30.7, 41.9], [1, -5.0, -3.0, 12.0, 9.0],
                         [0, 40.0, 30.0, 40.0, 30.0], [1, 0.0, 0.0, 96.0, 80.0],
                         [0, 90.0, 70.0, 130.0, 120.0], [1, 10.2, 10.7, 11.1, 11.9]])
    out['ra_feat'], out['ra_rois'] = fm.numpy(), rois.numpy()
    out['ra_out7_s4'] = tvo.roi_align(fm, rois, (7, 7), 0.25, 0, True).numpy()
    out['ra_out14_s8'] = tvo.roi_align(fm, rois, (14, 14), 0.125, 0, True).numpy()
    out['ra_out28_s1'] = tvo.roi_align(fm, rois, (28, 28), 1.0, 0, True).numpy()
    nb = torch.rand(300, 4, generator=g) * 200
    nb[:, 2:] = nb[:, :2] + torch.rand(300, 2, generator=g) * 80 + 1
    ns = (torch.rand(300, generator=g) * 20).round() / 20        # many exact ties
    ids = torch.randint(0, 5, (300,), generator=g)
    out['nms_boxes'], out['nms_scores'], out['nms_ids'] = nb.numpy(), ns.numpy(), ids.numpy()
    out['nms_keep'] = tvo.nms(nb, ns, 0.7).numpy()
    from mmcv.ops import batched_nms
    dets, keep = batched_nms(nb, ns, ids, dict(type='nms', iou_threshold=0.7))
    out['bnms_keep'], out['bnms_dets'] = keep.numpy(), dets.numpy()
    np.savez_compressed(os.path.join(GOLD, 'units.npz'), **out)
    print('units:', len(out), 'arrays')


def gold_inference():
    """Reference simple_test (bbox, segm, offset results) on the portable init, one 256^2 tile."""
    p = O.randomize_bn(O.init_params(0), 0)
    model, _ = build_reference_model(p)
    model.eval()
    img, _, _, _, _ = O.make_inputs(0, 1, 256, 10)
    metas = [dict(img_shape=(256, 256, 3), ori_shape=(256, 256, 3), pad_shape=(256, 256, 3),
                  scale_factor=1.0, flip=False, filename='synthetic')]
    with torch.no_grad():
        bbox_r, segm_r, off_r = model.simple_test(img, metas, rescale=False)
    dets = bbox_r[0]
    areas = np.array([int(m.sum()) for m in segm_r[0]], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, 'loft_infer_256.npz'), dets=dets,
                        offsets=np.asarray(off_r, dtype=np.float32), mask_areas=areas,
                        meta=np.array(['init_params(0)+randomize_bn(0); make_inputs(0,1,256,10) img; '
                                       'reference simple_test(rescale=False); torch ' +
                                       torch.__version__]))
    print('loft_infer_256:', dets.shape, np.asarray(off_r).shape, areas[:5])


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ['units', 'step', 'infer']
    if 'units' in which:
        gold_units()
    if 'step' in which:
        gold_loft_step()
    if 'infer' in which:
        gold_inference()
