"""Whole-step GPU parity (teacher-forced against the CPU oracle and against the reference's golden
losses) and size-independent properties at BASELINE.json's full size (1024x1024, batch 2, G=80)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))

LOSS_TOL = 1e-3          # BASELINE.json north_star: losses within 1e-3 relative


def test_step_parity_256_vs_oracle_and_reference_golden(golden_step):
    import e2e_check
    rep = e2e_check.run(size=256, n_img=1, num_gt=10, seed=0, verbose=False)
    assert rep['worst_loss_rel'] < LOSS_TOL, rep['losses']
    gold = dict(zip([str(n) for n in golden_step['loss_names']], golden_step['loss_values']))
    for k, (gpu, oracle, _) in rep['losses'].items():
        assert abs(gpu - gold[k]) <= LOSS_TOL * max(abs(gold[k]), 1e-6), (k, gpu, gold[k])
    assert 'missing' not in rep['grads'].values()
    assert rep['worst_grad_rel'][0] < 0.15, rep['worst_grad_rel']


def test_step_parity_two_images_batched_nms():
    import e2e_check
    rep = e2e_check.run(size=256, n_img=2, num_gt=7, seed=3, force_proposals=True, verbose=False)
    assert rep['worst_loss_rel'] < LOSS_TOL, rep['losses']


@pytest.fixture(scope='module')
def full_size():
    from bonai_b200 import Config
    from bonai_b200.apis import Trainer
    from bonai_b200.core import BitmapMasks
    from bonai_b200.models import build_detector
    from oracle import loft_cpu as O
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.train()
    trainer = Trainer(model, cfg, torch.device('cuda:0'))
    img, gb, gl, gm, go = O.make_inputs(0, 2, 1024, 80)
    dev = 'cuda:0'
    metas = [dict(img_shape=(1024, 1024, 3), pad_shape=(1024, 1024, 3), scale_factor=1.0,
                  flip=False)] * 2
    data = dict(img=img.to(dev), img_metas=metas, gt_bboxes=[b.to(dev) for b in gb],
                gt_labels=[l.to(dev) for l in gl],
                gt_masks=[BitmapMasks(m.to(dev), 1024, 1024) for m in gm],
                gt_offsets=[o.to(dev) for o in go])
    return model, trainer, data


def test_full_size_step_properties(full_size):
    model, trainer, data = full_size
    logs0 = trainer.train_step(data, read_logs=True)
    assert all(torch.isfinite(torch.tensor(v)) for v in logs0.values()), logs0
    # at random init: RPN cls ~ ln 2, RCNN cls ~ ln 2 (two classes), losses positive
    assert abs(logs0['loss_rpn_cls'] - 0.6931) < 0.02 and abs(logs0['loss_cls'] - 0.6931) < 0.15
    assert logs0['loss_offset'] > 0 and logs0['loss_mask'] > 0
    G = trainer.store.G
    assert bool(torch.isfinite(G).all())
    assert float(trainer.store.grad_norm()) > 0
    srs = model.roi_head._last_sampling_results
    for sr in srs:
        n = sr.pos_bboxes.shape[0] + sr.neg_bboxes.shape[0]
        assert n == 1024 and 80 <= sr.pos_bboxes.shape[0] <= 256     # R = 1024, GT added
    for _ in range(3):
        logs = trainer.train_step(data, read_logs=True)
    assert logs['loss'] < logs0['loss']                              # SGD on a fixed batch descends


def test_full_size_proposals_properties(full_size):
    from bonai_b200.ops import batched_nms
    model, trainer, data = full_size
    with torch.no_grad():
        trainer.store.refresh_weights()
        feats = model.extract_feat(data['img'])
        outs = model.rpn_head(feats)
        props = model.rpn_head.get_bboxes(*outs, data['img_metas'], cfg=model.train_cfg.rpn_proposal)
    for p in props:
        assert p.shape[0] <= 3000 and p.shape[1] == 5
        assert bool((p[1:, 4] <= p[:-1, 4]).all())                   # score-descending
        assert float(p[:, :4].min()) >= 0 and float(p[:, :4].max()) <= 1024
        assert bool((p[:, 2] >= p[:, 0]).all()) and bool((p[:, 3] >= p[:, 1]).all())
        # idempotence: NMS of the survivors (same threshold, one class) removes nothing within a
        # level; run it class-agnostically on the boxes of one level-sized chunk
        sub = p[:500]
        ids = torch.zeros(sub.shape[0], dtype=torch.long, device=p.device)
        d1, k1 = batched_nms(sub[:, :4].contiguous(), sub[:, 4].contiguous(), ids,
                             dict(type='nms', iou_threshold=0.7))
        d2, k2 = batched_nms(d1[:, :4].contiguous(), d1[:, 4].contiguous(), ids[:d1.shape[0]],
                             dict(type='nms', iou_threshold=0.7))
        assert d2.shape[0] == d1.shape[0] and torch.equal(d2, d1)


def test_full_size_conv_linearity():
    """conv(a + b) == conv(a) + conv(b) on the dominant P2-level shape (TF32-rounded inputs)."""
    import ctypes
    from bonai_b200 import _lib as L
    N, H, W, C = 2, 256, 256, 256
    g = torch.Generator().manual_seed(0)

    def r(*s):
        t = torch.randn(*s, generator=g).cuda()
        return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)   # TF32 grid
    a, b = r(N, H, W, C) * 0.5, r(N, H, W, C) * 0.5
    ab = (a + b)
    ab = ((ab.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    a = ab - b                                      # exact on the grid up to one rounding
    w = r(C, 3, 3, C) * 0.02
    outs = []
    for x in (a, b, ab):
        y = torch.empty(N, H, W, C, device='cuda')
        L.call('conv3x3_fprop', L.ptr(x.contiguous()), L.ptr(w), L.ptr(y), ctypes.c_int(N),
               ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(C), ctypes.c_int(C), None, L.stream())
        outs.append(y)
    torch.cuda.synchronize()
    err = (outs[0] + outs[1] - outs[2]).norm() / outs[2].norm()
    assert float(err) < 2e-3
