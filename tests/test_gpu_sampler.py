"""The one-launch RCNN sampler (`loft_rcnn_sample`, product path) against the ATen formulation of
BaseSampler.sample + RandomSampler (base_sampler.py:34-101, random_sampler.py:31-75): identical
candidate sets, counts and every deterministic field; the random subsets are valid, ordered and
uniform (their random stream is the kernel's own -- whole-step parity tests inject the oracle's
draws and therefore run the ATen formulation)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _head():
    from bonai_b200 import Config
    from bonai_b200.models import build_detector
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    return model.roi_head


def _scene(seed, n_img=2, K=2000, G=(80, 57), crowded=False, dev='cuda'):
    g = torch.Generator().manual_seed(seed)
    props, gts, labels = [], [], []
    block = torch.zeros(n_img, K, 5)
    for i in range(n_img):
        c = torch.rand(G[i], 2, generator=g) * 900 + 60
        wh = torch.exp(torch.rand(G[i], 2, generator=g) * 2.0) * 12
        gt = torch.cat([c - wh / 2, c + wh / 2], 1)
        if crowded:      # most proposals are jittered gt boxes: far more than 128 positives
            src = gt[torch.randint(0, G[i], (K,), generator=g)]
            p = src + torch.randn(K, 4, generator=g) * 1.5
        else:
            pc = torch.rand(K, 2, generator=g) * 1000
            pwh = torch.exp(torch.rand(K, 2, generator=g) * 2.5) * 10
            p = torch.cat([pc - pwh / 2, pc + pwh / 2], 1)
            p[: G[i] * 2] = gt.repeat(2, 1) + torch.randn(G[i] * 2, 4, generator=g) * 2.0
        block[i, :, :4] = p.clamp(0, 1024)
        block[i, :, 4] = torch.rand(K, generator=g)
        gts.append(gt.to(dev))
        labels.append(torch.zeros(G[i], dtype=torch.long, device=dev))
    block = block.to(dev)
    nv = torch.tensor([K, K - 237], dtype=torch.int32, device=dev)[:n_img]
    for i in range(n_img):
        block[i, int(nv[i]):] = 0
        r = block[i]
        r._loft_num_valid = nv[i]
        props.append(r)
    return props, gts, labels, nv


def _run(head, props, gts, labels, fused):
    os.environ['LOFT_FUSED_SAMPLER'] = '1' if fused else '0'
    try:
        return head.assign_and_sample(None, [{}] * len(props), props, gts, labels)
    finally:
        os.environ.pop('LOFT_FUSED_SAMPLER', None)


@pytest.mark.gpu
@pytest.mark.parametrize('crowded', [False, True])
def test_fused_sampler_fields_and_counts(crowded):
    head = _head()
    props, gts, labels, nv = _scene(1 + int(crowded), crowded=crowded)
    ref = _run(head, props, gts, labels, fused=False)
    got = _run(head, props, gts, labels, fused=True)
    num = head.bbox_sampler.num
    max_pos = int(num * head.bbox_sampler.pos_fraction)       # LOFT: 1024 RoIs, 256 positives
    for i, (a, b) in enumerate(zip(ref, got)):
        G, K = gts[i].shape[0], props[i].shape[0]
        allb = torch.cat([gts[i], props[i][:, :4]])
        ar = head.bbox_assigner.assign(props[i], gts[i], None, labels[i])
        gi = torch.cat([torch.arange(1, G + 1, device='cuda'), ar.gt_inds])
        gi[G + int(nv[i]):] = -1
        n_pos_c, n_neg_c = int((gi > 0).sum()), int((gi == 0).sum())
        assert b.pos_inds.numel() == a.pos_inds.numel() == min(n_pos_c, max_pos)
        assert b.neg_inds.numel() == a.neg_inds.numel() == min(n_neg_c, num - b.pos_inds.numel())
        if crowded:
            assert n_pos_c > max_pos
        for inds, want in ((b.pos_inds, gi > 0), (b.neg_inds, gi == 0)):
            if inds.numel() == 0:
                continue
            assert bool(want[inds].all())                               # drawn from the right set
            assert bool((inds[1:] > inds[:-1]).all())                   # ascending, no duplicates
            assert int(inds.max()) < G + int(nv[i])                     # never a padding row
        assert torch.equal(b.pos_bboxes, allb[b.pos_inds]) and torch.equal(b.neg_bboxes, allb[b.neg_inds])
        assert torch.equal(b.bboxes, torch.cat([b.pos_bboxes, b.neg_bboxes]))
        assert torch.equal(b.pos_assigned_gt_inds, gi[b.pos_inds] - 1)
        assert torch.equal(b.pos_gt_bboxes, gts[i][b.pos_assigned_gt_inds])
        assert torch.equal(b.pos_is_gt.bool(), b.pos_inds < G)
        assert torch.equal(b.pos_gt_labels, labels[i][b.pos_assigned_gt_inds]) and b.num_gts == G
        if not crowded:       # every positive is taken: identical to the ATen formulation
            assert torch.equal(b.pos_inds, a.pos_inds) and torch.equal(b.pos_bboxes, a.pos_bboxes)
            assert torch.equal(b.pos_assigned_gt_inds, a.pos_assigned_gt_inds)
            assert torch.equal(b.pos_gt_bboxes, a.pos_gt_bboxes)
            assert torch.equal(b.pos_is_gt, a.pos_is_gt)


@pytest.mark.gpu
def test_fused_sampler_draws_are_uniform_and_vary():
    head = _head()
    props, gts, labels, nv = _scene(5, n_img=1, K=2000, G=(20,))     # ~1900 negatives for ~960 slots
    runs = 300
    hits = torch.zeros(20 + 2000, device='cuda')
    first = None
    for r in range(runs):
        res = _run(head, props, gts, labels, fused=True)[0]
        hits[res.neg_inds] += 1
        if first is None:
            first = res.neg_inds.clone()
            n_neg = res.neg_inds.numel()
        elif r == 1:
            assert not torch.equal(first, res.neg_inds)                 # a new draw every call
    cand = hits > 0
    n_cand = int(cand.sum())
    p = n_neg / n_cand
    assert n_cand > n_neg + 20                                           # a real subset is drawn
    freq = hits[cand] / runs
    sigma = (p * (1 - p) / runs) ** 0.5
    assert abs(float(freq.mean()) - p) < 1e-6                            # exactly n_neg per draw
    assert float((freq - p).abs().max()) < 5 * sigma, (float((freq - p).abs().max()), sigma)


@pytest.mark.gpu
def test_fused_sampler_empty_gt_image():
    head = _head()
    props, gts, labels, nv = _scene(7, n_img=2, K=500, G=(12, 9))
    gts[1] = gts[1][:0]
    labels[1] = labels[1][:0]
    res = _run(head, props, gts, labels, fused=True)
    assert res[1].pos_inds.numel() == 0 and res[1].pos_gt_bboxes.shape == (0, 4)
    assert res[1].neg_inds.numel() == min(int(nv[1]), head.bbox_sampler.num)
    assert res[0].pos_inds.numel() > 0
