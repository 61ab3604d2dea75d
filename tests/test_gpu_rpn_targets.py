"""SURVEY 8(a) rows a7 / a9 on the GPU: the no-read-back RPN sampling + target construction
(`loft_rpn_targets`: AnchorHead._get_targets_single after the assigner + RandomSampler +
images_to_levels, mmdet/models/dense_heads/anchor_head.py:206-278,363-380,
core/bbox/samplers/random_sampler.py:31-75) against the op-by-op path of the same head: identical
targets when nothing has to be drawn, the sampler's counts and candidate sets when something has,
and a uniform draw."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py')


@pytest.fixture(scope='module')
def head():
    from bonai_b200 import Config
    from bonai_b200.engine import get_store
    from bonai_b200.models import build_detector
    cfg = Config.fromfile(CFG)
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    get_store(model, torch.device('cuda:0'))
    return model.rpn_head


def _gts(n_img, size, counts, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for c in counts[:n_img]:
        xy = torch.rand(c, 2, generator=g) * (size - 80)
        wh = 12 + torch.rand(c, 2, generator=g) * 60
        out.append(torch.cat([xy, xy + wh], 1).cuda())
    return out


def _unfused(head, sizes, gts, metas, dev):
    os.environ['LOFT_FUSED_RPN_TARGETS'] = '0'
    try:
        return head._build_targets(sizes, gts, metas, dev)
    finally:
        del os.environ['LOFT_FUSED_RPN_TARGETS']


def _assigned(head, sizes, gts, dev):
    flat = head._flat_anchors(sizes, dev)
    return [head.assigner.assign(flat, g, None, None).gt_inds for g in gts]


def _per_image(per_level, n_img, which, width=1):
    """[n_img, A(*width)] view of one target quantity, levels concatenated like the anchors."""
    return torch.cat([lv[which].view(n_img, -1) for lv in per_level], 1)


@pytest.mark.parametrize('n_img,size,counts', [(2, 256, (10, 7)), (1, 320, (3,)), (2, 192, (0, 5))])
def test_targets_equal_the_op_by_op_path_when_nothing_is_drawn(head, n_img, size, counts):
    dev = torch.device('cuda:0')
    sizes = head.featmap_sizes_for((size, size))
    metas = [dict(img_shape=(size, size, 3), pad_shape=(size, size, 3))] * n_img
    gts = _gts(n_img, size, counts, seed=1)
    smp = head.sampler
    num0 = smp.num
    smp.num = 1 << 28                  # quotas above every candidate count: the draw is the identity
    try:
        want, want_total = _unfused(head, sizes, gts, metas, dev)
        got, got_total = head._build_targets(sizes, gts, metas, dev)
    finally:
        smp.num = num0
    assert isinstance(got_total, torch.Tensor) and not isinstance(want_total, torch.Tensor)
    assert float(got_total) == float(want_total)
    for l, (g, w) in enumerate(zip(got, want)):
        for q, name in enumerate(('labels', 'label_weights', 'bbox_targets', 'bbox_weights')):
            assert torch.equal(g[q], w[q].float()), (l, name)
    from bonai_b200 import _lib as L
    assert L.lib().loft_rpn_targets_overflowed(L.ptr(head._last_target_ws), L.stream()) == 0


def test_sampled_counts_candidates_and_encoding(head):
    dev = torch.device('cuda:0')
    n_img, size = 2, 512
    sizes = head.featmap_sizes_for((size, size))
    metas = [dict(img_shape=(size, size, 3), pad_shape=(size, size, 3))] * n_img
    gts = _gts(n_img, size, (90, 40), seed=2)
    smp = head.sampler
    num, num_pos_max = int(smp.num), int(smp.num * smp.pos_fraction)
    gi = _assigned(head, sizes, gts, dev)
    smp.num = 1 << 28                  # every candidate's targets, from the op-by-op path
    try:
        full, _ = _unfused(head, sizes, gts, metas, dev)
    finally:
        smp.num = num
    got, total = head._build_targets(sizes, gts, metas, dev)
    lab, lw = _per_image(got, n_img, 0), _per_image(got, n_img, 1)
    bt = torch.cat([lv[2].view(n_img, -1, 4) for lv in got], 1)
    bw = torch.cat([lv[3].view(n_img, -1, 4) for lv in got], 1)
    bt_full = torch.cat([lv[2].view(n_img, -1, 4) for lv in full], 1)
    tot = 0
    for i in range(n_img):
        pos_c, neg_c = gi[i] > 0, gi[i] == 0
        take_pos = min(int(pos_c.sum()), num_pos_max)
        take_neg = min(int(neg_c.sum()), num - take_pos)
        sel_pos, sel_neg = lab[i] > 0, (lw[i] > 0) & (lab[i] == 0)
        assert int(sel_pos.sum()) == take_pos and int(sel_neg.sum()) == take_neg
        assert bool((sel_pos & ~pos_c).sum() == 0) and bool((sel_neg & ~neg_c).sum() == 0)
        assert torch.equal(lw[i] > 0, sel_pos | sel_neg) and bool((lw[i][lw[i] > 0] == 1).all())
        assert torch.equal(bw[i], sel_pos[:, None].expand(-1, 4).float())
        assert torch.equal(bt[i][sel_pos], bt_full[i][sel_pos].float())     # same encoder
        assert bool((bt[i][~sel_pos] == 0).all())
        tot += max(take_pos, 1) + max(take_neg, 1)
    assert int(pos_c.sum()) > 0
    assert float(total) == float(tot)


def test_draw_is_uniform_and_differs_between_calls(head):
    """Each negative anchor is drawn with the same probability: over many calls the hit counts of
    16 index buckets follow the buckets' sizes (chi-square far below a fixed-subset value)."""
    dev = torch.device('cuda:0')
    size = 256
    sizes = head.featmap_sizes_for((size, size))
    metas = [dict(img_shape=(size, size, 3), pad_shape=(size, size, 3))]
    gts = _gts(1, size, (6,), seed=3)
    neg_c = _assigned(head, sizes, gts, dev)[0] == 0
    hits = torch.zeros(neg_c.numel(), device=dev)
    prev, changed, calls = None, 0, 200
    for _ in range(calls):
        got, _ = head._build_targets(sizes, gts, metas, dev)
        sel = (_per_image(got, 1, 1)[0] > 0) & (_per_image(got, 1, 0)[0] == 0)
        hits += sel.float()
        if prev is not None and not torch.equal(prev, sel):
            changed += 1
        prev = sel
    assert changed == calls - 1
    idx = torch.nonzero(neg_c)[:, 0]
    buckets = torch.chunk(idx, 16)
    drawn = float(hits.sum())
    chi2 = 0.0
    for b in buckets:
        expect = drawn * b.numel() / idx.numel()
        chi2 += (float(hits[b].sum()) - expect) ** 2 / expect
    assert chi2 < 50.0, chi2            # 15 degrees of freedom: P(chi2 > 50) ~ 1e-5
    assert float(hits.max()) <= calls * 0.25      # nobody is drawn (almost) every time
