"""Post-load training pipeline (first slice): the CPU restatement against golden
vectors produced by the reference's own transform classes, and the device kernels against both."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'pipeline.npz'))


def _cases(g):
    for i in range(int(g['n_cases'])):
        d = 'horizontal' if int(g[f'c{i}_dir']) == 0 else 'vertical'
        yield i, bool(g[f'c{i}_flip']), d


def test_pipeline_oracle_matches_reference_golden(golden):
    """Bit-exact: image (float32), boxes, bitmaps, offsets -- reference transforms.py:222-229,
    378-404, 458-466, 484-488, 571-600, 655-676 via oracle/make_golden_pipeline.py."""
    from oracle import pipeline_cpu as P
    g = golden
    for i, flip, d in _cases(g):
        x, b, m, o, meta = P.train_pipeline(g[f'c{i}_img'], g[f'c{i}_bboxes'], g[f'c{i}_masks'],
                                            g[f'c{i}_offsets'], flip, d, g['mean'], g['std'])
        assert np.array_equal(x, g[f'c{i}_out_img'])
        assert np.array_equal(b, g[f'c{i}_out_bboxes'])
        assert np.array_equal(m, g[f'c{i}_out_masks'])
        assert np.array_equal(o, g[f'c{i}_out_offsets'])
        assert meta['pad_shape'][0] % 32 == 0 and meta['pad_shape'][1] % 32 == 0


def test_pipeline_builds_from_reference_config():
    from bonai_b200 import Config
    from bonai_b200.datasets import GpuTrainPipeline
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    p = GpuTrainPipeline.from_cfg(cfg.data.train.pipeline, device='cpu',
                                  rng=np.random.RandomState(0))
    assert p.direction in ('horizontal', 'vertical') and p.flip_ratio == 0.5
    assert p.size_divisor == 32 and p.to_rgb and tuple(p.img_scale) == (1024, 1024)
    assert np.allclose(p.mean, [123.675, 116.28, 103.53])
    ref_cfg = '/root/reference/configs/loft_foa/loft_foa_r50_fpn_2x_bonai.py'
    if os.path.exists(ref_cfg):                         # the reference's own config, unchanged
        q = GpuTrainPipeline.from_cfg(Config.fromfile(ref_cfg).data.train.pipeline, device='cpu')
        assert (q.flip_ratio, q.size_divisor, q.to_rgb) == (p.flip_ratio, p.size_divisor, p.to_rgb)
        assert np.array_equal(q.mean, p.mean) and np.array_equal(q.std, p.std)
    with pytest.raises(NotImplementedError):
        GpuTrainPipeline.from_cfg([dict(type='RandomCrop', crop_size=(512, 512))], device='cpu')
    with pytest.raises(NotImplementedError):          # a tile that would need resampling
        p(np.zeros((2048, 2048, 3), np.uint8), np.zeros((0, 4)), np.zeros((0,)),
          np.zeros((0, 2048, 2048), np.uint8), np.zeros((0, 2)))


@pytest.mark.gpu
def test_pipeline_kernels_bit_exact_vs_reference_golden(golden):
    from bonai_b200.datasets import GpuTrainPipeline
    g = golden
    for i, flip, d in _cases(g):
        H, W = g[f'c{i}_img'].shape[:2]
        p = GpuTrainPipeline(img_scale=(max(H, W), max(H, W)), direction=d, mean=g['mean'],
                             std=g['std'])
        out = p(g[f'c{i}_img'], g[f'c{i}_bboxes'], np.zeros(len(g[f'c{i}_bboxes']), np.int64),
                g[f'c{i}_masks'], g[f'c{i}_offsets'], flip=flip)
        assert torch.equal(out['img'].cpu(), torch.from_numpy(g[f'c{i}_out_img']))
        assert torch.equal(out['gt_bboxes'].cpu(), torch.from_numpy(g[f'c{i}_out_bboxes']))
        assert torch.equal(out['gt_masks'].to_tensor(device='cuda').cpu(),
                           torch.from_numpy(g[f'c{i}_out_masks']))
        assert torch.equal(out['gt_offsets'].cpu(), torch.from_numpy(g[f'c{i}_out_offsets']))
        assert out['img_metas']['flip'] == flip and out['img_metas']['pad_shape'][0] % 32 == 0


@pytest.mark.gpu
def test_pipeline_full_size_vs_oracle_and_feeds_the_model():
    """1024^2 tile, 80 buildings: kernels == oracle bit for bit for both flip directions; flipping
    twice is the identity; the collated batch is accepted by LOFT.forward_train."""
    from bonai_b200.datasets import GpuTrainPipeline, image_prep, mask_flip_pad
    from oracle import pipeline_cpu as P
    rng = np.random.RandomState(7)
    H = W = 1024
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    G = 80
    xy = rng.uniform(0, 900, (G, 2))
    bb = np.concatenate([xy, xy + rng.uniform(16, 120, (G, 2))], 1).astype(np.float32)
    masks = np.zeros((G, H, W), np.uint8)
    for k in range(G):
        x1, y1, x2, y2 = bb[k].astype(int)
        masks[k, y1:min(y2, H), x1:min(x2, W)] = 1
    offs = rng.uniform(-40, 40, (G, 2)).astype(np.float32)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    samples = []
    for d in ('horizontal', 'vertical'):
        p = GpuTrainPipeline(direction=d)
        out = p(img, bb, np.zeros(G, np.int64), masks, offs, flip=True)
        x, b, m, o, _ = P.train_pipeline(img, bb, masks, offs, True, d, mean, std)
        assert torch.equal(out['img'].cpu(), torch.from_numpy(x))
        assert torch.equal(out['gt_masks'].to_tensor(device='cuda').cpu(), torch.from_numpy(m))
        assert torch.equal(out['gt_bboxes'].cpu(), torch.from_numpy(b))
        assert torch.equal(out['gt_offsets'].cpu(), torch.from_numpy(o))
        md = torch.from_numpy(masks).cuda()
        assert torch.equal(mask_flip_pad(mask_flip_pad(md, d), d), md)          # involution
        samples.append(out)
    batch = GpuTrainPipeline.collate(samples)
    assert batch['img'].shape == (2, 3, 1024, 1024)
    from bonai_b200 import Config
    from bonai_b200.models import build_detector
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.train()
    losses = model.forward_train(**batch)
    loss, logs = model._parse_losses(losses)
    assert torch.isfinite(loss)
