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
    assert p._scale_factor(2048, 2048) == 0.5 and p._scale_factor(1024, 1024) == 1.0
    assert p._scale_factor(512, 1024) == 1.0          # keep_ratio: the long edge decides


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


# ------------------------------------------------------------------ f3: BONAI annotations, polygons
def _ann_golden():
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'bonai_ann.json')) as f:
        return json.load(f)


def _write_bonai_tiles(tmp_path, coco, seed=0):
    """The golden json + one PNG per tile under tmp_path; returns (ann_file, img_prefix)."""
    import json
    from PIL import Image
    rng = np.random.RandomState(seed)
    for im in coco['images']:
        a = rng.randint(0, 256, (im['height'], im['width'], 3)).astype(np.uint8)
        Image.fromarray(a).save(os.path.join(str(tmp_path), im['file_name']))
    ann = os.path.join(str(tmp_path), 'bonai_test.json')
    with open(ann, 'w') as f:
        json.dump(coco, f)
    return ann, str(tmp_path)


def test_bonai_parse_ann_info_matches_reference(tmp_path):
    """`BONAI._parse_ann_info` (bonai.py:106-256) for three (bbox_type, mask_type,
    offset_coordinate) settings == the unmodified reference method
    (oracle/make_golden_bonai_ann.py): ignore / crowd / zero-area / foreign-category filtering,
    the sticky only_footprint flag, polar offsets, mean angle, empty tiles filtered out."""
    from bonai_b200.datasets import BONAI
    from bonai_b200 import Config
    g = _ann_golden()
    ann, prefix = _write_bonai_tiles(tmp_path, g['coco'])
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    for st, parsed in zip(g['settings'], g['parsed']):
        ds = BONAI(ann_file=ann, pipeline=cfg.data.train.pipeline, img_prefix=prefix, **st)
        # tile 2 has no annotations: filtered (filter_empty_gt), as bonai.py:88-104
        assert [d['id'] for d in ds.data_infos] == [10, 11]
        assert ds.flag.tolist() == [1, 1]
        for idx in range(len(ds)):
            got, ref = ds.get_ann_info(idx), parsed[idx]
            assert set(got) == set(ref)
            for k, v in ref.items():
                if isinstance(got[k], np.ndarray):
                    assert got[k].dtype == (np.int64 if k == 'labels' else np.float32), k
                    assert got[k].tolist() == v, k
                else:
                    assert got[k] == v, k


def test_bonai_build_dataset_from_reference_config_layout(tmp_path):
    """A list of annotation files (one per city, bonai_instance.py:33-48) gives one dataset each."""
    from bonai_b200.datasets import build_dataset
    from bonai_b200 import Config
    g = _ann_golden()
    ann, prefix = _write_bonai_tiles(tmp_path, g['coco'])
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    c = dict(cfg.data.train)
    assert c['type'] == 'BONAI' and c['bbox_type'] == 'building' and c['mask_type'] == 'roof'
    c.update(ann_file=[ann, ann], img_prefix=[prefix, prefix])
    dss = build_dataset(c)
    assert len(dss) == 2 and all(len(d) == 2 for d in dss)


def test_polygon_oracle_known_answers():
    """pycocotools rleFrPoly semantics (PARITY UNPINNED: the library is not in this image; these
    are the hand-checkable cases of the published algorithm, oracle/polygon_cpu.py)."""
    from oracle.polygon_cpu import poly2mask
    m = poly2mask([[10, 10, 20, 10, 20, 20, 10, 20]], 32, 32)      # integer square: [10, 20)^2
    ref = np.zeros((32, 32), np.uint8)
    ref[10:20, 10:20] = 1
    assert np.array_equal(m, ref)
    m = poly2mask([[10.5, 10.5, 20.5, 10.5, 20.5, 20.5, 10.5, 20.5]], 32, 32)   # half-pixel shift
    ref = np.zeros((32, 32), np.uint8)
    ref[11:21, 11:21] = 1
    assert np.array_equal(m, ref)
    m = poly2mask([[2, 2, 12, 2, 2, 12]], 16, 16)                   # right triangle, legs of 10
    assert int(m.sum()) == 45 and m[2, 2:11].all() and m[10, 2] == 1 and m[11, 2] == 0
    m = poly2mask([[-5, -5, 8, -5, 8, 8, -5, 8]], 16, 16)           # clipped at the border
    ref = np.zeros((16, 16), np.uint8)
    ref[:8, :8] = 1
    assert np.array_equal(m, ref)
    m = poly2mask([[10, 10, 5, 5]], 32, 32)                         # 4 numbers = [x, y, w, h] box
    ref = np.zeros((32, 32), np.uint8)
    ref[10:15, 10:15] = 1
    assert np.array_equal(m, ref)
    two = poly2mask([[1, 1, 4, 1, 4, 4, 1, 4], [6, 6, 9, 6, 9, 9, 6, 9]], 12, 12)   # union of parts
    assert int(two.sum()) == 18 and two[2, 2] == 1 and two[7, 7] == 1 and two[5, 5] == 0
    assert poly2mask([], 8, 8).sum() == 0


def _random_polygons(rng, n, H, W):
    out = []
    for i in range(n):
        cx, cy = rng.uniform(-10, W + 10), rng.uniform(-10, H + 10)
        k = int(rng.randint(3, 24))
        if i % 3 == 0:        # star-shaped (concave), fractional vertices
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            r = rng.uniform(3, 0.3 * min(H, W), k)
            pts = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1)
        elif i % 3 == 1:      # rotated rectangle, integer-ish
            w, h, t = rng.uniform(2, 80), rng.uniform(2, 80), rng.uniform(0, np.pi)
            c = np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2
            R = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
            pts = np.round(c @ R.T + [cx, cy], 1)
        else:                 # arbitrary (self-intersecting) vertex order
            pts = rng.uniform(0, 1, (k, 2)) * [0.25 * W, 0.25 * H] + [cx, cy]
        parts = [pts.reshape(-1).tolist()]
        if i % 5 == 4:        # a second part
            parts.append((pts + rng.uniform(5, 30)).reshape(-1).tolist())
        out.append(parts)
    out.append([[3, 3, 9, 3, 9, 9, 3, 9]])
    out.append([[5.0, 6.0, 7.0, 4.0]])            # box form
    out.append([[W - 2.0, H - 2.0, W + 30.0, H - 2.0, W + 30.0, H + 30.0]])
    return out


@pytest.mark.gpu
def test_polygon_rasterize_kernel_bit_exact_vs_oracle():
    """`loft_poly_rasterize` == the CPU restatement of pycocotools frPyObjects/merge/decode, bit for
    bit, on random concave / rotated / self-intersecting / multi-part / partly outside polygons."""
    from oracle.polygon_cpu import poly2mask
    from bonai_b200.datasets import polygons_to_bitmaps
    for seed, (H, W, n) in enumerate([(96, 160, 40), (1024, 1024, 24), (33, 47, 12)]):
        rng = np.random.RandomState(seed)
        polys = _random_polygons(rng, n, H, W)
        got = polygons_to_bitmaps(polys, H, W, 'cuda').cpu().numpy()
        assert got.shape == (len(polys), H, W) and got.dtype == np.uint8
        for i, parts in enumerate(polys):
            ref = poly2mask(parts, H, W)
            assert np.array_equal(got[i], ref), (seed, i, int(got[i].sum()), int(ref.sum()))
    assert polygons_to_bitmaps([], 64, 64, 'cuda').shape == (0, 64, 64)
    empty_inst = polygons_to_bitmaps([[], [[1, 1, 5, 1, 5, 5]]], 16, 16, 'cuda').cpu().numpy()
    assert empty_inst[0].sum() == 0 and np.array_equal(empty_inst[1], poly2mask([[1, 1, 5, 1, 5, 5]], 16, 16))


@pytest.mark.gpu
def test_bonai_dataset_tile_equals_cpu_pipeline(tmp_path):
    """json + PNG -> `BONAI.prepare_train_img` (host decode, device rasterisation + pipeline) ==
    the CPU restatements (oracle/polygon_cpu.py, oracle/pipeline_cpu.py) on the same tile."""
    from PIL import Image
    from oracle import pipeline_cpu as P
    from oracle.polygon_cpu import poly2mask
    from bonai_b200.datasets import BONAI
    from bonai_b200 import Config
    g = _ann_golden()
    ann, prefix = _write_bonai_tiles(tmp_path, g['coco'])
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    # the golden tiles are 192 x 256: same pipeline with the Resize target at their size (identity)
    pipe = [dict(t, img_scale=(256, 256)) if t['type'] == 'Resize' else dict(t)
            for t in cfg.data.train.pipeline]
    ds = BONAI(ann_file=ann, pipeline=pipe, img_prefix=prefix,
               bbox_type='building', mask_type='roof', rng=np.random.RandomState(0))
    for idx, flip in [(0, False), (1, True)]:
        out = ds.prepare_train_img(idx, flip=flip)
        info, a = ds.data_infos[idx], ds.get_ann_info(idx)
        bgr = np.asarray(Image.open(os.path.join(prefix, info['filename'])).convert('RGB'))[:, :, ::-1]
        masks = np.stack([poly2mask(m, info['height'], info['width']) for m in a['masks']])
        x, b, m, o, meta = P.train_pipeline(np.ascontiguousarray(bgr), a['bboxes'], masks,
                                            a['offsets'], flip, ds.pipeline.direction,
                                            ds.pipeline.mean, ds.pipeline.std)
        assert np.array_equal(out['img'].cpu().numpy(), x)
        assert np.array_equal(out['gt_bboxes'].cpu().numpy(), b)
        assert np.array_equal(out['gt_masks'].to_tensor(device='cuda').cpu().numpy(), m)
        assert np.array_equal(out['gt_offsets'].cpu().numpy(), o)
        assert out['gt_labels'].tolist() == a['labels'].tolist()
        assert out['img_metas']['pad_shape'] == tuple(meta['pad_shape'])


def test_resize_oracle_known_answers():
    """cv2.resize restatement (PARITY UNPINNED, cv2 is not in this image): properties any correct
    INTER_LINEAR / INTER_NEAREST has -- identity size, constants stay constant, exact 2x nearest
    replication, exact 2x down-sampling = mean of the two centre taps, monotone ramps stay monotone."""
    from oracle import pipeline_cpu as P
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
    assert np.array_equal(P.resize_bilinear_u8(img, 53, 37), img)
    const = np.full((20, 30, 3), 77, np.uint8)
    assert (P.resize_bilinear_u8(const, 71, 45) == 77).all()
    m = rng.randint(0, 2, (3, 16, 24)).astype(np.uint8)
    up = P.resize_nearest_u8(m, 48, 32)
    assert np.array_equal(up, m.repeat(2, axis=1).repeat(2, axis=2))
    assert np.array_equal(P.resize_nearest_u8(m, 24, 16), m)
    even = rng.randint(0, 256, (16, 32, 3)).astype(np.uint8)
    dn = P.resize_bilinear_u8(even, 16, 8).astype(np.int64)
    e = even.astype(np.int64)
    ref = (e[0::2, 0::2] + e[0::2, 1::2] + e[1::2, 0::2] + e[1::2, 1::2] + 2) >> 2
    assert np.abs(dn - ref).max() <= 1                 # 2x: both taps weigh 1024/2048
    ramp = np.tile(np.arange(64, dtype=np.uint8)[None, :, None] * 4, (8, 1, 3))
    r = P.resize_bilinear_u8(ramp, 150, 8).astype(np.int64)
    assert (np.diff(r[0, :, 0]) >= 0).all() and r[0, 0, 0] == 0 and r[0, -1, 0] == 252
    assert P.rescale_size(2048, 1536, (1024, 1024)) == (1024, 768)
    assert P.rescale_size(1024, 1024, (1024, 1024)) == (1024, 1024)


@pytest.mark.gpu
def test_resize_kernels_bit_exact_vs_oracle_and_pipeline():
    """`loft_resize_bilinear_u8` / `loft_resize_nearest_u8` == the CPU restatement, bit for bit, up
    and down, odd sizes; and the whole pipeline on a tile that needs resampling == the oracle."""
    from oracle import pipeline_cpu as P
    from bonai_b200.datasets import GpuTrainPipeline, resize_bilinear_u8, resize_nearest_u8
    rng = np.random.RandomState(1)
    for (h, w), (nh, nw) in [((37, 53), (74, 106)), ((128, 96), (64, 48)), ((100, 75), (333, 250)),
                             ((333, 250), (100, 75)), ((1536, 2048), (768, 1024)), ((5, 7), (5, 7))]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        got = resize_bilinear_u8(torch.from_numpy(img).cuda(), nh, nw).cpu().numpy()
        assert np.array_equal(got, P.resize_bilinear_u8(img, nw, nh)), (h, w, nh, nw)
        m = rng.randint(0, 2, (4, h, w)).astype(np.uint8)
        gm = resize_nearest_u8(torch.from_numpy(m).cuda(), nh, nw).cpu().numpy()
        assert np.array_equal(gm, P.resize_nearest_u8(m, nw, nh)), (h, w, nh, nw)
    # whole pipeline, 96 x 128 tile onto img_scale (64, 64): keep_ratio -> 48 x 64
    H, W, G = 96, 128, 5
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    xy = np.stack([rng.uniform(0, W - 20, G), rng.uniform(0, H - 20, G)], 1)
    bb = np.concatenate([xy, xy + rng.uniform(4, 30, (G, 2))], 1).astype(np.float32)
    masks = (rng.rand(G, H, W) > 0.6).astype(np.uint8)
    offs = rng.uniform(-20, 20, (G, 2)).astype(np.float32)
    for flip in (False, True):
        p = GpuTrainPipeline(img_scale=(64, 64), direction='horizontal', device='cuda')
        out = p(img, bb, np.zeros(G, np.int64), torch.from_numpy(masks), offs, flip=flip)
        x, b, m, o, meta = P.train_pipeline(img, bb, masks, offs, flip, 'horizontal', p.mean, p.std,
                                            img_scale=(64, 64))
        assert out['img'].shape == (3, 64, 64) and meta['img_shape'] == (48, 64, 3)
        assert np.array_equal(out['img'].cpu().numpy(), x)
        assert np.array_equal(out['gt_bboxes'].cpu().numpy(), b)
        assert np.array_equal(out['gt_masks'].to_tensor(device='cuda').cpu().numpy(), m)
        assert np.array_equal(out['gt_offsets'].cpu().numpy(), o)
        assert np.array_equal(out['img_metas']['scale_factor'], meta['scale_factor'])
        assert out['img_metas']['img_shape'] == (48, 64, 3)
