"""Generate tests/golden/pipeline.npz by running the UNMODIFIED reference transforms
(Resize / RandomFlip / Normalize / Pad / DefaultFormatBundle image layout) over the import shim.
Run in the build container only:  python oracle/make_golden_pipeline.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402

ref_env.activate()
from mmdet.core import BitmapMasks  # noqa: E402
from mmdet.datasets.pipelines.transforms import Normalize, Pad, RandomFlip, Resize  # noqa: E402

MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]


def run(seed, H, W, flip, direction):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    G = 6
    xy = np.stack([rng.uniform(-5, W - 10, G), rng.uniform(-5, H - 10, G)], 1)
    bb = np.concatenate([xy, xy + rng.uniform(4, 30, (G, 2))], 1).astype(np.float32)
    masks = (rng.rand(G, H, W) > 0.6).astype(np.uint8)
    offs = rng.uniform(-20, 20, (G, 2)).astype(np.float32)
    res = dict(img=img.copy(), img_shape=img.shape, ori_shape=img.shape, img_fields=['img'],
               gt_bboxes=bb.copy(), bbox_fields=['gt_bboxes'],
               gt_masks=BitmapMasks(masks.copy(), H, W), mask_fields=['gt_masks'],
               gt_offsets=offs.copy(), offset_fields=['gt_offsets'])
    res = Resize(img_scale=(max(H, W), max(H, W)), keep_ratio=True)(res)
    assert float(res['scale_factor'][0]) == 1.0
    res['flip'] = flip
    res = RandomFlip(flip_ratio=0.5, direction=direction)(res)
    res = Normalize(mean=MEAN, std=STD, to_rgb=True)(res)
    res = Pad(size_divisor=32)(res)
    out_img = np.ascontiguousarray(res['img'].transpose(2, 0, 1))     # DefaultFormatBundle
    return dict(img=img, bboxes=bb, masks=masks, offsets=offs, out_img=out_img,
                out_bboxes=res['gt_bboxes'].astype(np.float32),
                out_masks=res['gt_masks'].masks.astype(np.uint8),
                out_offsets=np.asarray(res['gt_offsets'], dtype=np.float32))


def main():
    cases = [(0, 64, 96, False, 'horizontal'), (1, 64, 96, True, 'horizontal'),
             (2, 70, 50, True, 'vertical'), (3, 128, 128, True, 'vertical')]
    blob = {'mean': np.asarray(MEAN, np.float32), 'std': np.asarray(STD, np.float32),
            'n_cases': np.asarray(len(cases))}
    for i, (seed, H, W, flip, d) in enumerate(cases):
        r = run(seed, H, W, flip, d)
        blob[f'c{i}_flip'] = np.asarray(int(flip))
        blob[f'c{i}_dir'] = np.asarray(0 if d == 'horizontal' else 1)
        for k, v in r.items():
            blob[f'c{i}_{k}'] = v
    path = os.path.join(ROOT, 'tests', 'golden', 'pipeline.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
