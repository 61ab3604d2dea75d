"""Generate tests/golden/bonai_ann.json: a synthetic BONAI-format COCO json (the extension fields
of mmdet/datasets/bonai.py) and what the UNMODIFIED reference `BONAI._parse_ann_info` returns for
each tile under several (bbox_type, mask_type, offset_coordinate) settings.  The method is called
unbound on a stand-in `self` carrying exactly the attributes it reads, so no pycocotools is needed.
Run in the build container only:  python oracle/make_golden_bonai_ann.py"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402

ref_env.activate()
from mmdet.datasets.bonai import BONAI  # noqa: E402


def make_json(seed=0, n_img=3, W=256, H=192):
    rng = np.random.RandomState(seed)
    images, anns = [], []
    aid = 1
    for i in range(n_img):
        images.append(dict(id=10 + i, file_name=f'tile_{i}.png', width=W, height=H))
        for j in range(0 if i == 2 else 7):
            x, y = float(rng.uniform(-10, W - 20)), float(rng.uniform(-10, H - 20))
            w, h = float(rng.uniform(0.5, 60)), float(rng.uniform(0.5, 60))
            ox, oy = float(rng.uniform(-15, 15)), float(rng.uniform(-15, 15))
            roof = [x, y, x + w, y, x + w, y + h, x, y + h]
            foot = [v + (ox if k % 2 == 0 else oy) for k, v in enumerate(roof)]
            a = dict(id=aid, image_id=10 + i, category_id=1 if j != 5 else 2,
                     iscrowd=1 if j == 3 else 0, area=w * h if j != 4 else 0.0,
                     bbox=[x, y, w, h], roof_bbox=[x, y, w, h],
                     footprint_bbox=[x + ox, y + oy, w, h],
                     building_bbox=[min(x, x + ox), min(y, y + oy), w + abs(ox), h + abs(oy)],
                     segmentation=[roof], footprint_mask=foot, offset=[ox, oy],
                     building_height=float(rng.uniform(3, 90)))
            if j == 1:
                a['only_footprint'] = 1
            if j == 2:
                a['only_footprint'] = 0
                del a['building_height']
            if j == 6:
                a['ignore'] = True
            anns.append(a)
            aid += 1
    return dict(images=images, annotations=anns,
                categories=[dict(id=1, name='building'), dict(id=2, name='other')])


def tolist(v):
    return v.tolist() if isinstance(v, np.ndarray) else v


def main():
    coco = make_json()
    settings = [dict(bbox_type='building', mask_type='roof', offset_coordinate='rectangle'),
                dict(bbox_type='roof', mask_type='footprint', offset_coordinate='polar'),
                dict(bbox_type='footprint', mask_type='roof', offset_coordinate='rectangle')]
    out = dict(coco=coco, settings=settings, parsed=[])
    for st in settings:
        fake = SimpleNamespace(cat_ids=[1], cat2label={1: 0}, resolution=0.6,
                               ignore_buildings=True, **st)
        per_img = []
        for im in coco['images']:
            info = dict(im, filename=im['file_name'])
            ann_info = [a for a in coco['annotations'] if a['image_id'] == im['id']]
            if not ann_info:
                per_img.append(None)
                continue
            r = BONAI._parse_ann_info(fake, info, ann_info)
            per_img.append({k: tolist(v) for k, v in r.items()})
        out['parsed'].append(per_img)
    path = os.path.join(ROOT, 'tests', 'golden', 'bonai_ann.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
