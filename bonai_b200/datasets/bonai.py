"""BONAI dataset front end of the device-side input pipeline.

Mirrors `mmdet/datasets/bonai.py` (class BONAI, :14-257) + the parts of `CocoDataset` /
`CustomDataset` it relies on (coco.py:36-76, custom.py:95-190) for the TRAINING path of
`configs/_base_/datasets/bonai_instance.py`: COCO-style json with the BONAI extension fields
(`building_bbox`, `roof_bbox`, `footprint_bbox`, `segmentation` = roof polygon, `footprint_mask`,
`offset`, `building_height`, `only_footprint`), parsed with the reference's rules, then

    LoadImageFromFile   host: file bytes -> uint8 HWC BGR (PIL / libpng / libjpeg; the reference
                        decodes on the host too, with cv2 -- entropy decoding is serial bit-stream
                        work, neither HBM- nor tensor-bound, and stays there)
    LoadAnnotations     device: polygons -> uint8 bitmaps (`loft_poly_rasterize`; in the reference
                        pycocotools on the host, 1 MB per building then shipped over PCIe)
    Resize .. Collect   device: `GpuTrainPipeline` (gpu_pipeline.py)

so a tile crosses PCIe as 3 MB of pixels plus a few KB of vertices instead of 12 MB + 1 MB per
building.  Evaluation / result dumping (`bonai.py:258-500`) is out of scope.
"""
import json
import math
import os.path as osp
from collections import defaultdict

import numpy as np
import torch

from ..registry import Registry, build_from_cfg
from .gpu_pipeline import GpuTrainPipeline, polygons_to_bitmaps

DATASETS = Registry('dataset')


class _CocoIndex:
    """The five pycocotools.coco.COCO calls `CocoDataset.load_annotations` / `get_ann_info` make
    (coco.py:46-76), over the parsed json."""

    def __init__(self, ann_file):
        with open(ann_file) as f:
            d = json.load(f)
        self.imgs = {im['id']: im for im in d.get('images', [])}
        self.anns = {a['id']: a for a in d.get('annotations', [])}
        self.cats = {c['id']: c for c in d.get('categories', [])}
        self.img_to_anns = defaultdict(list)
        for a in d.get('annotations', []):
            self.img_to_anns[a['image_id']].append(a)

    def get_cat_ids(self, cat_names):
        names = [cat_names] if isinstance(cat_names, str) else list(cat_names)
        return [c['id'] for c in self.cats.values() if c['name'] in names]

    def get_img_ids(self):
        return list(self.imgs.keys())

    def load_imgs(self, ids):
        return [self.imgs[i] for i in ids]

    def get_ann_ids(self, img_ids):
        return [a['id'] for i in img_ids for a in self.img_to_anns.get(i, [])]

    def load_anns(self, ids):
        return [self.anns[i] for i in ids]


@DATASETS.register_module()
class BONAI:
    """Same constructor arguments as the reference class (bonai.py:17-35); `pipeline` is the
    reference's list of transform dicts (consumed by `GpuTrainPipeline.from_cfg`)."""
    CLASSES = ('building')          # sic: a str in the reference (bonai.py:16)

    def __init__(self, ann_file, pipeline, classes=None, data_root=None, img_prefix='',
                 seg_prefix=None, edge_prefix=None, side_face_prefix=None,
                 offset_field_prefix=None, proposal_file=None, test_mode=False,
                 filter_empty_gt=True, gt_footprint_csv_file=None, bbox_type='roof',
                 mask_type='roof', offset_coordinate='rectangle', resolution=0.6,
                 ignore_buildings=True, device='cuda', rng=None):
        if test_mode:
            raise NotImplementedError('BONAI: only the training path is on the device pipeline')
        self.ann_file, self.data_root, self.img_prefix = ann_file, data_root, img_prefix
        if data_root is not None:
            if not osp.isabs(self.ann_file):
                self.ann_file = osp.join(data_root, self.ann_file)
            if not (self.img_prefix is None or osp.isabs(self.img_prefix)):
                self.img_prefix = osp.join(data_root, self.img_prefix)
        if classes is not None:
            self.CLASSES = tuple(classes) if not isinstance(classes, str) else classes
        self.bbox_type, self.mask_type = bbox_type, mask_type
        self.offset_coordinate, self.resolution = offset_coordinate, resolution
        self.ignore_buildings, self.filter_empty_gt = ignore_buildings, filter_empty_gt
        self.test_mode = test_mode
        self.device = torch.device(device)
        self.data_infos = self.load_annotations(self.ann_file)
        valid = self._filter_imgs()
        self.data_infos = [self.data_infos[i] for i in valid]
        self.img_ids = [self.img_ids[i] for i in valid]
        self.load_cfg = next((dict(t) for t in pipeline if t.get('type') == 'LoadAnnotations'), {})
        self.pipeline = GpuTrainPipeline.from_cfg(pipeline, device=device, rng=rng)
        self._set_group_flag()

    def __len__(self):
        return len(self.data_infos)

    # ---------------------------------------------------------------- coco.py:36-76
    def load_annotations(self, ann_file):
        self.coco = _CocoIndex(ann_file)
        self.cat_ids = self.coco.get_cat_ids(cat_names=self.CLASSES)
        self.cat2label = {cat_id: i for i, cat_id in enumerate(self.cat_ids)}
        self.img_ids = self.coco.get_img_ids()
        data_infos = []
        for i in self.img_ids:
            info = self.coco.load_imgs([i])[0]
            info['filename'] = info['file_name']
            data_infos.append(info)
        return data_infos

    def get_ann_info(self, idx):
        img_id = self.data_infos[idx]['id']
        ann_info = self.coco.load_anns(self.coco.get_ann_ids(img_ids=[img_id]))
        return self._parse_ann_info(self.data_infos[idx], ann_info)

    def _filter_imgs(self, min_size=32):
        """bonai.py:88-104: drop tiles that are too small, have no annotation or only crowds."""
        valid_inds = []
        ids_with_ann = set(a['image_id'] for a in self.coco.anns.values())
        for i, img_info in enumerate(self.data_infos):
            ann_info = self.coco.load_anns(self.coco.get_ann_ids(img_ids=[img_info['id']]))
            all_iscrowd = all([a['iscrowd'] for a in ann_info])
            if self.filter_empty_gt and (self.img_ids[i] not in ids_with_ann or all_iscrowd):
                continue
            if min(img_info['width'], img_info['height']) >= min_size:
                valid_inds.append(i)
        return valid_inds

    def _set_group_flag(self):
        """custom.py:176-186: aspect-ratio groups for GroupSampler."""
        self.flag = np.zeros(len(self), dtype=np.uint8)
        for i, info in enumerate(self.data_infos):
            if info['width'] / info['height'] > 1:
                self.flag[i] = 1

    # ---------------------------------------------------------------- bonai.py:106-256
    def _parse_ann_info(self, img_info, ann_info):
        gt_bboxes, gt_labels, gt_bboxes_ignore = [], [], []
        gt_masks_ann, gt_roof_masks_ann, gt_footprint_masks_ann = [], [], []
        gt_offsets, gt_building_heights, gt_angles = [], [], []
        gt_roof_bboxes, gt_footprint_bboxes = [], []
        only_fp = 0
        key = {'roof': 'bbox', 'building': 'building_bbox', 'footprint': 'footprint_bbox'}
        if self.bbox_type not in key:
            raise TypeError(f"don't support bbox_type={self.bbox_type}")
        for ann in ann_info:
            if ann.get('ignore', False):
                continue
            x1, y1, w, h = ann[key[self.bbox_type]]
            inter_w = max(0, min(x1 + w, img_info['width']) - max(x1, 0))
            inter_h = max(0, min(y1 + h, img_info['height']) - max(y1, 0))
            if inter_w * inter_h == 0:
                continue
            if ann['area'] <= 0 or w < 1 or h < 1:
                continue
            if ann['category_id'] not in self.cat_ids:
                continue
            bbox = [x1, y1, x1 + w, y1 + h]
            if ann.get('iscrowd', False) and self.ignore_buildings:
                gt_bboxes_ignore.append(bbox)
                continue
            if 'roof_bbox' in ann:
                x, y, w_, h_ = ann['roof_bbox']
                gt_roof_bboxes.append([x, y, x + w_, y + h_])
            if 'footprint_bbox' in ann:
                x, y, w_, h_ = ann['footprint_bbox']
                gt_footprint_bboxes.append([x, y, x + w_, y + h_])
            if 'only_footprint' in ann:            # sticky across annotations, as in the reference
                only_fp = 1 if ann['only_footprint'] == 1 else 0
            gt_bboxes.append(bbox)
            gt_labels.append(self.cat2label[ann['category_id']])
            if only_fp == 0:
                if self.mask_type == 'roof':
                    gt_masks_ann.append(ann['segmentation'])
                elif self.mask_type == 'footprint':
                    gt_masks_ann.append([ann['footprint_mask']])
                else:
                    raise TypeError(f"don't support mask_type={self.mask_type}")
            else:
                gt_masks_ann.append([ann['footprint_mask']])
            gt_roof_masks_ann.append(ann['segmentation'])
            gt_footprint_masks_ann.append([ann['footprint_mask']])
            if 'offset' in ann:
                if self.offset_coordinate == 'rectangle':
                    gt_offsets.append(ann['offset'])
                elif self.offset_coordinate == 'polar':
                    ox, oy = ann['offset']
                    gt_offsets.append([math.sqrt(ox ** 2 + oy ** 2), math.atan2(oy, ox)])
                else:
                    raise RuntimeError(f'do not support this coordinate: {self.offset_coordinate}')
            else:
                gt_offsets.append([0, 0])
            gt_building_heights.append(ann.get('building_height', 0.0))
            if 'offset' in ann and 'building_height' in ann:
                ox, oy = ann['offset']
                gt_angles.append(math.atan2(math.sqrt(ox ** 2 + oy ** 2) * self.resolution,
                                            ann['building_height']))
        if gt_bboxes:
            gt_bboxes = np.array(gt_bboxes, dtype=np.float32)
            gt_roof_bboxes = np.array(gt_roof_bboxes, dtype=np.float32)
            gt_footprint_bboxes = np.array(gt_footprint_bboxes, dtype=np.float32)
            gt_labels = np.array(gt_labels, dtype=np.int64)
            gt_offsets = np.array(gt_offsets, dtype=np.float32)
            gt_building_heights = np.array(gt_building_heights, dtype=np.float32)
            gt_mean_angle = float(np.array(gt_angles, dtype=np.float32).mean())
            only_fp = float(only_fp)
        else:
            gt_bboxes = np.zeros((0, 4), dtype=np.float32)
            gt_roof_bboxes = np.zeros((0, 4), dtype=np.float32)
            gt_footprint_bboxes = np.zeros((0, 4), dtype=np.float32)
            gt_labels = np.array([], dtype=np.int64)
            gt_offsets = np.zeros((0, 2), dtype=np.float32)
            gt_building_heights = np.zeros((0, 2), dtype=np.float32)
            gt_mean_angle = 0.0001
            only_fp = 0
        gt_bboxes_ignore = np.array(gt_bboxes_ignore, dtype=np.float32) if gt_bboxes_ignore \
            else np.zeros((0, 4), dtype=np.float32)
        fn = img_info['filename']
        return dict(bboxes=gt_bboxes, labels=gt_labels, bboxes_ignore=gt_bboxes_ignore,
                    masks=gt_masks_ann, roof_masks=gt_roof_masks_ann,
                    footprint_masks=gt_footprint_masks_ann, seg_map=fn.replace('jpg', 'png'),
                    offsets=gt_offsets, building_heights=gt_building_heights,
                    angle=gt_mean_angle, edge_map=fn.replace('jpg', 'png'),
                    side_face_map=fn.replace('jpg', 'png'), roof_bboxes=gt_roof_bboxes,
                    footprint_bboxes=gt_footprint_bboxes,
                    offset_field=fn.replace('png', 'npy'), only_footprint_flag=only_fp)

    # ---------------------------------------------------------------- custom.py:147-190
    def load_image(self, idx):
        """LoadImageFromFile (loading.py:14-71): uint8 [H,W,3] in BGR order, pinned."""
        from PIL import Image
        info = self.data_infos[idx]
        path = osp.join(self.img_prefix, info['filename']) if self.img_prefix else info['filename']
        with Image.open(path) as im:
            rgb = np.asarray(im.convert('RGB'))
        bgr = torch.from_numpy(np.ascontiguousarray(rgb[:, :, ::-1]))
        return bgr.pin_memory() if torch.cuda.is_available() else bgr

    def prepare_train_img(self, idx, flip=None):
        info = self.data_infos[idx]
        ann = self.get_ann_info(idx)
        if self.load_cfg.get('poly2mask', True) is not True:
            raise NotImplementedError('LoadAnnotations(poly2mask=False) is not on the BONAI path')
        img = self.load_image(idx)
        masks = polygons_to_bitmaps(ann['masks'], info['height'], info['width'], self.device)
        out = self.pipeline(img, ann['bboxes'], ann['labels'], masks, ann['offsets'], flip=flip)
        out['img_metas'].update(filename=info['filename'], ori_filename=info['filename'])
        return out

    def __getitem__(self, idx):
        return self.prepare_train_img(idx)


def build_dataset(cfg, default_args=None):
    """datasets/builder.py:47-75 for the BONAI config: a list of annotation files (one per city,
    bonai_instance.py:33-38) becomes a list of datasets the caller concatenates."""
    if isinstance(cfg.get('ann_file'), (list, tuple)):
        out = []
        for i, f in enumerate(cfg['ann_file']):
            c = dict(cfg)
            c['ann_file'] = f
            if isinstance(cfg.get('img_prefix'), (list, tuple)):
                c['img_prefix'] = cfg['img_prefix'][i]
            out.append(build_from_cfg(c, DATASETS, default_args))
        return out
    return build_from_cfg(dict(cfg), DATASETS, default_args)
