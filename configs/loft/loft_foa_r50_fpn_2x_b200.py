"""LOFT + FOA, ResNet-50-FPN, 2x schedule -- self-contained restatement of the reference's
configs/loft_foa/loft_foa_r50_fpn_2x_bonai.py (+ its four _base_ files) for boxes where the
reference tree is not mounted.  tests/test_config.py asserts that the model / train_cfg / test_cfg
/ optimizer sections are equal to the reference's when /root/reference is present.
(`pretrained` is None here: there is no network to fetch torchvision://resnet50.)"""


def _roi_extractor(size):
    return dict(type='SingleRoIExtractor', out_channels=256, featmap_strides=[4, 8, 16, 32],
                roi_layer=dict(type='RoIAlign', output_size=size, sampling_ratio=0))


def _xywh_coder(stds):
    return dict(type='DeltaXYWHBBoxCoder', target_means=[0.0] * 4, target_stds=list(stds))


def _max_iou(pos, neg, min_pos):
    return dict(type='MaxIoUAssigner', pos_iou_thr=pos, neg_iou_thr=neg, min_pos_iou=min_pos,
                match_low_quality=True, ignore_iof_thr=-1, gpu_assign_thr=512)


def _sampler(num, frac, add_gt):
    return dict(type='RandomSampler', num=num, pos_fraction=frac, neg_pos_ub=-1,
                add_gt_as_proposals=add_gt)


_proposal = dict(nms_across_levels=False, nms_pre=3000, nms_post=3000, max_num=3000, nms_thr=0.7,
                 min_bbox_size=0)

model = dict(
    type='LOFT',
    pretrained=None,
    backbone=dict(type='ResNet', depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                  norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch'),
    neck=dict(type='FPN', in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5),
    rpn_head=dict(
        type='RPNHead', in_channels=256, feat_channels=256,
        anchor_generator=dict(type='AnchorGenerator', scales=[8], ratios=[0.5, 1.0, 2.0],
                              strides=[4, 8, 16, 32, 64]),
        bbox_coder=_xywh_coder([1.0, 1.0, 1.0, 1.0]),
        loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', loss_weight=1.0)),
    roi_head=dict(
        type='LoftRoIHead',
        bbox_roi_extractor=_roi_extractor(7),
        bbox_head=dict(
            type='Shared2FCBBoxHead', in_channels=256, fc_out_channels=1024, roi_feat_size=7,
            num_classes=1, bbox_coder=_xywh_coder([0.1, 0.1, 0.2, 0.2]), reg_class_agnostic=False,
            loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
            loss_bbox=dict(type='L1Loss', loss_weight=1.0)),
        mask_roi_extractor=_roi_extractor(14),
        mask_head=dict(type='FCNMaskHead', num_convs=4, in_channels=256, conv_out_channels=256,
                       num_classes=1,
                       loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0)),
        offset_roi_extractor=_roi_extractor(7),
        offset_head=dict(type='OffsetHeadExpandFeature', expand_feature_num=4, share_expand_fc=True,
                         rotations=[0, 90, 180, 270], num_fcs=2, fc_out_channels=1024, num_convs=10,
                         loss_offset=dict(type='SmoothL1Loss', loss_weight=16.0))))

train_cfg = dict(
    rpn=dict(assigner=_max_iou(0.7, 0.3, 0.3), sampler=_sampler(512, 0.5, False),
             allowed_border=-1, pos_weight=-1, debug=False),
    rpn_proposal=dict(_proposal),
    rcnn=dict(assigner=_max_iou(0.5, 0.5, 0.5), sampler=_sampler(1024, 0.25, True), mask_size=28,
              pos_weight=-1, debug=False))

test_cfg = dict(
    rpn=dict(_proposal),
    rcnn=dict(score_thr=0.05, nms=dict(type='soft_nms', iou_threshold=0.5), max_per_img=2000,
              mask_thr_binary=0.5))

# schedule (SGD "for 4 GPUs": lr 0.02/4), runtime
optimizer = dict(type='SGD', lr=0.005, momentum=0.9, weight_decay=0.0001)
optimizer_config = dict(grad_clip=dict(max_norm=35, norm_type=2))
lr_config = dict(policy='step', warmup='linear', warmup_iters=300, warmup_ratio=0.001, step=[16, 22])
total_epochs = 24
checkpoint_config = dict(interval=1)
log_config = dict(interval=10, hooks=[dict(type='TextLoggerHook')])
dist_params = dict(backend='nccl')
log_level = 'INFO'
load_from = None
resume_from = None
workflow = [('train', 1)]

# post-load training pipeline (same values as the reference's
# configs/_base_/datasets/bonai_instance.py:3-17); bonai_b200.datasets.GpuTrainPipeline.from_cfg
# runs RandomFlip / Normalize / Pad / formatting on the device
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
train_pipeline = [
    dict(type='LoadImageFromFile'),
    dict(type='LoadAnnotations', with_bbox=True, with_mask=True, with_offset=True),
    dict(type='Resize', img_scale=(1024, 1024), keep_ratio=True),
    dict(type='RandomFlip', flip_ratio=0.5, direction=['horizontal', 'vertical']),
    dict(type='Normalize', **img_norm_cfg),
    dict(type='Pad', size_divisor=32),
    dict(type='DefaultFormatBundle'),
    dict(type='Collect', keys=['img', 'gt_bboxes', 'gt_labels', 'gt_masks', 'gt_offsets']),
]
# the five city files of configs/_base_/datasets/bonai_instance.py:31-48
data_root = 'data/BONAI/'
cities = ['shanghai', 'beijing', 'jinan', 'haerbin', 'chengdu']
train_ann_file = [data_root + 'coco/bonai_{}_trainval.json'.format(c) for c in cities]
img_prefix = [data_root + 'trainval/images/' for _ in cities]
data = dict(samples_per_gpu=2, workers_per_gpu=2,
            train=dict(type='BONAI', ann_file=train_ann_file, img_prefix=img_prefix,
                       bbox_type='building', mask_type='roof', pipeline=train_pipeline))
