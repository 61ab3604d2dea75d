from .anchor import AnchorGenerator, build_anchor_generator, anchor_inside_flags, images_to_levels
from .bbox import (MaxIoUAssigner, AssignResult, RandomSampler, SamplingResult, DeltaXYWHBBoxCoder,
                   DeltaXYOffsetCoder, BboxOverlaps2D, bbox_overlaps, bbox2roi, roi2bbox,
                   bbox2result, build_assigner, build_sampler, build_bbox_coder,
                   build_iou_calculator, bbox2delta, delta2bbox, offset2delta, delta2offset)
from .mask import BitmapMasks, mask_target, encode_mask_results
from .utils import multi_apply, unmap

__all__ = [k for k in dir() if not k.startswith('_')]
