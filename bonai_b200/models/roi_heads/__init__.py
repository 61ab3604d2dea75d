from .roi_extractors import SingleRoIExtractor
from .bbox_heads import BBoxHead, ConvFCBBoxHead, Shared2FCBBoxHead
from .mask_heads import FCNMaskHead
from .attribute_heads import OffsetHeadExpandFeature, OffsetHead
from .standard_roi_head import BaseRoIHead, StandardRoIHead
from .loft_roi_head import LoftRoIHead

__all__ = ['SingleRoIExtractor', 'BBoxHead', 'ConvFCBBoxHead', 'Shared2FCBBoxHead', 'FCNMaskHead',
           'OffsetHeadExpandFeature', 'OffsetHead', 'BaseRoIHead', 'StandardRoIHead', 'LoftRoIHead']
