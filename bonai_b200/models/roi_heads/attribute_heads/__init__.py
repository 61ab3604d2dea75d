from .offset_head_expand_feature import OffsetHeadExpandFeature
from .offset_head import OffsetHead

__all__ = ['OffsetHeadExpandFeature', 'OffsetHead']
