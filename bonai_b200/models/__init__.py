from .builder import (BACKBONES, NECKS, ROI_EXTRACTORS, SHARED_HEADS, HEADS, LOSSES, DETECTORS,
                      build_backbone, build_neck, build_roi_extractor, build_shared_head,
                      build_head, build_loss, build_detector)
from .losses import *  # noqa: F401,F403
from .backbones import *  # noqa: F401,F403
from .necks import *  # noqa: F401,F403
from .dense_heads import *  # noqa: F401,F403
from .roi_heads import *  # noqa: F401,F403
from .detectors import *  # noqa: F401,F403
