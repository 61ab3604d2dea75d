from ..builder import HEADS, ROI_EXTRACTORS, build_head, build_loss, build_roi_extractor, \
    build_shared_head  # noqa: F401
