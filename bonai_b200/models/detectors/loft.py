"""LOFT detector (mmdet/models/detectors/loft.py:11-32)."""
from ..builder import DETECTORS
from .two_stage import TwoStageDetector


@DETECTORS.register_module()
class LOFT(TwoStageDetector):
    def __init__(self, backbone, rpn_head, roi_head, train_cfg, test_cfg, neck=None,
                 pretrained=None):
        super().__init__(backbone=backbone, neck=neck, rpn_head=rpn_head, roi_head=roi_head,
                         train_cfg=train_cfg, test_cfg=test_cfg, pretrained=pretrained)
        self.with_vis_feat = True
