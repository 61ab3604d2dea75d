from .base import BaseDetector
from .two_stage import TwoStageDetector
from .loft import LOFT

__all__ = ['BaseDetector', 'TwoStageDetector', 'LOFT']
