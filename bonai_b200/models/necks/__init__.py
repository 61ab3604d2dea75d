from .fpn import FPN
from .hrfpn import HRFPN

__all__ = ['FPN', 'HRFPN']
