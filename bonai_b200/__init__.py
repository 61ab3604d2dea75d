"""bonai_b200 -- B200-native (sm_100a) implementation of the LOFT/FOA training hot path of
jwwangchn/BONAI, behind the reference's registry / config surface."""
__version__ = '0.1.0'

from .config import Config, ConfigDict, DictAction  # noqa: F401
from .registry import Registry, build_from_cfg  # noqa: F401
