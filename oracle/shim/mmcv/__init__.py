"""Import shim standing in for mmcv-full==1.0.5 (absent here; pinned by mmdet/__init__.py:18-26).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference under /root/reference execute on CPU so
that golden vectors can be generated (oracle/make_golden.py) and the portable restatement
(oracle/loft_cpu.py) can be validated.  Nothing in bonai_b200/ imports it.  Semantics of the few
real pieces follow mmcv v1.0.5 as recorded in SURVEY.md Appendix A.
"""
__version__ = '1.0.5'

from .config import Config, ConfigDict, DictAction  # noqa: F401
from . import utils, cnn, ops, runner, parallel, image  # noqa: F401
from .image import (imflip, imnormalize, imnormalize_, impad, impad_to_multiple,  # noqa: F401
                    imrescale, imresize, rescale_size)
from .utils import is_tuple_of, is_list_of, is_seq_of, is_str  # noqa: F401
from _shim_dummy import install_getattr as _ig

_ig(globals(), 'mmcv')
