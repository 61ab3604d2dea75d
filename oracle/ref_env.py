"""Make the UNMODIFIED reference (/root/reference) importable over the shim.  Only usable in the
build container (the GPU box has no /root/reference); used by oracle/make_golden.py and by the
CPU tests that validate the portable restatement against the reference itself."""
import os
import sys

REF_ROOT = os.environ.get('LOFT_REFERENCE_ROOT', '/root/reference')
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'mmdet'))


def activate():
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    for p in (REF_ROOT, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mmdet  # noqa: F401
    return mmdet
