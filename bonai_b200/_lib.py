"""ctypes binding of ``libloft_b200.so`` (the C ABI declared in ``include/loft_b200.h``).

There is deliberately no fallback: if the shared library is missing, or a call is made without a
CUDA device, the product path raises.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C bonai_b200/csrc``.
"""
import contextlib
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LOFT_LIB_PATH') or os.path.join(_HERE, 'libloft_b200.so')   # override: A/B builds

_lib = None


class LoftError(RuntimeError):
    pass


class Epilogue(ctypes.Structure):
    """Mirror of ``loft_epilogue_t``."""
    _fields_ = [
        ('raw_out', ctypes.c_void_p),
        ('scale', ctypes.c_void_p),
        ('shift', ctypes.c_void_p),
        ('residual', ctypes.c_void_p),
        ('mask', ctypes.c_void_p),
        ('ldr', ctypes.c_longlong),
        ('res_upsample2x', ctypes.c_int),
        ('relu', ctypes.c_int),
        ('deconv_shuffle', ctypes.c_int),
        ('round_out', ctypes.c_int),
        ('colsum', ctypes.c_void_p),
        ('colsum2', ctypes.c_void_p),
        ('colsum_gstride', ctypes.c_longlong),
    ]


def lib():
    """Load (once) and return the ctypes handle; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LoftError(
                f'{LIB_PATH} is not built; run `make -C {_HERE}/csrc` (needs nvcc, sm_100a). '
                'There is no CPU fallback for the LOFT hot path.')
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.loft_last_error.restype = ctypes.c_char_p
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ll(v):
    return ctypes.c_longlong(int(v))


def f32(v):
    return ctypes.c_float(float(v))


LAUNCHES = [0]          # kernels launched through this binding (bench.py reads / resets it)
_KERNELS_PER_CALL = {'iou_assign': 2, 'grad_sqnorm': 1, 'poly_rasterize': 2, 'rpn_targets': 3}


class Copy2D(ctypes.Structure):
    """Mirror of ``loft_copy2d_t``."""
    _fields_ = [('src', ctypes.c_void_p), ('lds', ctypes.c_longlong), ('dst', ctypes.c_void_p),
                ('ldd', ctypes.c_longlong), ('rows', ctypes.c_longlong), ('cols', ctypes.c_int),
                ('accumulate', ctypes.c_int), ('round_tf32', ctypes.c_int)]


BATCH = None            # inside `batched_copies()`: the copy2d calls collected so far


@contextlib.contextmanager
def batched_copies():
    """Collect every ``call('copy2d', ...)`` issued inside the block and run them as ONE
    ``loft_copy2d_multi`` launch at its end.  Only for copies that are independent of each other
    and of every other kernel launched inside the block (weight re-packs, gradient scatters)."""
    global BATCH
    if BATCH is not None or RECORD is not None:          # nested, or a launch program is recording
        yield
        return
    BATCH = []
    try:
        yield
    finally:
        items, BATCH = BATCH, None
        if items:
            arr = (Copy2D * len(items))()
            for a, (src, lds, dst, ldd, rows, cols, acc, rnd, st) in zip(arr, items):
                a.src, a.lds, a.dst, a.ldd = src.value, lds.value, dst.value, ldd.value
                a.rows, a.cols, a.accumulate, a.round_tf32 = rows.value, cols.value, acc.value, \
                    rnd.value
                assert st.value == items[0][-1].value, 'batched copies must share one stream'
            call('copy2d_multi', arr, ctypes.c_int(len(items)), items[0][-1])


RECORD = None           # set to a list to record (fn, args) of every call (trunk.Tape)
TRACE = None            # set to a list to record (name, int args, start event, end event) per call


def _arg_summary(args):
    out = []
    for a in args:
        if isinstance(a, (ctypes.c_int, ctypes.c_longlong)):
            out.append(int(a.value))
    return tuple(out)


def call(name, *args):
    """Call ``loft_<name>`` and raise LoftError on a non-zero return code."""
    if BATCH is not None and name == 'copy2d':
        BATCH.append(args)
        return
    fn = getattr(lib(), 'loft_' + name)
    if RECORD is not None:
        RECORD.append((name, fn, args))
    if TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        TRACE.append((name, _arg_summary(args), e0, e1))
        if rc != 0:
            msg = lib().loft_last_error()
            raise LoftError(f'loft_{name} failed ({rc}): {msg.decode() if msg else ""}')
        return
    if name == 'nms_sorted':
        LAUNCHES[0] += 3 * int(args[2].value) + 1      # per image: fill+max+mask; one scan
    elif name == 'nms_segmented':
        LAUNCHES[0] += 2 * int(args[4].value) + 3      # per image: fill+max; mask, scan, compact
    else:
        LAUNCHES[0] += _KERNELS_PER_CALL.get(name, 1)
    rc = fn(*args)
    if rc != 0:
        msg = lib().loft_last_error()
        raise LoftError(f'loft_{name} failed ({rc}): {msg.decode() if msg else ""}')


def make_epilogue(raw_out=None, scale=None, shift=None, residual=None, mask=None, ldr=0,
                  res_upsample2x=False, relu=False, deconv_shuffle=False, round_out=False,
                  colsum=None, colsum2=None, colsum_gstride=0):
    e = Epilogue()
    e.raw_out = raw_out.data_ptr() if raw_out is not None else None
    e.scale = scale.data_ptr() if scale is not None else None
    e.shift = shift.data_ptr() if shift is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.mask = mask.data_ptr() if mask is not None else None
    e.ldr = int(ldr)
    e.res_upsample2x = int(res_upsample2x)        # 0 none, 1 nearest x2, 2 zero-stuffed x2
    e.relu = int(bool(relu))
    e.deconv_shuffle = int(bool(deconv_shuffle))
    e.round_out = int(bool(round_out))
    e.colsum = colsum.data_ptr() if colsum is not None else None
    e.colsum2 = colsum2.data_ptr() if colsum2 is not None else None
    e.colsum_gstride = int(colsum_gstride)
    return e
