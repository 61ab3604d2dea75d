"""GPU bring-up of the tcgen05 GEMM/conv core against torch fp32 (TF32 disabled) references.

Usage:  python tools/bringup_gemm.py <case> [variant]     (one case per process, so a hang in one
        case cannot take the others down; tools/bringup_all.sh wraps each in `timeout`).
"""
import ctypes
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonai_b200 import _lib as L  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def desc(lbo, sbo, lt=1):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | (1 << 46) | (lt << 61)


VARIANTS = {
    # name: (a_desc, b_desc, a_kstep, b_kstep) overrides for MN-major operands; None = default
    'default': None,
    'swap': (desc(512, 4096), desc(512, 4096), -1, -1),
    'sbo1024': (desc(4096, 1024), desc(4096, 1024), -1, -1),
}


def relerr(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def run(case, variant):
    dev = 'cuda'
    g = torch.Generator(device='cpu').manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    i32 = ctypes.c_int
    mn_a = mn_b = False
    out = {}
    if case.startswith('fprop2d'):
        P, K, Co = {'fprop2d': (1000, 256, 256), 'fprop2d_small': (4096, 64, 64),
                    'fprop2d_big': (131072, 256, 256), 'fprop2d_tiny': (50, 1024, 8), 'fprop2d_wide': (2048, 1024, 1024)}[case]
        x, w, b = rnd(P, K), rnd(Co, K) * 0.1, rnd(Co)
        res = rnd(P, Co)
        y = torch.empty(P, Co, device=dev)
        raw = torch.empty(P, Co, device=dev)
        e = L.make_epilogue(raw_out=raw, shift=b, residual=res, relu=True)
        L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co), L.ll(K),
               L.ll(K), L.ll(Co), i32(1), i32(P), ctypes.byref(e), L.stream())
        torch.cuda.synchronize()
        ref_raw = x @ w.t()
        ref = torch.relu(ref_raw + b + res)
        out['raw'] = relerr(raw, ref_raw)
        out['y'] = relerr(y, ref)
    elif case.startswith('fprop_conv'):
        N, H, W, Ci, Co = {'fprop_conv': (2, 32, 32, 64, 128), 'fprop_conv7': (11, 7, 7, 256, 256),
                           'fprop_conv14': (3, 14, 14, 256, 256),
                           'fprop_conv_odd': (2, 20, 25, 32, 64),
                           'fprop_conv_big': (2, 256, 256, 256, 256)}[case]
        x, w, b = rnd(N, Ci, H, W), rnd(Co, Ci, 3, 3) * 0.05, rnd(Co)
        xn, wn = nhwc(x), nhwc(w)
        y = torch.empty(N, H, W, Co, device=dev)
        e = L.make_epilogue(shift=b, relu=True)
        L.call('conv3x3_fprop', L.ptr(xn), L.ptr(wn), L.ptr(y), i32(N), i32(H), i32(W), i32(Ci),
               i32(Co), ctypes.byref(e), L.stream())
        torch.cuda.synchronize()
        ref = nhwc(torch.relu(F.conv2d(x, w, b, padding=1)))
        out['y'] = relerr(y, ref)
    elif case.startswith('dgrad2d'):
        mn_a = True
        P, Ci, Co = {'dgrad2d': (1000, 256, 128), 'dgrad2d_tiny': (300, 1024, 8)}[case]
        dy, w = rnd(P, Co), rnd(Co, Ci) * 0.1
        mask = rnd(P, Ci)
        dx = torch.empty(P, Ci, device=dev)
        e = L.make_epilogue(mask=mask)
        set_variant(variant, mn_a, mn_b)
        L.call('gemm_dgrad', L.ptr(dy), L.ptr(w), L.ptr(dx), L.ll(P), i32(Ci), i32(Co), L.ll(Co),
               L.ll(Ci), L.ll(Ci), ctypes.byref(e), L.stream())
        torch.cuda.synchronize()
        ref = (dy @ w) * (mask > 0)
        out['dx'] = relerr(dx, ref)
    elif case.startswith('dgrad_conv'):
        mn_a = True
        N, H, W, Ci, Co = {'dgrad_conv': (2, 32, 32, 64, 128),
                           'dgrad_conv7': (11, 7, 7, 256, 256)}[case]
        dy, w = rnd(N, Co, H, W), rnd(Co, Ci, 3, 3) * 0.05
        dyn, wn = nhwc(dy), nhwc(w)
        dx = torch.empty(N, H, W, Ci, device=dev)
        set_variant(variant, mn_a, mn_b)
        L.call('conv3x3_dgrad', L.ptr(dyn), L.ptr(wn), L.ptr(dx), i32(N), i32(H), i32(W), i32(Ci),
               i32(Co), None, L.stream())
        torch.cuda.synchronize()
        ref = nhwc(F.conv_transpose2d(dy, w, padding=1))
        out['dx'] = relerr(dx, ref)
    elif case.startswith('wgrad2d'):
        mn_a = mn_b = True
        P, Ci, Co = {'wgrad2d': (1000, 256, 128), 'wgrad2d_wide': (2048, 1024, 1024)}[case]
        dy, x = rnd(P, Co), rnd(P, Ci)
        dw = torch.zeros(Co, Ci, device=dev)
        set_variant(variant, mn_a, mn_b)
        L.call('gemm_wgrad', L.ptr(dy), L.ptr(x), L.ptr(dw), L.ll(P), i32(Ci), i32(Co), L.ll(Co),
               L.ll(Ci), L.ll(Ci), L.stream())
        torch.cuda.synchronize()
        ref = dy.t() @ x
        out['dw'] = relerr(dw, ref)
    elif case.startswith('wgrad_conv'):
        mn_a = mn_b = True
        N, H, W, Ci, Co = {'wgrad_conv': (2, 32, 32, 64, 128),
                           'wgrad_conv7': (11, 7, 7, 256, 256),
                           'wgrad_conv14': (3, 14, 14, 256, 256)}[case]
        dy, x = rnd(N, Co, H, W), rnd(N, Ci, H, W)
        dyn, xn = nhwc(dy), nhwc(x)
        dw = torch.zeros(Co, 3, 3, Ci, device=dev)
        set_variant(variant, mn_a, mn_b)
        L.call('conv3x3_wgrad', L.ptr(dyn), L.ptr(xn), L.ptr(dw), i32(N), i32(H), i32(W), i32(Ci),
               i32(Co), L.stream())
        torch.cuda.synchronize()
        ref = torch.nn.grad.conv2d_weight(x, (Co, Ci, 3, 3), dy, padding=1)
        out['dw'] = relerr(dw, nhwc(ref))
    else:
        raise SystemExit(f'unknown case {case}')
    ok = all(v[0] < 3e-3 for v in out.values())
    print(json.dumps({'case': case, 'variant': variant, 'ok': ok,
                      'err': {k: [round(v[0], 6), round(v[1], 5)] for k, v in out.items()}}))
    return ok


def set_variant(variant, mn_a, mn_b):
    v = VARIANTS[variant]
    if v is None:
        return
    a_desc, b_desc, a_k, b_k = v
    L.lib().loft_debug_set_desc(L.ll(a_desc if mn_a else -1), L.ll(b_desc if mn_b else -1),
                                L.ll(a_k if mn_a else -1), L.ll(b_k if mn_b else -1), L.ll(-1))


if __name__ == '__main__':
    ok = run(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'default')
    sys.exit(0 if ok else 1)
