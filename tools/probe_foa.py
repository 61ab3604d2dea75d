"""The grouped FOA 3x3 conv (4 branches x P RoIs of 7x7x256 -> 256) alone, for ncu: fprop, dgrad,
wgrad, a few launches each.  P from argv (default 105 = positives per tile of the bench run)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
P = int(sys.argv[1]) if len(sys.argv) > 1 else 105
G, C, S = 4, 256, 7
x = torch.randn(G * P * 2, S, S, C, device='cuda')          # 2 tiles per step
w = torch.randn(G, C, 3, 3, C, device='cuda') * 0.02
b = torch.zeros(G, C, device='cuda')
y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.zeros_like(w)
e = L.make_epilogue(shift=b, relu=True, round_out=True)
e2 = L.make_epilogue(round_out=True, mask=x)
N = x.shape[0]
for _ in range(4):
    L.call('conv3x3_fprop_grouped', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(S), i32(S), i32(C),
           i32(C), i32(G), L.ll(C * 9 * C), L.ll(C), ctypes.byref(e), L.stream())
    L.call('conv3x3_dgrad_grouped', L.ptr(y), L.ptr(w), L.ptr(dx), i32(N), i32(S), i32(S), i32(C),
           i32(C), i32(G), L.ll(C * 9 * C), ctypes.byref(e2), L.stream())
    L.call('conv3x3_wgrad_grouped', L.ptr(y), L.ptr(x), L.ptr(dw), i32(N), i32(S), i32(S), i32(C),
           i32(C), i32(G), L.ll(C * 9 * C), L.stream())
torch.cuda.synchronize()
