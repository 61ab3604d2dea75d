"""The dominant launch alone (FPN/RPN P2 3x3 conv, 2x256x256x256 -> 256) for `ncu --set full`."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
N, H, W, C = 2, 256, 256, 256
x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(C, 3, 3, C, device='cuda') * 0.02
b = torch.randn(C, device='cuda'); y = torch.empty(N, H, W, C, device='cuda')
e = L.make_epilogue(shift=b, round_out=True)
for _ in range(6):
    L.call('conv3x3_fprop', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H), i32(W), i32(C), i32(C),
           ctypes.byref(e), L.stream())
torch.cuda.synchronize()
