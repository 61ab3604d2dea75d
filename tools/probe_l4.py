"""layer4 3x3 conv (512->512 on 2x32x32 pixels) alone, for ncu."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
N, H, W, C = 2, 32, 32, 512
x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(C, 3, 3, C, device='cuda') * 0.02
y = torch.empty(N, H, W, C, device='cuda')
for _ in range(6):
    L.call('conv3x3_fprop', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H), i32(W), i32(C), i32(C), None, L.stream())
torch.cuda.synchronize()
