"""An epilogue-bound launch alone (layer1 conv3: 1x1 64->256 on 2x256x256 pixels, + residual + ReLU)
for `ncu --set full --import-source on`."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
P, K, Co = 131072, 64, 256
x = torch.randn(P, K, device='cuda'); w = torch.randn(Co, K, device='cuda') * 0.05
b = torch.randn(Co, device='cuda'); y = torch.empty(P, Co, device='cuda'); r = torch.randn(P, Co, device='cuda')
mode = sys.argv[1] if len(sys.argv) > 1 else 'res'
e = L.make_epilogue(shift=b, residual=r if mode == 'res' else None, ldr=Co, relu=True, round_out=True)
for _ in range(6):
    L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co), L.ll(K), L.ll(K), L.ll(Co),
           i32(256), i32(256), ctypes.byref(e), L.stream())
torch.cuda.synchronize()
