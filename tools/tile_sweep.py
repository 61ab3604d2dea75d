"""Tile-N sweep for the small-pixel backbone shapes: run with LOFT_TILE_N=256|128|64 (or unset)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
dev = 'cuda'
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
out = []
def conv3(name, N, H, W, Ci, Co):
    x = torch.randn(N, H, W, Ci, device=dev); w = torch.randn(Co, 3, 3, Ci, device=dev) * .02
    y = torch.empty(N, H, W, Co, device=dev); dx = torch.empty_like(x)
    out.append((name + ' fwd', timeit(lambda: L.call('conv3x3_fprop', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H), i32(W), i32(Ci), i32(Co), None, L.stream()))))
    out.append((name + ' dgrad', timeit(lambda: L.call('conv3x3_dgrad', L.ptr(y), L.ptr(w), L.ptr(dx), i32(N), i32(H), i32(W), i32(Ci), i32(Co), None, L.stream()))))
def gemm(name, P, K, Co):
    x = torch.randn(P, K, device=dev); w = torch.randn(Co, K, device=dev) * .02
    y = torch.empty(P, Co, device=dev); dx = torch.empty_like(x)
    out.append((name + ' fwd', timeit(lambda: L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co), L.ll(K), L.ll(K), L.ll(Co), i32(1), i32(P), None, L.stream()))))
    out.append((name + ' dgrad', timeit(lambda: L.call('gemm_dgrad', L.ptr(y), L.ptr(w), L.ptr(dx), L.ll(P), i32(K), i32(Co), L.ll(Co), L.ll(K), L.ll(K), None, L.stream()))))
conv3('l2 3x3 128', 2, 128, 128, 128, 128)
conv3('l3 3x3 256', 2, 64, 64, 256, 256)
conv3('l4 3x3 512', 2, 32, 32, 512, 512)
conv3('fpn P4 3x3', 2, 64, 64, 256, 256)
conv3('fpn P5 3x3', 2, 32, 32, 256, 256)
gemm('l3 1x1 1024->256', 8192, 1024, 256); gemm('l3 1x1 256->1024', 8192, 256, 1024)
gemm('l4 1x1 2048->512', 2048, 2048, 512); gemm('l4 1x1 512->2048', 2048, 512, 2048)
gemm('l4 ds 1024->2048', 2048, 1024, 2048); gemm('l3 ds 512->1024', 8192, 512, 1024)
gemm('l2 1x1 512->128', 32768, 512, 128); gemm('l2 1x1 128->512', 32768, 128, 512)
gemm('bbox fc1', 2048, 12544, 1024); gemm('foa fc1', 812, 12544, 1024)
print(os.environ.get('LOFT_TILE_N', 'auto'), ' '.join(f'{n}={t:.1f}' for n, t in out))
