"""Micro-benchmark of every distinct dense shape of the LOFT step (fwd / dgrad / wgrad):
TFLOP/s per launch and share of the summed time -> where the GEMM time goes."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L
i32 = ctypes.c_int
dev = 'cuda'

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

rows = []
def conv3(name, N, H, W, Ci, Co, count, bwd=True):
    x = torch.randn(N, H, W, Ci, device=dev); w = torch.randn(Co, 3, 3, Ci, device=dev) * 0.02
    y = torch.empty(N, H, W, Co, device=dev); dx = torch.empty_like(x); dw = torch.zeros_like(w)
    fl = 2.0 * N * H * W * 9 * Ci * Co
    t = timeit(lambda: L.call('conv3x3_fprop', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H), i32(W), i32(Ci), i32(Co), None, L.stream()))
    rows.append((name + ' fwd', fl, t, count))
    if bwd:
        t = timeit(lambda: L.call('conv3x3_dgrad', L.ptr(y), L.ptr(w), L.ptr(dx), i32(N), i32(H), i32(W), i32(Ci), i32(Co), None, L.stream()))
        rows.append((name + ' dgrad', fl, t, count))
        t = timeit(lambda: L.call('conv3x3_wgrad', L.ptr(y), L.ptr(x), L.ptr(dw), i32(N), i32(H), i32(W), i32(Ci), i32(Co), L.stream()))
        rows.append((name + ' wgrad', fl, t, count))

def gemm(name, P, K, Co, count, bwd=True):
    x = torch.randn(P, K, device=dev); w = torch.randn(Co, K, device=dev) * 0.02
    y = torch.empty(P, Co, device=dev); dx = torch.empty_like(x); dw = torch.zeros_like(w)
    fl = 2.0 * P * K * Co
    t = timeit(lambda: L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co), L.ll(K), L.ll(K), L.ll(Co), i32(1), i32(P), None, L.stream()))
    rows.append((name + ' fwd', fl, t, count))
    if bwd:
        t = timeit(lambda: L.call('gemm_dgrad', L.ptr(y), L.ptr(w), L.ptr(dx), L.ll(P), i32(K), i32(Co), L.ll(Co), L.ll(K), L.ll(K), None, L.stream()))
        rows.append((name + ' dgrad', fl, t, count))
        if Co % 32 == 0:
            t = timeit(lambda: L.call('gemm_wgrad', L.ptr(y), L.ptr(x), L.ptr(dw), L.ll(P), i32(K), i32(Co), L.ll(Co), L.ll(K), L.ll(K), L.stream()))
            rows.append((name + ' wgrad', fl, t, count))

P1, P2, P3, P4, P5 = 131072, 32768, 8192, 2048, 512
gemm('stem im2col 148->64', 524288, 148, 64, 1, bwd=False)
gemm('l1 1x1 64->64', P1, 64, 64, 1, bwd=False); gemm('l1 1x1 256->64', P1, 256, 64, 2, bwd=False)
gemm('l1 1x1 64->256', P1, 64, 256, 4, bwd=False); conv3('l1 3x3 64', 2, 256, 256, 64, 64, 3, bwd=False)
gemm('l2 1x1 256->128', P1, 256, 128, 1); gemm('l2 1x1 512->128', P2, 512, 128, 3); gemm('l2 1x1 128->512', P2, 128, 512, 4)
gemm('l2 ds 256->512', P2, 256, 512, 1); gemm('l2 3x3s2 im2col', P2, 1152, 128, 1); conv3('l2 3x3 128', 2, 128, 128, 128, 128, 3)
gemm('l3 1x1 512->256', P2, 512, 256, 1); gemm('l3 1x1 1024->256', P3, 1024, 256, 5); gemm('l3 1x1 256->1024', P3, 256, 1024, 6)
gemm('l3 ds 512->1024', P3, 512, 1024, 1); gemm('l3 3x3s2 im2col', P3, 2304, 256, 1); conv3('l3 3x3 256', 2, 64, 64, 256, 256, 5)
gemm('l4 1x1 1024->512', P3, 1024, 512, 1); gemm('l4 1x1 2048->512', P4, 2048, 512, 2); gemm('l4 1x1 512->2048', P4, 512, 2048, 3)
gemm('l4 ds 1024->2048', P4, 1024, 2048, 1); gemm('l4 3x3s2 im2col', P4, 4608, 512, 1); conv3('l4 3x3 512', 2, 32, 32, 512, 512, 2)
for nm, P, C in (('lat0', P1, 256), ('lat1', P2, 512), ('lat2', P3, 1024), ('lat3', P4, 2048)):
    gemm('fpn ' + nm, P, C, 256, 1)
for nm, H, c in (('P2', 256, 2), ('P3', 128, 2), ('P4', 64, 2), ('P5', 32, 2)):
    conv3('fpn/rpn 3x3 ' + nm, 2, H, H, 256, 256, c)
conv3('rpn 3x3 P6', 2, 16, 16, 256, 256, 1)
gemm('rpn head P2 256->16', P1, 256, 16, 1)
gemm('bbox fc1 12544->1024', 2048, 12544, 1024, 1); gemm('bbox fc2', 2048, 1024, 1024, 1)
conv3('mask conv 14x14 P=200', 200, 14, 14, 256, 256, 4); gemm('mask deconv', 200 * 196, 256, 1024, 1)
gemm('mask logits', 200 * 784, 256, 4, 1)
conv3('foa conv 7x7 P=200', 200, 7, 7, 256, 256, 40); gemm('foa fc1 P=800', 800, 12544, 1024, 1); gemm('foa fc2', 800, 1024, 1024, 1)
tot = sum(t * c for _, _, t, c in rows)
print(f'{"shape":34s} {"us":>8s} {"TF/s":>7s} {"n":>3s} {"share":>6s}')
for name, fl, t, c in sorted(rows, key=lambda r: -r[2] * r[3]):
    print(f'{name:34s} {t:8.1f} {fl / t / 1e6:7.1f} {c:3d} {100 * t * c / tot:5.1f}%')
print('total ms', tot / 1e3)
