"""Epilogue variants of the dgrad launches in isolation: plain / +mask / +mask+colsum."""
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

def variants(x, cs, gstride=0):
    return [('plain', L.make_epilogue(round_out=True)),
            ('mask', L.make_epilogue(round_out=True, mask=x)),
            ('colsum', L.make_epilogue(round_out=True, colsum=cs, colsum_gstride=gstride)),
            ('mask+colsum', L.make_epilogue(round_out=True, mask=x, colsum=cs, colsum_gstride=gstride))]

def gemm_dgrad(P, Cin, Cout):
    dy = torch.randn(P, Cout, device=dev); w = torch.randn(Cout, Cin, device=dev)
    dx = torch.empty(P, Cin, device=dev); x = torch.randn(P, Cin, device=dev)
    cs = torch.zeros(Cin, device=dev)
    for name, e in variants(x, cs):
        t = timeit(lambda: L.call('gemm_dgrad', L.ptr(dy), L.ptr(w), L.ptr(dx), L.ll(P), i32(Cin), i32(Cout),
                                  L.ll(Cout), L.ll(Cin), L.ll(Cin), ctypes.byref(e), L.stream()))
        print(f'gemm_dgrad P={P} Cin={Cin} Cout={Cout} {name:12s} {t:8.1f} us')

def conv_dgrad(N, H, W, C, G):
    dy = torch.randn(N, H, W, C, device=dev); w = torch.randn(G, C, 3, 3, C, device=dev) * .02
    dx = torch.empty(N, H, W, C, device=dev); x = torch.randn(N, H, W, C, device=dev)
    cs = torch.zeros(G, C, device=dev)
    for name, e in variants(x, cs, C):
        t = timeit(lambda: L.call('conv3x3_dgrad_grouped', L.ptr(dy), L.ptr(w), L.ptr(dx), i32(N), i32(H), i32(W),
                                  i32(C), i32(C), i32(G), L.ll(9 * C * C), ctypes.byref(e), L.stream()))
        print(f'conv3x3_dgrad N={N} {H}x{W} C={C} G={G} {name:12s} {t:8.1f} us')

gemm_dgrad(159152, 256, 4)
gemm_dgrad(131072, 256, 16)
gemm_dgrad(812, 12544, 1024)
gemm_dgrad(39788, 256, 1024)
conv_dgrad(812, 7, 7, 256, 4)
conv_dgrad(203, 14, 14, 256, 1)
