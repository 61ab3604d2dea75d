"""Where the time of ONE dense launch goes, per CTA (globaltimer stamps written by the kernel when
loft_debug_set_trace is armed): launch -> operands readable -> first stage landed -> first tile's
MMAs issued -> accumulator complete -> first tile stored -> last tile stored.

    python tools/gemm_timeline.py [shape ...]       (LOFT_2CTA=0|1|2 selects the form)
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import _lib as L  # noqa: E402

i32 = ctypes.c_int
dev = 'cuda'
SHAPES = {
    # name: (kind, args)
    'foa': ('conv', (200, 7, 7, 256, 256, 1)),
    'foa_g4': ('conv', (800, 7, 7, 256, 256, 4)),
    'foa_g4_p256': ('conv', (2048, 7, 7, 256, 256, 4)),
    'mask': ('conv', (200, 14, 14, 256, 256, 1)),
    'p2': ('conv', (2, 256, 256, 256, 256, 1)),
    'p4': ('conv', (2, 64, 64, 256, 256, 1)),
    'l3': ('conv', (2, 64, 64, 256, 256, 1)),
    'l4': ('conv', (2, 32, 32, 512, 512, 1)),
    'l3_1x1': ('gemm', (8192, 256, 1024)),
    'l4_1x1': ('gemm', (2048, 2048, 512)),
    'fc2': ('gemm', (2048, 1024, 1024)),
    'fc1': ('gemm', (2048, 12544, 1024)),
    'l3_2d': ('gemm', (8192, 2304, 256)),     # the l3 / l4 3x3 contractions in the 2-D mode
    'l4_2d': ('gemm', (2048, 4608, 512)),
}


def main():
    names = sys.argv[1:] or list(SHAPES)
    lib = L.lib()
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    buf = torch.zeros(16 * 1024, dtype=torch.int64, device=dev)
    print(f'{"shape":12s} {"ctas":>4s} {"tiles":>5s} | {"wait":>6s} {"1st_ld":>6s} {"mma_t0":>6s} '
          f'{"drain":>6s} {"epi_t0":>6s} {"rest":>7s} | {"total":>7s} {"event":>7s}  (us, median CTA; '
          f'total = first entry -> last store)')
    for nm in names:
        kind, a = SHAPES[nm]
        if kind == 'conv':
            N, H, W, Ci, Co, G = a
            x = torch.randn(N, H, W, Ci, device=dev)
            w = torch.randn(G, Co, 3, 3, Ci, device=dev) * 0.02
            y = torch.empty(N, H, W, Co, device=dev)
            b = torch.zeros(G, Co, device=dev)
            e = L.make_epilogue(shift=b, relu=True, round_out=True)

            def launch():
                L.call('conv3x3_fprop_grouped', L.ptr(x), L.ptr(w), L.ptr(y), i32(N), i32(H),
                       i32(W), i32(Ci), i32(Co), i32(G), L.ll(Co * 9 * Ci), L.ll(Co),
                       ctypes.byref(e), L.stream())
            flops = 2.0 * N * H * W * 9 * Ci * Co
        else:
            P, K, Co = a
            x = torch.randn(P, K, device=dev)
            w = torch.randn(Co, K, device=dev) * 0.02
            y = torch.empty(P, Co, device=dev)
            b = torch.zeros(Co, device=dev)
            e = L.make_epilogue(shift=b, relu=True, round_out=True)

            def launch():
                L.call('gemm_fprop', L.ptr(x), L.ptr(w), L.ptr(y), L.ll(P), i32(K), i32(Co),
                       L.ll(K), L.ll(K), L.ll(Co), i32(1), i32(P), ctypes.byref(e), L.stream())
            flops = 2.0 * P * K * Co
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ev_us = e0.elapsed_time(e1) * 100
        buf.zero_()
        lib.loft_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
        launch()
        torch.cuda.synchronize()
        lib.loft_debug_set_trace(ctypes.c_void_p(0))
        t = buf.view(-1, 16).cpu()
        t = t[t[:, 0] > 0].double()
        n_cta = t.shape[0]
        lead = t[t[:, 2] > 0]                   # CTAs that issued MMAs (all, or the pair leaders)
        med = lambda v: float(v.median()) / 1e3
        t_first = float(t[:, 0].min())
        total = (float(t[:, 6].max()) - t_first) / 1e3
        print(f'{nm:12s} {n_cta:4d} {int(t[:, 7].sum()):5d} | {med(t[:, 1] - t[:, 0]):6.2f} '
              f'{med(lead[:, 2] - lead[:, 1]):6.2f} {med(lead[:, 3] - lead[:, 2]):6.2f} '
              f'{med(t[:, 4] - t[:, 1]) - med(lead[:, 3] - lead[:, 1]):6.2f} '
              f'{med(t[:, 5] - t[:, 4]):6.2f} {med(t[:, 6] - t[:, 5]):7.2f} | {total:7.2f} '
              f'{ev_us:7.2f}  {flops / ev_us / 1e6:6.0f} TF/s')
        if lead[:, 9].max() > 0:   # built with -DLOFT_KTRACE: cycle accounting of the first tile's k-loop
            kb = float(lead[:, 12].median())
            print(f'   k-loop cycles per k-block ({int(kb)} k-blocks): issuer total {med(lead[:, 9]) * 1e3 / kb:6.1f} '
                  f'waiting for operands {med(lead[:, 8]) * 1e3 / kb:6.1f} | producer total '
                  f'{med(t[:, 11]) * 1e3 / kb:6.1f} waiting for a free slot {med(t[:, 10]) * 1e3 / kb:6.1f}'
                  f' | epilogue of the first tile {med(t[:, 13]) * 1e3:7.0f} cycles')


if __name__ == '__main__':
    main()
