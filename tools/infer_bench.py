#!/usr/bin/env python
"""Inference throughput sweep (BASELINE.json configs[4]): LOFT R50-FPN simple_test on synthetic
1024x1024 tiles, "batch" 1 -> 32 (the reference's forward_test asserts batch size 1,
detectors/base.py:141-143, so a batch is a loop of tiles), with the RoIAlign / NMS latency
breakdown measured by CUDA events around the C-ABI calls.

    python tools/infer_bench.py [--batches 1 2 4 8 16 32] [--size 1024]
Prints one JSON line per batch size.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import Config, _lib as L  # noqa: E402
from bonai_b200.models import build_detector  # noqa: E402

CFG = os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py')
GROUPS = {'roi_align': ('roi_align_fwd',),
          'nms': ('nms_sorted', 'nms_segmented', 'soft_nms_linear', 'rpn_decode'),
          'mask_paste': ('paste_masks',), 'offset_decode': ('offset_fusion_decode',),
          'dense': ('gemm_fprop', 'conv3x3_fprop', 'conv3x3_fprop_grouped')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, nargs='+', default=[1, 2, 4, 8, 16, 32])
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--breakdown', action='store_true', help='per-op CUDA-event timing (adds syncs)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = Config.fromfile(CFG)
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.eval()
    S = args.size
    metas = [dict(img_shape=(S, S, 3), ori_shape=(S, S, 3), pad_shape=(S, S, 3), scale_factor=1.0,
                  flip=False)]
    imgs = [torch.randn(1, 3, S, S, device=dev) for _ in range(4)]
    for _ in range(3):
        model.simple_test(imgs[0], metas)
    torch.cuda.synchronize()

    # optional per-op timing: wrap the binding
    events = []
    orig_call = L.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        events.append((name, e0, e1))

    for B in args.batches:
        events.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ndet = 0
        for i in range(B):
            out = model.simple_test(imgs[i % len(imgs)], metas)
            ndet += out[0][0].shape[0]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line = {'metric': 'inference images/sec LOFT R50-FPN', 'batch': B, 'size': S,
                'value': round(B / dt, 2), 'unit': 'img/s', 'ms_per_img': round(dt / B * 1e3, 2),
                'dets_per_img': ndet / B,
                'note': 'includes the host-side result packing the reference API mandates '
                        '(numpy bbox/offset arrays, one bool mask per detection)'}
        if args.breakdown:
            L.call = timed_call
            events.clear()
            model.simple_test(imgs[0], metas)
            torch.cuda.synchronize()
            L.call = orig_call
            agg = {}
            for name, e0, e1 in events:
                agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
            line['breakdown_ms'] = {g: round(sum(agg.get(n, 0.0) for n in names), 3)
                                    for g, names in GROUPS.items()}
            line['breakdown_ms']['all_c_abi_calls'] = round(sum(agg.values()), 3)
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
