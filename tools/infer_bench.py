#!/usr/bin/env python
"""Inference throughput sweep (BASELINE.json configs[4]): LOFT R50-FPN on synthetic 1024x1024
tiles, batch 1 -> 32, with the RoIAlign / NMS / mask-paste latency breakdown measured by CUDA
events around the C-ABI calls.

Two forms are timed:
  device  `model.simple_test_batch(imgs[B], metas)`: the dense work batched over the B tiles,
          results left on the device (dets, labels, bool masks, offsets) and the masks
          run-length encoded on the device (core.encode_mask_results) -- what a B200 serving
          path would do;
  api     `model.simple_test(img, metas)` tile by tile with the reference's result packing
          (detectors/base.py:141-143 asserts batch 1; one host bool bitmap per detection).

    python tools/infer_bench.py [--batches 1 2 4 8 16 32] [--size 1024] [--max-dets 100]
Prints one JSON line per batch size.  --max-dets caps the detections kept per tile: a random-init
model passes all 2000 (test_cfg.rcnn.max_per_img) through the mask and FOA heads, a trained one
keeps about as many as there are buildings (BONAI mean: 81.5 per tile).
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
from bonai_b200.core import encode_mask_results  # noqa: E402
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
    ap.add_argument('--max-dets', type=int, default=100)
    ap.add_argument('--api', action='store_true', help='also time the reference-format path')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = Config.fromfile(CFG)
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.eval()
    S = args.size
    meta = dict(img_shape=(S, S, 3), ori_shape=(S, S, 3), pad_shape=(S, S, 3), scale_factor=1.0,
                flip=False)
    events = []
    orig_call = L.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        events.append((name, e0, e1))

    def run(imgs, metas):
        res = model.simple_test_batch(imgs, metas, max_dets=args.max_dets)
        rle = [encode_mask_results(r[2]) for r in res]
        return res, rle

    for B in args.batches:
        imgs = torch.randn(B, 3, S, S, device=dev)
        metas = [meta] * B
        for _ in range(2):
            run(imgs, metas)
        torch.cuda.synchronize()
        reps = max(1, 8 // B)
        t0 = time.perf_counter()
        for _ in range(reps):
            res, rle = run(imgs, metas)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        line = {'metric': 'inference images/sec LOFT R50-FPN', 'batch': B, 'size': S,
                'value': round(B / dt, 2), 'unit': 'img/s', 'ms_per_img': round(dt / B * 1e3, 2),
                'dets_per_img': sum(r[0].shape[0] for r in res) / B, 'max_dets': args.max_dets,
                'form': 'device: batched dense work, device-resident results, device RLE'}
        # per-op CUDA-event breakdown of one batch (adds event overhead, not part of `value`)
        L.call = timed_call
        events.clear()
        run(imgs, metas)
        torch.cuda.synchronize()
        L.call = orig_call
        agg = {}
        for name, e0, e1 in events:
            agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
        line['breakdown_ms_per_img'] = {g: round(sum(agg.get(n, 0.0) for n in names) / B, 3)
                                        for g, names in GROUPS.items()}
        line['breakdown_ms_per_img']['all_c_abi_calls'] = round(sum(agg.values()) / B, 3)
        cnt = {}
        for name, _, _ in events:
            cnt[name] = cnt.get(name, 0) + 1
        line['top_calls_ms_per_img'] = {n: [round(v / B, 3), cnt[n]] for n, v in
                                        sorted(agg.items(), key=lambda kv: -kv[1])[:8]}
        if args.api:
            for _ in range(2):
                model.simple_test(imgs[:1], [meta])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(B):
                model.simple_test(imgs[i:i + 1], [meta])
            torch.cuda.synchronize()
            line['api_ms_per_img'] = round((time.perf_counter() - t0) / B * 1e3, 2)
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
