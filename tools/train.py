#!/usr/bin/env python
"""Training entry point mirroring the reference's `tools/train.py CONFIG [--launcher pytorch]
[--options k=v ...] [--work-dir D] [--resume-from CKPT]` (tools/train.py:25-156) on the B200 path.

The BONAI dataset is not shipped with the reference (README.md:21-23), so batches are synthetic
tiles with the reference's input contract unless --data-module names a Python module
exposing `iter_batches(cfg, rank, world)` that yields the reference's input dicts.
"""
import argparse
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import Config, DictAction  # noqa: E402
from bonai_b200.apis import Trainer, init_dist, launcher_env, set_random_seed  # noqa: E402
from bonai_b200.core import BitmapMasks  # noqa: E402
from bonai_b200.models import build_detector  # noqa: E402


def synthetic_batches(cfg, rank, world, size, num_gt, device):
    """Yield one fixed-recipe batch per iteration (samples_per_gpu tiles)."""
    import math
    n = cfg.get('data', {}).get('samples_per_gpu', 2) if 'data' in cfg else 2
    g = torch.Generator().manual_seed(1000 + rank)
    it = 0
    while True:
        img = torch.randn(n, 3, size, size, generator=g)
        gb, gl, gm, go = [], [], [], []
        yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32),
                                torch.arange(size, dtype=torch.float32), indexing='ij')
        for _ in range(n):
            c = torch.rand(num_gt, 2, generator=g) * size
            wh = torch.exp(torch.rand(num_gt, 2, generator=g) * math.log(10)) * 16
            b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, size)
            b[:, 2:] = torch.max(b[:, 2:], b[:, :2] + 2).clamp(max=size)
            b[:, :2] = torch.min(b[:, :2], b[:, 2:] - 2)
            cx, cy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
            rx, ry = (b[:, 2] - b[:, 0]) / 2, (b[:, 3] - b[:, 1]) / 2
            m = (((xx[None] + .5 - cx[:, None, None]) / rx[:, None, None]) ** 2 +
                 ((yy[None] + .5 - cy[:, None, None]) / ry[:, None, None]) ** 2) <= 1
            gb.append(b)
            gl.append(torch.zeros(num_gt, dtype=torch.long))
            gm.append(BitmapMasks(m.to(torch.uint8), size, size))
            go.append(torch.rand(num_gt, 2, generator=g) * 80 - 40)
        metas = [dict(img_shape=(size, size, 3), pad_shape=(size, size, 3),
                      ori_shape=(size, size, 3), scale_factor=1.0, flip=False)] * n
        yield dict(img=img, img_metas=metas, gt_bboxes=gb, gt_labels=gl, gt_masks=gm,
                   gt_offsets=go)
        it += 1


def main():
    ap = argparse.ArgumentParser(description='Train LOFT on the B200 path')
    ap.add_argument('config')
    ap.add_argument('--work-dir', default=None)
    ap.add_argument('--resume-from', default=None)
    ap.add_argument('--launcher', choices=['none', 'pytorch', 'slurm', 'mpi'], default='none')
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--options', nargs='+', action=DictAction)
    ap.add_argument('--iters', type=int, default=None,
                    help='iteration-based smoke run: stop after this many iterations')
    ap.add_argument('--iters-per-epoch', type=int, default=None,
                    help='epoch length (len(data_loader) in the reference); default: '
                         '3300 BONAI training tiles / global batch, or the data module\'s '
                         '`iters_per_epoch(cfg, world)`')
    ap.add_argument('--epochs', type=int, default=None, help='default: cfg.total_epochs')
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--num-gt', type=int, default=80)
    ap.add_argument('--data-module', default=None)
    ap.add_argument('--local_rank', type=int, default=0)
    args = ap.parse_args()

    cfg = Config.fromfile(args.config)
    if args.options:
        cfg.merge_from_dict(args.options)
    cfg.model.pretrained = None
    launcher_env(args.launcher)
    rank, world = (init_dist(cfg.get('dist_params', {}).get('backend', 'nccl'))
                   if args.launcher != 'none' else (0, 1))
    if args.seed is not None:
        set_random_seed(args.seed)
    torch.manual_seed(args.seed if args.seed is not None else 0)
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.train()
    data_mod = importlib.import_module(args.data_module) if args.data_module else None
    per_gpu = cfg.get('data', {}).get('samples_per_gpu', 2) if 'data' in cfg else 2
    ipe = args.iters_per_epoch
    if ipe is None and data_mod is not None and hasattr(data_mod, 'iters_per_epoch'):
        ipe = data_mod.iters_per_epoch(cfg, world)
    if ipe is None:
        ipe = max(1, 3300 // (per_gpu * world))      # BONAI: 3 300 training tiles (README.md:13)
    trainer = Trainer(model, cfg, device, iters_per_epoch=ipe)
    work_dir = args.work_dir or os.path.join('work_dirs', os.path.splitext(
        os.path.basename(args.config))[0])
    if rank == 0:
        os.makedirs(work_dir, exist_ok=True)
    resume = args.resume_from or cfg.get('resume_from')
    if resume:
        meta = trainer.load_checkpoint(resume, resume=True)
        if rank == 0:
            print(f'resumed from {resume} at epoch {trainer.epoch}, iter {meta.get("iter", 0)}')
    elif cfg.get('load_from'):
        trainer.load_checkpoint(cfg.load_from, resume=False)
    if data_mod is not None:
        batches = data_mod.iter_batches(cfg, rank, world)
    else:
        batches = synthetic_batches(cfg, rank, world, args.size, args.num_gt, device)
    interval = cfg.get('log_config', {}).get('interval', 10)
    if args.iters is None:
        # the reference's schedule: total_epochs epochs, LR stepped by epoch, epoch_{n}.pth
        ck = cfg.get('checkpoint_config', {}).get('interval', 1)
        trainer.run(batches, max_epochs=args.epochs, work_dir=work_dir, checkpoint_interval=ck,
                    log_interval=interval, rank=rank, log=lambda m: print(m, flush=True))
        return
    nxt = trainer.stage(next(batches))
    t0, n_img = time.time(), 0
    for it in range(args.iters):
        cur, nxt = nxt, trainer.stage(next(batches))
        log = (it + 1) % interval == 0 or it + 1 == args.iters
        out = trainer.train_step(cur, read_logs=log, prefetch=nxt)
        n_img += len(cur['img_metas']) * world
        if log and rank == 0:
            dt = time.time() - t0
            items = ', '.join(f'{k}: {v:.4f}' for k, v in out.items())
            print(f'Epoch [{trainer.epoch + 1}] Iter [{trainer.iter}]\tlr: {trainer.current_lr():.3e}, '
                  f'{n_img / dt:.1f} img/s, {items}', flush=True)
    if rank == 0:
        path = os.path.join(work_dir, f'iter_{trainer.iter}.pth')
        trainer.save_checkpoint(path, meta=dict(config=cfg.filename))
        print('saved', path)


if __name__ == '__main__':
    main()
