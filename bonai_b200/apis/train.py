"""Training step driver: the B200-native counterpart of mmdet/apis/train.py:34-143 (+ the mmcv
pieces it wires together: EpochBasedRunner.train, OptimizerHook(grad_clip), StepLrUpdaterHook with
linear warm-up, MMDistributedDataParallel).

One process per GPU.  Gradients already live in one flat fp32 buffer (engine.ParamStore), so the
data-parallel exchange is a bucketed in-place NCCL all-reduce over views of that buffer issued on
a side stream, followed by ONE fused clip + SGD launch that also applies the 1/world scaling.
Log scalars are reduced as one packed vector and read back only when asked."""
import os
import random
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from ..engine import get_store


def set_random_seed(seed, deterministic=False):
    """mmdet/apis/train.py:15-31."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def init_dist(backend='nccl', **kwargs):
    """torch.distributed bootstrap from the torchrun environment (tools/train.py:94-98,
    default_runtime.py:10)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', rank))
    if backend == 'nccl':
        torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world


def step_lr(base_lr, it, epoch, step=(16, 22), gamma=0.1, warmup='linear', warmup_iters=300,
            warmup_ratio=0.001):
    """mmcv StepLrUpdaterHook + linear warm-up as configured by
    configs/_base_/schedules/schedule_2x_bonai.py:5-10."""
    exp = sum(1 for s in step if epoch >= s)
    lr = base_lr * gamma ** exp
    if warmup is not None and it < warmup_iters:
        if warmup == 'linear':
            k = (1 - it / warmup_iters) * (1 - warmup_ratio)
            lr = lr * (1 - k)
        elif warmup == 'constant':
            lr = lr * warmup_ratio
        else:
            raise ValueError(warmup)
    return lr


def build_optimizer_args(cfg):
    opt = dict(cfg.optimizer)
    if opt.pop('type') != 'SGD':
        raise NotImplementedError('LOFT path: SGD (schedule_2x_bonai.py:2)')
    clip = (cfg.get('optimizer_config') or {}).get('grad_clip')
    if clip is not None and clip.get('norm_type', 2) != 2:
        raise NotImplementedError('LOFT path: L2 gradient clipping')
    return dict(lr=opt['lr'], momentum=opt.get('momentum', 0.0),
                weight_decay=opt.get('weight_decay', 0.0),
                max_norm=clip['max_norm'] if clip else None)


def bucket_views(flat, bucket_bytes=64 << 20):
    """Contiguous views of a flat buffer, last parameters first (their gradients are ready first in
    backward), each about `bucket_bytes` large."""
    n = flat.numel()
    per = max(bucket_bytes // flat.element_size(), 1)
    out, end = [], n
    while end > 0:
        start = max(end - per, 0)
        out.append(flat[start:end])
        end = start
    return out


def allreduce_flat(flat, group=None, bucket_bytes=64 << 20, async_op=False):
    """Sum-all-reduce a flat gradient buffer in place, bucket by bucket."""
    works = []
    for v in bucket_views(flat, bucket_bytes):
        w = dist.all_reduce(v, group=group, async_op=True)
        works.append(w)
    if async_op:
        return works
    for w in works:
        w.wait()
    return []


class Trainer:
    def __init__(self, model, cfg=None, device=None, lr=0.005, momentum=0.9, weight_decay=1e-4,
                 max_norm=35.0, bucket_bytes=64 << 20):
        self.model = model
        if cfg is not None:
            a = build_optimizer_args(cfg)
            lr, momentum, weight_decay, max_norm = a['lr'], a['momentum'], a['weight_decay'], \
                a['max_norm']
            self.lr_cfg = dict(cfg.get('lr_config') or {})
        else:
            self.lr_cfg = {}
        self.base_lr, self.momentum, self.weight_decay, self.max_norm = lr, momentum, \
            weight_decay, max_norm
        self.bucket_bytes = bucket_bytes
        self.store = get_store(model, device)
        self.distributed = dist.is_available() and dist.is_initialized() and \
            dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.distributed else 1
        self.iter = 0
        self.epoch = 0
        self._last_logs = None

    def current_lr(self):
        c = self.lr_cfg
        if not c:
            return self.base_lr
        if c.get('policy', 'step') != 'step':
            raise NotImplementedError('LOFT path: step LR policy')
        return step_lr(self.base_lr, self.iter, self.epoch, step=c.get('step', (16, 22)),
                       gamma=c.get('gamma', 0.1), warmup=c.get('warmup'),
                       warmup_iters=c.get('warmup_iters', 0),
                       warmup_ratio=c.get('warmup_ratio', 0.1))

    def stage(self, host_batch):
        """Copy a (pinned) host batch to the device on a dedicated copy stream, like a prefetching
        data loader would: the copies overlap the previous step's backward and the returned batch
        carries the event the compute streams wait on.  Accepts the reference's input dict
        (img, gt_bboxes, gt_labels, gt_masks, gt_offsets; lists of tensors or BitmapMasks)."""
        from ..core import BitmapMasks
        dev = self.store.device
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
        out, nbytes = {}, 0

        def mv(t):
            nonlocal nbytes
            if isinstance(t, BitmapMasks):
                src = t._t if t._t is not None else torch.from_numpy(t._np)
                nbytes += src.numel() * src.element_size()
                return BitmapMasks(src.to(dev, non_blocking=True), t.height, t.width)
            if isinstance(t, torch.Tensor):
                nbytes += t.numel() * t.element_size()
                return t.to(dev, non_blocking=True)
            return t

        with torch.cuda.stream(self._copy_stream):
            for k, v in host_batch.items():
                out[k] = [mv(t) for t in v] if isinstance(v, (list, tuple)) and k != 'img_metas' \
                    else mv(v)
            out['ready_event'] = self._copy_stream.record_event()
        main = torch.cuda.current_stream(dev)
        for k, v in out.items():
            for t in (v if isinstance(v, list) else [v]):
                t = t._t if isinstance(t, BitmapMasks) else t
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)
        self.staged_bytes = nbytes
        return out

    def train_step(self, data, read_logs=False, prefetch=None):
        """forward + backward + gradient all-reduce + clip + SGD.  Returns the device-resident
        packed log vector (and its key order); nothing synchronises the host unless read_logs.
        `prefetch`: the NEXT batch (device-resident or from `stage`) -- its input-only work (RPN
        targets) is issued right after this step's optimizer launch, when the launch thread would
        otherwise wait for the GPU to finish the backward."""
        model = self.model
        losses = model(**data)
        log_vars = OrderedDict()
        def _m(v):                                   # fused losses are already 1-element sums
            return v.reshape(()) if v.numel() == 1 else v.mean()

        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = _m(value)
            else:
                log_vars[name] = sum(_m(v) for v in value)
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        loss.backward()
        if self.distributed:
            if os.environ.get('LOFT_TIME_COMM'):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                allreduce_flat(self.store.G, bucket_bytes=self.bucket_bytes)
                e1.record()
                self.__dict__.setdefault('_comm_events', []).append((e0, e1))
            else:
                allreduce_flat(self.store.G, bucket_bytes=self.bucket_bytes)
        self.store.sgd_step(self.current_lr(), self.momentum, self.weight_decay, self.max_norm,
                            grad_scale=1.0 / self.world)
        self.iter += 1
        if prefetch is not None and hasattr(model, 'prefetch') and \
                os.environ.get('LOFT_PREFETCH', '1') != '0':
            model.prefetch(prefetch)
        packed = torch.stack([v.detach().reshape(()) for v in log_vars.values()])
        self._last_logs = (list(log_vars.keys()), packed)
        if read_logs == 'async':
            return self.read_logs_async()
        if read_logs:
            return self.read_logs()
        return packed

    # ------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, path, meta=None):
        """Reference checkpoint layout (mmcv save_checkpoint as used by CheckpointHook,
        default_runtime.py:1): {'meta', 'state_dict', 'optimizer'}; state_dict keys / shapes are
        the reference's, tensors are saved contiguous in the reference (OIHW) order."""
        sd = {k: v.detach().cpu().contiguous() for k, v in self.model.state_dict().items()}
        opt = {k: v.cpu() for k, v in self.store.momentum_state().items()}
        m = dict(iter=self.iter, epoch=self.epoch, lr=self.current_lr())
        m.update(meta or {})
        torch.save({'meta': m, 'state_dict': sd, 'optimizer': {'momentum_buffer': opt}}, path)

    def load_checkpoint(self, path, resume=True):
        ck = torch.load(path, map_location='cpu')
        sd = ck['state_dict'] if 'state_dict' in ck else ck
        self.model.load_state_dict(sd)
        self.store.refresh_weights(force=True)
        if resume:
            if 'optimizer' in ck and 'momentum_buffer' in ck['optimizer']:
                self.store.load_momentum_state(ck['optimizer']['momentum_buffer'])
            self.iter = ck.get('meta', {}).get('iter', 0)
            self.epoch = ck.get('meta', {}).get('epoch', 0)
        return ck.get('meta', {})

    def read_logs_async(self):
        """Non-blocking variant: enqueue the device->host copy of THIS step's packed log vector
        into pinned memory and return the PREVIOUS step's values (None on the first call).  The
        launch thread never waits for the GPU; `flush_logs()` returns the last step's values."""
        keys, packed = self._last_logs
        if self.distributed:
            packed = packed / self.world
            dist.all_reduce(packed)
        prev = self.flush_logs() if getattr(self, '_pending', None) is not None else None
        if not hasattr(self, '_pin'):
            self._pin = [torch.empty(packed.numel(), dtype=torch.float32).pin_memory()
                         for _ in range(2)]
            self._pin_i = 0
        buf = self._pin[self._pin_i]
        self._pin_i ^= 1
        buf.copy_(packed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending = (keys, buf, ev)
        return prev

    def flush_logs(self):
        keys, buf, ev = self._pending
        ev.synchronize()
        self._pending = None
        return OrderedDict(zip(keys, buf.tolist()))

    def read_logs(self):
        """Packed mean over ranks + one host read-back (replaces the 8 all_reduce + .item() pairs
        of detectors/base.py:201-206)."""
        keys, packed = self._last_logs
        if self.distributed:
            packed = packed / self.world
            dist.all_reduce(packed)
        return OrderedDict(zip(keys, packed.tolist()))
