"""Training step driver: the B200-native counterpart of mmdet/apis/train.py:34-143 (+ the mmcv
pieces it wires together: EpochBasedRunner.train, OptimizerHook(grad_clip), StepLrUpdaterHook with
linear warm-up, MMDistributedDataParallel).

One process per GPU.  Gradients already live in one flat fp32 buffer (engine.ParamStore), so the
data-parallel exchange is a bucketed in-place NCCL all-reduce over views of that buffer issued on
a side stream, followed by ONE fused clip + SGD launch that also applies the 1/world scaling.
Log scalars are reduced as one packed vector and read back only when asked."""
import ctypes
import os
import random
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib as L
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


def launcher_env(launcher):
    """Translate a `--launcher slurm|mpi` environment into the RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_* variables `init_dist` reads (mmcv.runner.init_dist as called from tools/train.py:94-98:
    `_init_dist_slurm` derives them from SLURM_PROCID / SLURM_NTASKS / SLURM_NODELIST,
    `_init_dist_mpi` from OMPI_COMM_WORLD_*).  'pytorch' / 'none' leave the environment alone."""
    e = os.environ
    if launcher == 'slurm':
        e['RANK'] = e['SLURM_PROCID']
        e['WORLD_SIZE'] = e['SLURM_NTASKS']
        e['LOCAL_RANK'] = e.get('SLURM_LOCALID', str(int(e['SLURM_PROCID']) %
                                                     max(torch.cuda.device_count(), 1)))
        if 'MASTER_ADDR' not in e:
            nodes = e.get('SLURM_STEP_NODELIST', e.get('SLURM_NODELIST', '127.0.0.1'))
            # first host of a "prefix[a-b,c],other" list
            head = nodes.split(',')[0]
            if '[' in head:
                pre, rng = head.split('[', 1)
                head = pre + rng.rstrip(']').split(',')[0].split('-')[0]
            e['MASTER_ADDR'] = head
        e.setdefault('MASTER_PORT', '29500')
    elif launcher == 'mpi':
        e['RANK'] = e['OMPI_COMM_WORLD_RANK']
        e['WORLD_SIZE'] = e['OMPI_COMM_WORLD_SIZE']
        e['LOCAL_RANK'] = e.get('OMPI_COMM_WORLD_LOCAL_RANK', '0')
        e.setdefault('MASTER_ADDR', '127.0.0.1')
        e.setdefault('MASTER_PORT', '29500')
    elif launcher not in ('pytorch', 'none', None):
        raise ValueError(f'Invalid launcher type: {launcher}')


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
    """`iters_per_epoch`: length of one epoch in iterations (len(data_loader) of the reference's
    EpochBasedRunner); with it `epoch` advances by itself and the step LR policy / per-epoch
    checkpoints behave as in the reference.  Without it the caller drives `set_epoch`."""

    def __init__(self, model, cfg=None, device=None, lr=0.005, momentum=0.9, weight_decay=1e-4,
                 max_norm=35.0, bucket_bytes=64 << 20, iters_per_epoch=None):
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
        self.iters_per_epoch = iters_per_epoch
        self.max_epochs = cfg.get('total_epochs') if cfg is not None else None
        self._last_logs = None
        # Overlap of the gradient exchange with the backward (what the reference gets from DDP's
        # bucketed reducer, mmdet/apis/train.py:75-79): the RoI heads' parameters -- 64 % of the
        # gradient bytes, the tail of the flat buffer -- are all-reduced on NCCL's stream as soon
        # as autograd reaches the trunk's backward; only the trunk's share is exchanged after it.
        self.overlap = self.distributed and os.environ.get('LOFT_OVERLAP_COMM', '1') != '0'
        self._head_works = []
        self._upper_works = []
        if self.overlap:
            self.store.heads_done_hooks.append(self._exchange_heads)
            # ... and, opt-in (LOFT_SPLIT_COMM=1), layer3 / layer4 / FPN / RPN (94 % of the trunk's
            # bytes) under the backward of layer2, between the trunk's two backward programs.
            # Measured neutral (N=8: 936.1 vs 936.9 img/s, N=2: +1 %, gpurun_out/n{2,8}_split*):
            # what stays exposed after the heads' overlap is the fixed latency of one last
            # collective (0.45 ms at N=8) and the wait for the slowest rank, not bytes.
            if self.store.mid_start is not None and \
                    os.environ.get('LOFT_SPLIT_COMM', '0') != '0':
                self.store.upper_done_hooks.append(self._exchange_upper)
                # bn_finalize rewrites gradient rows in place: not while NCCL is reducing them
                self.store.pre_finalize.append(self._wait_upper)
            # the collective runs beside the trunk backward's persistent GEMM grids: cap its CTAs
            # and leave it as many SMs (read by NCCL when the communicator is created, i.e. at
            # the first collective below)
            self._sm_reserve = int(os.environ.get('LOFT_COMM_SMS', '8'))
            if os.environ.get('LOFT_COMM_CAP_CTAS', '0') != '0':
                os.environ.setdefault('NCCL_MAX_CTAS', str(self._sm_reserve))
        if self.distributed:
            self.sync_replicas()

    # ------------------------------------------------------------------ replicas / epochs
    def sync_replicas(self, src=0):
        """Make every rank's master weights, momentum and BN statistics those of rank `src` --
        what MMDistributedDataParallel does at construction (mmdet/apis/train.py:75-79): without it
        replicas seeded differently (or a checkpoint loaded on one rank) diverge silently."""
        if not self.distributed:
            return
        st = self.store
        dist.broadcast(st.P, src)
        dist.broadcast(st.M, src)
        for g in st._bn_groups:
            if g is not None:
                dist.broadcast(g['mean'], src)
                dist.broadcast(g['var'], src)
        st.refresh_weights(force=True)

    def _exchange_heads(self):
        st = self.store
        trunk = getattr(self.model, '_trunk', None)
        if trunk is not None and not trunk.bwd_sm_reserve:
            trunk.bwd_sm_reserve = self._sm_reserve
        if st.head_start < st.n_train:
            self._head_works = allreduce_flat(st.G[st.head_start:st.n_train],
                                              bucket_bytes=self.bucket_bytes, async_op=True)

    def _exchange_upper(self):
        st = self.store
        if self._head_works:            # same step: the heads' exchange is already in flight
            self._upper_works = allreduce_flat(st.G[st.mid_start:st.head_start],
                                               bucket_bytes=self.bucket_bytes, async_op=True)
            self._upper_exchanged = True

    def _wait_upper(self):
        for w in self._upper_works:
            w.wait()
        self._upper_works = []

    def _exchange_rest(self):
        """After backward: the part of the flat gradient not yet exchanged, then wait for all."""
        st = self.store
        if self._head_works:
            end = st.mid_start if getattr(self, '_upper_exchanged', False) else st.head_start
            self._upper_exchanged = False
            works = allreduce_flat(st.G[:end], bucket_bytes=self.bucket_bytes,
                                   async_op=True) if end > 0 else []
            for w in self._head_works + self._upper_works + works:
                w.wait()
            self._head_works, self._upper_works = [], []
        else:
            allreduce_flat(st.G, bucket_bytes=self.bucket_bytes)

    def reserve_memory(self, gigabytes):
        """Grow the caching allocator's pool once, up front: one block of `gigabytes` is allocated
        and released to the cache, from which the per-step temporaries (whose sizes follow the
        number of positives) are then carved without a cudaMalloc in the middle of a step."""
        if gigabytes and gigabytes > 0:
            free, _ = torch.cuda.mem_get_info(self.store.device)
            n = int(min(gigabytes * (1 << 30), free * 0.5))
            blk = torch.empty(n, dtype=torch.uint8, device=self.store.device)
            del blk

    def _gc_policy(self):
        """Keep Python's cyclic collector off the launch thread's critical path.  A generation-2
        sweep over the heap of a built model (hundreds of thousands of objects: modules, recorded
        launch programs, ctypes argument tuples) takes 10-30 ms, and when it fires in the middle of
        a step every other rank waits for this one in the gradient exchange.  After the programs
        are recorded the existing heap is frozen out of the collector's reach, automatic
        collection is switched off, and the young generations are collected explicitly every 16
        steps right here -- after the optimizer launch, when the launch thread is ~5 ms ahead of
        the GPU.  LOFT_GC=auto restores Python's default behaviour."""
        import gc
        if os.environ.get('LOFT_GC', 'managed') != 'managed':
            return
        if self.iter == 3:
            gc.collect()
            gc.freeze()
            gc.disable()
        elif self.iter > 3 and self.iter % 16 == 0:
            gc.collect(1)

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def end_epoch(self):
        """Advance the epoch counter (EpochBasedRunner.train: `self._epoch += 1`)."""
        self.epoch += 1
        return self.epoch

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

    def stage(self, host_batch, slots=3, mask_windows=False):
        """Copy a (pinned) host batch to the device on a dedicated copy stream, like a prefetching
        data loader would: the copies overlap the previous step's backward and the returned batch
        carries the event the compute streams wait on.  Accepts the reference's input dict
        (img, gt_bboxes, gt_labels, gt_masks, gt_offsets; lists of tensors or BitmapMasks).

        The device side is a ring of `slots` persistent byte buffers (grown on demand, then
        stable): the tensors of a staged batch are views into one of them, so the steady state
        makes no allocator calls at all (fresh `Tensor.to()` allocations on the copy stream with
        `record_stream` kept the caching allocator creating segments -- `cudaMalloc`, 10-100 ms on
        the launch thread -- whenever the rotation's tensor sizes changed).  A slot is reused only
        after the step that consumed it has finished (event recorded by `train_step`); a staged
        batch that was never passed to `train_step` keeps its buffer and the slot gets a new one.

        `mask_windows=True`: transfer only the gt-box window (+2 px) of every GT bitmap into a
        zeroed device stack (`loft_h2d_mask_windows`, one strided DMA per box, no host copy).  A
        footprint bitmap vanishes outside its box -- BONAI's boxes are the polygons' extents -- so
        the device tensors are identical while a 1024^2 tile with 80 buildings moves ~0.5 MB of
        masks instead of 80 MB (eight ranks staging 193 MB each per step saturate the host:
        20 ms per step at N=8).  Only valid for masks with that property; off by default."""
        from ..core import BitmapMasks
        dev = self.store.device
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
        ring = self.__dict__.setdefault('_stage_ring', None)
        if ring is None or len(ring['bufs']) != slots:
            ring = self._stage_ring = dict(bufs=[None] * slots, done=[None] * slots,
                                           pending=[False] * slots, i=0)
        k = ring['i']
        ring['i'] = (k + 1) % slots
        # layout: every tensor at a 256-byte aligned offset of the slot's buffer
        plan, off = [], 0

        windows = {}                         # plan index of a bitmap stack -> (host boxes, bytes)
        wcache = self.__dict__.setdefault('_window_bytes', {})

        def add(t, boxes=None):
            nonlocal off
            src = t
            if isinstance(t, BitmapMasks):
                src = t._t if t._t is not None else torch.from_numpy(t._np)
            if not isinstance(src, torch.Tensor):
                return None
            nb = src.numel() * src.element_size()
            plan.append((src, off, nb))
            if mask_windows and isinstance(t, BitmapMasks) and boxes is not None and \
                    src.dtype == torch.uint8 and src.dim() == 3 and src.is_contiguous() and \
                    src.numel() > 0 and src.is_pinned() and not boxes.is_cuda and \
                    boxes.dtype == torch.float32 and boxes.shape == (src.shape[0], 4):
                key = (src.data_ptr(), boxes.data_ptr(), src._version, boxes._version)
                if key not in wcache:           # bytes actually moved (for the caller's accounting)
                    if len(wcache) > 256:
                        wcache.clear()
                    G, H, W = src.shape
                    x0 = (boxes[:, 0].floor() - 2).clamp(0, W)
                    y0 = (boxes[:, 1].floor() - 2).clamp(0, H)
                    x1 = (boxes[:, 2].ceil() + 2).clamp(0, W)
                    y1 = (boxes[:, 3].ceil() + 2).clamp(0, H)
                    wcache[key] = int(((y1 - y0).clamp(min=0) * (x1 - x0).clamp(min=0)).sum())
                windows[len(plan) - 1] = (boxes, wcache[key])
            off = (off + nb + 255) // 256 * 256
            return len(plan) - 1

        index = {}
        for key, v in host_batch.items():
            if key == 'img_metas':
                continue
            if key == 'gt_masks' and isinstance(v, (list, tuple)):
                gb = host_batch.get('gt_bboxes') or [None] * len(v)
                index[key] = [add(t, b) for t, b in zip(v, gb)]
            else:
                index[key] = [add(t) for t in v] if isinstance(v, (list, tuple)) else add(v)
        nbytes = sum(windows[i][1] if i in windows else p[2] for i, p in enumerate(plan))
        if ring['pending'][k]:              # staged but never consumed: leave its memory alone
            ring['bufs'][k], ring['done'][k] = None, None
        buf = ring['bufs'][k]
        if buf is None or buf.numel() < off:
            buf = ring['bufs'][k] = torch.empty(int(off * 1.25) + 256, dtype=torch.uint8, device=dev)
            ring['done'][k] = None
        views = []
        with torch.cuda.stream(self._copy_stream):
            if ring['done'][k] is not None:
                self._copy_stream.wait_event(ring['done'][k])
            by_src = {}
            for i, (src, o, nb) in enumerate(plan):
                dst = buf[o:o + nb].view(src.dtype).view(src.shape)
                if i not in windows:
                    dst.copy_(src, non_blocking=True)
                by_src[src.data_ptr()] = dst
                views.append(dst)
            for i, (boxes, _) in windows.items():       # after the boxes' own copies (same stream)
                src, dst = plan[i][0], views[i]
                if boxes.data_ptr() not in by_src:          # boxes not part of this batch
                    dst.copy_(src, non_blocking=True)
                    continue
                dst.zero_()
                L.call('h2d_mask_windows', ctypes.c_void_p(src.data_ptr()),
                       L.ptr(by_src[boxes.data_ptr()]), L.ptr(dst), ctypes.c_int(src.shape[0]),
                       ctypes.c_int(src.shape[1]), ctypes.c_int(src.shape[2]), ctypes.c_int(2),
                       L.stream())
            ready = self._copy_stream.record_event()
        out = {}
        for key, v in host_batch.items():
            if key == 'img_metas':
                out[key] = v
                continue
            ix = index[key]

            def get(i, t):
                if i is None:
                    return t
                return BitmapMasks(views[i], t.height, t.width) if isinstance(t, BitmapMasks) \
                    else views[i]
            out[key] = [get(i, t) for i, t in zip(ix, v)] if isinstance(ix, list) else get(ix, v)
        out['ready_event'] = ready
        out['_stage_slot'] = k
        ring['pending'][k] = True
        self.staged_bytes = nbytes
        return out

    def train_step(self, data, read_logs=False, prefetch=None):
        """forward + backward + gradient all-reduce + clip + SGD.  Returns the device-resident
        packed log vector (and its key order); nothing synchronises the host unless read_logs.
        `prefetch`: the NEXT batch (device-resident or from `stage`) -- its input-only work (RPN
        targets) is issued right after this step's optimizer launch, when the launch thread would
        otherwise wait for the GPU to finish the backward."""
        model = self.model
        slot = None
        if '_stage_slot' in data:                    # a batch from `stage`: its ring slot
            data = dict(data)
            slot = data.pop('_stage_slot')
        at_fwd = prefetch is not None and hasattr(model, 'prefetch') and \
            os.environ.get('LOFT_PREFETCH', '1') != '0' and \
            os.environ.get('LOFT_PREFETCH_AT', 'step_end') == 'forward'
        losses = model(**data, prefetch_next=prefetch) if at_fwd else model(**data)

        def _m(v):                                   # fused losses are already 1-element sums
            return v.reshape(()) if v.numel() == 1 else v.mean()

        # Back-propagate FIRST, straight from the loss terms: the total is their plain sum
        # (detectors/base.py:175-208), so every term's upstream gradient is the same constant 1 --
        # one shared device scalar instead of a ones_like fill per root, and none of the ~25 tiny
        # mean / add / stack kernels of the logging sum sits between the end of the forward and
        # the first backward launch (measured: ~1.2 ms of launch-bound GPU idle there).
        terms = []
        for name, value in losses.items():
            if 'loss' in name:
                for v in (value if isinstance(value, (list, tuple)) else [value]):
                    if v.requires_grad:
                        terms.append(_m(v))
        if terms:
            if getattr(self, '_one', None) is None or self._one.device != terms[0].device:
                self._one = torch.ones((), device=terms[0].device)
            torch.autograd.backward(terms, [self._one] * len(terms))
        if self.distributed:
            if os.environ.get('LOFT_TIME_COMM'):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self._exchange_rest()
                e1.record()
                self.__dict__.setdefault('_comm_events', []).append((e0, e1))
            else:
                self._exchange_rest()
        self.store.sgd_step(self.current_lr(), self.momentum, self.weight_decay, self.max_norm,
                            grad_scale=1.0 / self.world)
        if slot is not None and getattr(self, '_stage_ring', None) is not None and \
                slot < len(self._stage_ring['done']):
            # every reader of the staged tensors is ordered before this point of the main stream
            self._stage_ring['done'][slot] = torch.cuda.current_stream().record_event()
            self._stage_ring['pending'][slot] = False
        self.iter += 1
        self._gc_policy()
        if self.iters_per_epoch and self.iter % self.iters_per_epoch == 0:
            self.end_epoch()
        if prefetch is not None and not at_fwd and hasattr(model, 'prefetch') and \
                os.environ.get('LOFT_PREFETCH', '1') != '0':
            model.prefetch(prefetch)
        # the log scalars are summed last, while the GPU is still busy with the backward
        log_vars = OrderedDict()
        with torch.no_grad():
            for name, value in losses.items():
                if isinstance(value, torch.Tensor):
                    log_vars[name] = _m(value.detach())
                else:
                    log_vars[name] = sum(_m(v.detach()) for v in value)
            log_vars['loss'] = sum(v for k, v in log_vars.items() if 'loss' in k)
            packed = torch.stack([v.reshape(()) for v in log_vars.values()])
        self._last_logs = (list(log_vars.keys()), packed)
        if read_logs == 'async':
            return self.read_logs_async()
        if read_logs:
            return self.read_logs()
        return packed

    # ------------------------------------------------------------------ checkpoints
    def _sgd_state_dict(self):
        """torch.optim.SGD.state_dict() layout, which is what mmcv's save_checkpoint stores under
        'optimizer' for the reference: one param group over ALL of model.parameters() in order
        (mmcv DefaultOptimizerConstructor without paramwise_cfg passes `model.parameters()`, frozen
        ones included; they simply never acquire state)."""
        mom = self.store.momentum_state()
        state, idx = {}, []
        for i, (name, p) in enumerate(self.model.named_parameters()):
            idx.append(i)
            if name in mom and self.iter > 0:
                state[i] = {'momentum_buffer': mom[name].cpu()}
        group = dict(lr=self.current_lr(), momentum=self.momentum, dampening=0,
                     weight_decay=self.weight_decay, nesterov=False, initial_lr=self.base_lr,
                     params=idx)
        return {'state': state, 'param_groups': [group]}

    def _load_sgd_state_dict(self, opt):
        import warnings
        names = [n for n, _ in self.model.named_parameters()]
        if 'state' in opt and 'param_groups' in opt:
            order = [i for g in opt['param_groups'] for i in g['params']]
            if len(order) != len(names):
                warnings.warn(f'optimizer state covers {len(order)} parameters, the model has '
                              f'{len(names)}: momentum not restored')
                return False
            mom = {names[k]: opt['state'][i]['momentum_buffer']
                   for k, i in enumerate(order) if i in opt['state'] and
                   opt['state'][i].get('momentum_buffer') is not None}
        elif 'momentum_buffer' in opt:          # round-1 layout of this repo
            mom = opt['momentum_buffer']
        else:
            warnings.warn('unrecognised optimizer state in checkpoint: momentum not restored')
            return False
        self.store.load_momentum_state(mom)
        return True

    def save_checkpoint(self, path, meta=None):
        """Reference checkpoint layout (mmcv save_checkpoint as used by CheckpointHook,
        default_runtime.py:1): {'meta', 'state_dict', 'optimizer'}; state_dict keys / shapes are
        the reference's, tensors are saved contiguous in the reference (OIHW) order; 'optimizer'
        is torch.optim.SGD.state_dict() so the reference's runner.resume accepts the file."""
        sd = {k: v.detach().cpu().contiguous() for k, v in self.model.state_dict().items()}
        m = dict(iter=self.iter, epoch=self.epoch, lr=self.current_lr())
        m.update(meta or {})
        torch.save({'meta': m, 'state_dict': sd, 'optimizer': self._sgd_state_dict()}, path)

    def load_checkpoint(self, path, resume=True):
        ck = torch.load(path, map_location='cpu')
        sd = ck['state_dict'] if 'state_dict' in ck else ck
        # checkpoints written from a (MM)DistributedDataParallel wrapper carry a 'module.' prefix
        # (mmcv load_checkpoint strips it)
        if sd and all(k.startswith('module.') for k in sd):
            sd = {k[len('module.'):]: v for k, v in sd.items()}
        self.model.load_state_dict(sd)
        self.store.refresh_weights(force=True)
        if resume:
            if 'optimizer' in ck:
                self._load_sgd_state_dict(ck['optimizer'])
            self.iter = ck.get('meta', {}).get('iter', 0)
            self.epoch = ck.get('meta', {}).get('epoch', 0)
        if self.distributed:
            self.sync_replicas()
        return ck.get('meta', {})

    # ------------------------------------------------------------------ epoch-based run
    def run(self, batches, max_epochs=None, work_dir=None, checkpoint_interval=1, log_interval=50,
            rank=0, log=print):
        """EpochBasedRunner.run for workflow [('train', 1)] (mmdet/apis/train.py:143, mmcv
        EpochBasedRunner.train): `iters_per_epoch` steps per epoch, LR stepped by epoch, a
        checkpoint `epoch_{n}.pth` every `checkpoint_interval` epochs (CheckpointHook,
        default_runtime.py:1).  `batches`: iterator of host input dicts."""
        import time
        assert self.iters_per_epoch, 'Trainer.run needs iters_per_epoch'
        max_epochs = max_epochs or self.max_epochs or 1
        nxt = self.stage(next(batches))
        while self.epoch < max_epochs:
            ep, t0 = self.epoch, time.time()
            for i in range(self.iter % self.iters_per_epoch, self.iters_per_epoch):
                last = (i + 1 == self.iters_per_epoch) and (ep + 1 == max_epochs)
                cur, nxt = nxt, (None if last else self.stage(next(batches)))
                want = (i + 1) % log_interval == 0 or i + 1 == self.iters_per_epoch
                lr = self.current_lr()
                out = self.train_step(cur, read_logs=want, prefetch=nxt)
                if want and rank == 0:
                    items = ', '.join(f'{k}: {v:.4f}' for k, v in out.items())
                    log(f'Epoch [{ep + 1}][{i + 1}/{self.iters_per_epoch}]\tlr: {lr:.3e}, '
                        f'time: {(time.time() - t0) / (i + 1):.3f}, {items}')
            if work_dir and rank == 0 and checkpoint_interval and \
                    self.epoch % checkpoint_interval == 0:
                self.save_checkpoint(os.path.join(work_dir, f'epoch_{self.epoch}.pth'))

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
