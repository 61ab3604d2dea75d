"""Is the GPU waiting for the launch thread, and where?  Runs pipelined training steps (no
synchronisation between them) and, at marked points of the step, records the host time and a CUDA
event on the main stream.  lead = (time the GPU reaches the mark) - (time the host issued it): a
lead of ~0 means the GPU had nothing queued there (it idles until the host catches up); a large
lead means the host runs ahead.  No profiler attached."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bonai_b200 import Config
from bonai_b200.apis import Trainer
from bonai_b200.models import build_detector

cfg = Config.fromfile(bench.CFG)
torch.manual_seed(0)
model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
model.train()
dev = torch.device('cuda:0')
trainer = Trainer(model, cfg, dev)
batches = [bench.to_model_inputs(bench.make_batch(i, device=dev)) for i in range(4)]
for i in range(8):
    trainer.train_step(batches[i % 4], prefetch=batches[(i + 1) % 4])
torch.cuda.synchronize()

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e, time.perf_counter()))

def wrap(obj, meth, name):
    f = getattr(obj, meth)
    def g(*a, **k):
        mark(name + ':begin'); r = f(*a, **k); mark(name + ':end'); return r
    setattr(obj, meth, g)

wrap(model, 'forward_train', 'forward')
wrap(model.rpn_head, 'loss', 'rpn_loss')
wrap(model.roi_head, 'assign_and_sample', 'rcnn_sample')
wrap(model.roi_head, '_bbox_forward_train', 'bbox_branch')
wrap(model.roi_head, '_mask_forward_train', 'mask_branch')
wrap(model.roi_head, '_offset_forward_train', 'offset_branch')
wrap(trainer.store, 'sgd_step', 'sgd')
if hasattr(model, 'prefetch'):
    wrap(model, 'prefetch', 'prefetch')

torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True)
e0.record()
torch.cuda.synchronize()
t0 = time.perf_counter()
NS = 6
for i in range(NS):
    mark(f'step{i}:begin')
    trainer.train_step(batches[i % 4], prefetch=batches[(i + 1) % 4])
    mark(f'step{i}:end')
torch.cuda.synchronize()
print(f'{NS} pipelined steps: {(time.perf_counter() - t0) * 1e3 / NS:.3f} ms/step (wall)')
print(f'{"mark":24s} {"host_ms":>9s} {"gpu_ms":>9s} {"lead_ms":>8s}')
for n, e, t in marks:
    h = (t - t0) * 1e3
    g = e0.elapsed_time(e)
    print(f'{n:24s} {h:9.3f} {g:9.3f} {g - h:8.3f}')
