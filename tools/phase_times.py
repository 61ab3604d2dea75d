"""Host (wall) vs GPU (CUDA events) time of one training step, split at the forward/backward
boundary, to see whether the launch thread or the GPU bounds the step."""
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
data = bench.to_model_inputs(bench.make_batch(0, device=dev))
for _ in range(5):
    trainer.train_step(data, prefetch=data)
torch.cuda.synchronize()

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))

def wrap(obj, meth, name):
    f = getattr(obj, meth)
    def g(*a, **k):
        mark(name + ':begin'); r = f(*a, **k); mark(name + ':end'); return r
    setattr(obj, meth, g)

wrap(model, 'forward_train', 'forward')
wrap(model.rpn_head, 'loss', 'rpn_loss')
wrap(model.rpn_head, 'get_bboxes', 'rpn_proposals')
wrap(model.roi_head, 'assign_and_sample', 'rcnn_sample')
wrap(model.roi_head, '_bbox_forward_train', 'bbox_branch')
wrap(model.roi_head, '_mask_forward_train', 'mask_branch')
wrap(model.roi_head, '_offset_forward_train', 'offset_branch')
wrap(trainer.store, 'sgd_step', 'sgd')
for it in range(3):
    marks.clear()
    torch.cuda.synchronize()
    mark('step:begin')
    trainer.train_step(data, prefetch=data)
    mark('step:end')
    torch.cuda.synchronize()
    d = {n: (e, t) for n, e, t in marks}
    names = []
    for n, _, _ in marks:
        b = n.split(':')[0]
        if b not in names: names.append(b)
    print(f'--- iteration {it}\n{"phase":22s} {"gpu_ms":>8s} {"host_ms":>8s}')
    for b in names:
        e0, t0 = d[b + ':begin']; e1, t1 = d[b + ':end']
        print(f'{b:22s} {e0.elapsed_time(e1):8.3f} {(t1 - t0) * 1e3:8.3f}')
    e0, t0 = d['forward:end']; e1, t1 = d['sgd:begin']
    print(f'{"loss-sum+backward":22s} {e0.elapsed_time(e1):8.3f} {(t1 - t0) * 1e3:8.3f}')
