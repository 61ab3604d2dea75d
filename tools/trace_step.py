"""Per-call GPU time of one training step (CUDA events around every C-ABI call, shapes attached):
which launches of which shapes the step time goes to.  Event-bracketed times include the launch
gap to the previous kernel, so they are upper bounds on the kernel's own duration."""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bonai_b200 import Config, _lib as L
from bonai_b200.apis import Trainer
from bonai_b200.models import build_detector

if len(sys.argv) > 1:
    bench.NUM_GT = int(sys.argv[1])
cfg = Config.fromfile(bench.CFG)
torch.manual_seed(1234)
model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
model.train()
dev = torch.device('cuda:0')
trainer = Trainer(model, cfg, dev)
data = bench.to_model_inputs(bench.make_batch(0, device=dev))
for _ in range(3):
    trainer.train_step(data, prefetch=data)
torch.cuda.synchronize()
L.TRACE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
trainer.train_step(data)
e1.record()
torch.cuda.synchronize()
tr, L.TRACE = L.TRACE, None
rows = [(n, a, s.elapsed_time(e) * 1e3) for n, a, s, e in tr]
tot = sum(r[2] for r in rows)
print(f'step {e0.elapsed_time(e1):.3f} ms; {len(rows)} C-ABI calls, summed {tot / 1e3:.3f} ms')
agg = collections.OrderedDict()
for n, a, us in rows:
    k = (n, a)
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += us
print(f'{"call":28s} {"n":>3s} {"us/call":>8s} {"ms":>7s}  args')
for (n, a), (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:28s} {c:3d} {us / c:8.1f} {us / 1e3:7.3f}  {a}')
byname = collections.Counter()
for n, a, us in rows:
    byname[n] += us
print('--- by entry point')
for n, us in byname.most_common():
    print(f'{n:28s} {us / 1e3:7.3f} ms')
