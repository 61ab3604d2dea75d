"""List the host<->device synchronisation points of one training step (torch sync debug mode)."""
import os, sys, warnings, collections, traceback, torch
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
for _ in range(4):
    trainer.train_step(data)
torch.cuda.synchronize()
sites = collections.Counter()
def showwarning(message, category, filename, lineno, file=None, line=None):
    if 'synchroniz' not in str(message):
        return
    st = [f for f in traceback.extract_stack() if '/bonai_b200/' in f.filename or '/bench.py' in f.filename]
    key = ' <- '.join(f'{os.path.relpath(f.filename, ROOT)}:{f.lineno}' for f in reversed(st[-3:]))
    sites[key] += 1
warnings.showwarning = showwarning
warnings.simplefilter('always')
torch.cuda.set_sync_debug_mode('warn')
trainer.train_step(data)
torch.cuda.set_sync_debug_mode('default')
torch.cuda.synchronize()
print('synchronising calls in one step:', sum(sites.values()))
for k, c in sites.most_common():
    print(f'{c:3d}  {k}')
