"""GPU idle gaps of one training step (torch.profiler / CUPTI): for every gap between consecutive
kernels longer than a threshold, the kernels on either side -- where the launch thread starves the GPU."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bonai_b200 import Config
from bonai_b200.apis import Trainer
from bonai_b200.models import build_detector
from torch.profiler import profile, ProfilerActivity

cfg = Config.fromfile(bench.CFG)
torch.manual_seed(0)
model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
model.train()
dev = torch.device('cuda:0')
trainer = Trainer(model, cfg, dev)
data = bench.to_model_inputs(bench.make_batch(0, device=dev))
for _ in range(6):
    trainer.train_step(data, prefetch=data)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        trainer.train_step(data, prefetch=data)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev], key=lambda t: t[0])
# steps are delimited by sgd_kernel
idx = [i for i, k in enumerate(ks) if 'sgd_kernel' in k[2]]
a, b = idx[-2] + 1, idx[-1] + 1
step = ks[a:b]
t0, t1 = step[0][0], step[-1][1]
busy = 0.0
gaps = []
end = step[0][0]
for s, e, n in step:
    if s > end:
        gaps.append((s - end, prev, n, end - t0))
    busy += max(0.0, e - max(s, end))
    if e > end:
        end, prev = e, n
print(f'step span {(t1 - t0) / 1e3:.3f} ms, kernels {len(step)}, GPU busy {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms')
thr = 15.0
big = [g for g in gaps if g[0] >= thr]
print(f'gaps >= {thr} us: {len(big)} totalling {sum(g[0] for g in big) / 1e3:.3f} ms; smaller gaps: {len(gaps) - len(big)} totalling {sum(g[0] for g in gaps if g[0] < thr) / 1e3:.3f} ms')
for g in sorted(big, key=lambda g: -g[0])[:40]:
    print(f'{g[0]:8.1f} us at +{g[3] / 1e3:7.3f} ms  after {g[1][:60]:60s} before {g[2][:60]}')

# GPU time per kernel name inside the step (warm durations, unlike the serialised ncu launch list)
import collections as _c
_agg = _c.defaultdict(lambda: [0, 0.0])
for s_, e_, n_ in step:
    _agg[n_][0] += 1
    _agg[n_][1] += e_ - s_
print('--- kernels of the step by GPU time (count, ms)')
for n_, (c_, us_) in sorted(_agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{n_[:100]:100s} {c_:5d} {us_ / 1e3:8.3f}')

if os.environ.get('LOFT_TIMELINE_LIST'):
    # every kernel of the step: start (ms since step start), duration (us), gap before it (us)
    end = step[0][0]
    for s_, e_, n_ in step:
        print(f'{(s_ - t0) / 1e3:8.3f} {e_ - s_:8.1f} {max(0.0, s_ - end):7.1f}  {n_[:90]}')
        end = max(end, e_)

# host side: which ops the launch thread spends its time in (one step)
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in cpu:
    if e.time_range.start >= t0_cpu_hint if False else True:
        agg[e.name][0] += 1
        agg[e.name][1] += e.self_cpu_time_total
print('--- host ops over the 3 profiled steps (count, self ms)')
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{n[:70]:70s} {c:6d} {us / 1e3:8.3f}')
