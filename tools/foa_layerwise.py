"""Layer-by-layer comparison of the GPU FOA conv stack with the TF32-emulated and the fp32 CPU
restatements (oracle/tf32_emu.py) on identical TF32-grid inputs: where do they diverge, and by
how much?  Prints, per conv layer of branch 0: relative L2 GPU-vs-emulated, GPU-vs-fp32, the
fraction of elements that differ at all, and the fraction whose ReLU mask differs."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import Config  # noqa: E402
from bonai_b200.engine import get_store  # noqa: E402
from bonai_b200.models import build_detector  # noqa: E402
from bonai_b200.ops import dense as D  # noqa: E402
from oracle import loft_cpu as O  # noqa: E402
from oracle import tf32_emu as E  # noqa: E402

cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
p = O.randomize_bn(O.init_params(0), 0)
model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
model.load_state_dict(p)
model.train()
store = get_store(model, torch.device('cuda:0'))
store.refresh_weights(force=True)
head = model.roi_head.offset_head
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(3)
x = E.rna(torch.relu(torch.randn(P, 256, 7, 7, generator=g)) * 0.7)
pre = 'roi_head.offset_head'

# GPU: the product's own grouped launches, layer by layer
with torch.no_grad():
    xg = x.cuda().contiguous(memory_format=torch.channels_last)
    y = torch.cat([head.expand_feature(xg, i) for i in range(4)], 0)
    gpu = [y[:P].float().cpu()]
    for gs in head._group_specs:
        y = D._GroupedConv3x3Fn.apply(y, gs)
        gpu.append(y[:P].float().cpu().contiguous())
    # the TF32 weight copy the kernels read vs rna(master weight)
    w0 = head.expand_convs[0][0].weight
    tw = w0._loft.w.detach().float().cpu().contiguous()
    print('T == rna(W):', bool(torch.equal(tw, E.rna(p[f'{pre}.expand_convs.0.0.weight']))),
          ' T == W:', bool(torch.equal(tw, p[f'{pre}.expand_convs.0.0.weight'])))
torch.cuda.synchronize()

a_e, a_f = x.clone(), x.clone()
print(f'{"layer":>5s} {"gpu-emu":>10s} {"gpu-fp32":>10s} {"emu-fp32":>10s} {"differ":>8s} {"mask!=":>8s}')
for c in range(10):
    w = p[f'{pre}.expand_convs.0.{c}.weight']
    b = p[f'{pre}.expand_convs.0.{c}.bias']
    a_e = E.rna(F.relu(F.conv2d(a_e, E.rna(w), b, padding=1)))
    a_f = F.relu(F.conv2d(a_f, w, b, padding=1))
    gy = gpu[c + 1]
    r = lambda u, v: float((u.double() - v.double()).norm() / v.double().norm())
    print(f'{c:5d} {r(gy, a_e):10.2e} {r(gy, a_f):10.2e} {r(a_e, a_f):10.2e} '
          f'{float((gy != a_e).float().mean()):8.4f} {float(((gy > 0) != (a_e > 0)).float().mean()):8.5f}')
