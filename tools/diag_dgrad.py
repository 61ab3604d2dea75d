"""Where is the error of a single-layer backward?  Grouped 3x3 conv (the FOA shape) and the
12544->1024 linear: error by position vs a float64 reference on identical TF32-grid operands."""
import os, sys, ctypes
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bonai_b200 import _lib as L
from bonai_b200.engine import WeightRef
from bonai_b200.ops import dense as D
import test_gpu_kernels as T

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
G, Pg, C = 4, 13, 256
wall = T.tf32_round(T.rnd(G, C + 1, 3, 3, C, seed=1, scale=0.03))
gwall = torch.zeros_like(wall)
ball = T.rnd(G, 260, seed=2)
gball = torch.zeros_like(ball)
wrefs = [WeightRef(wall[g, :C].permute(0, 3, 1, 2), gwall[g, :C].permute(0, 3, 1, 2)) for g in range(G)]
spec = D.GroupedConvSpec(wrefs, [ball[g, :C] for g in range(G)], [gball[g, :C] for g in range(G)],
                         relu=True, store=T._Store())
x = T.tf32_round(T.rnd(G * Pg, C, 7, 7, seed=3)).contiguous(memory_format=torch.channels_last)
xg = x.clone().requires_grad_(True)
y = D.grouped_conv3x3(xg, spec)
dy = T.tf32_round(T.rnd(*y.shape, seed=4))
y.backward(dy)
for g in range(G):
    xr = x[g * Pg:(g + 1) * Pg].double().clone().requires_grad_(True)
    wr = wrefs[g].w.double().clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, ball[g, :C].double(), padding=1))
    yr.backward(dy[g * Pg:(g + 1) * Pg].double())
    a, b = xg.grad[g * Pg:(g + 1) * Pg].double(), xr.grad
    d = (a - b)
    print(f'group {g}: dx rel {float(d.norm() / b.norm()):.3e}  max|d| {float(d.abs().max()):.3e} '
          f'(max|ref| {float(b.abs().max()):.2f})  mask mismatches '
          f'{int(((y[g * Pg:(g + 1) * Pg] > 0) != (yr > 0)).sum())}')
    e_hw = d.pow(2).sum((0, 1)).sqrt() / b.pow(2).sum((0, 1)).sqrt()
    print('   rel err by (h, w):', ' '.join(f'{v:.1e}' for v in e_hw.flatten().tolist()[:14]), '...')
    e_n = d.pow(2).sum((1, 2, 3)).sqrt() / b.pow(2).sum((1, 2, 3)).sqrt()
    print('   rel err by roi:', ' '.join(f'{v:.1e}' for v in e_n.tolist()))
    big = (d.abs() > 20 * d.abs().median()).float().mean()
    print(f'   median |d| {float(d.abs().median()):.3e}, fraction > 20x median: {float(big):.5f}')
    # is the error what TF32 rounding of the stored dx predicts?  compare with rna(ref)
    from oracle import tf32_emu as E
    rr = E.rna(b.float().cpu()).double().cuda()
    print(f'   vs rna(ref): rel {float((a - rr).norm() / b.norm()):.3e}; rna(ref) vs ref: '
          f'{float((rr - b).norm() / b.norm()):.3e}')
