"""Diagnostic: signed mean error of each mask-head layer (GPU kernels vs torch fp32 on GPU)."""
import os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import Config
from bonai_b200.models import build_detector
from bonai_b200.engine import get_store
from bonai_b200.ops import dense as D
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = Config.fromfile(os.path.join(ROOT, 'configs/loft/loft_foa_r50_fpn_2x_b200.py'))
torch.manual_seed(0)
m = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
st = get_store(m, torch.device('cuda:0')); st.begin_step()
mh = m.roi_head.mask_head
P = 100
x = (torch.randn(P, 256, 14, 14, device='cuda').abs() * 0.5).contiguous(memory_format=torch.channels_last)
i = x.view(torch.int32); x = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
def stat(name, a, b):
    d = (a.double() - b.double())
    print(f'{name:12s} rel_l2={float(d.norm()/b.double().norm()):.2e} signed_mean/abs_mean={float(d.mean()/b.double().abs().mean()):+.2e}')
xr = x.clone(); xg = x
with torch.no_grad():
    for k, (cm, spec) in enumerate(zip(mh.convs, mh._conv_specs)):
        xg = D.conv(xg, spec)
        xr = F.relu(F.conv2d(xr, cm.conv.weight, cm.conv.bias, padding=1))
        stat(f'conv{k}', xg, xr)
    xg = D.deconv2x2(xg, mh._up_spec)
    xr = F.relu(F.conv_transpose2d(xr, mh.upsample.weight, mh.upsample.bias, stride=2))
    stat('deconv', xg, xr)
    fg = D.conv(xg, mh._logit_spec)[:, :1]
    fr = F.conv2d(xr, mh.conv_logits.weight, mh.conv_logits.bias)
    stat('logits', fg, fr)
    t = (torch.rand(P, 28, 28, device='cuda') > 0.5).float()
    lg = F.binary_cross_entropy_with_logits(fg.squeeze(1), t); lr = F.binary_cross_entropy_with_logits(fr.squeeze(1), t)
    print('loss', float(lg), float(lr), float((lg - lr) / lr))
    # same chain but feeding the reference activations layer by layer (isolates each layer)
    xr = x.clone()
    for k, (cm, spec) in enumerate(zip(mh.convs, mh._conv_specs)):
        yg = D.conv(xr.contiguous(memory_format=torch.channels_last), spec)
        xr = F.relu(F.conv2d(xr, cm.conv.weight, cm.conv.bias, padding=1))
        stat(f'iso conv{k}', yg, xr)
