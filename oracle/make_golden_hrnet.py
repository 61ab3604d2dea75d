"""Golden vectors for the HRNetV2p-W32 backbone + HRFPN neck (BASELINE.json configs[3]) from the
UNMODIFIED reference run over the import shim (oracle/ref_env.py) -- only usable where
/root/reference is mounted; writes tests/golden/hrnet_w32.npz.

    python oracle/make_golden_hrnet.py

Weights are not stored (29 M backbone parameters): both sides fill every state_dict entry from a
generator seeded by the CRC32 of its NAME (`name_seeded_state`), so the product model -- whose
state_dict must have the reference's names and shapes -- reconstructs identical weights.  Stored:
the state_dict manifest (names, shapes), the input tile, the 4 backbone and 5 neck outputs, and the
gradient norm of every parameter for loss = sum_i mean(out_i^2) over the neck outputs.
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HRNET = dict(
    type='HRNet',
    extra=dict(
        stage1=dict(num_modules=1, num_branches=1, block='BOTTLENECK', num_blocks=(4, ),
                    num_channels=(64, )),
        stage2=dict(num_modules=1, num_branches=2, block='BASIC', num_blocks=(4, 4),
                    num_channels=(32, 64)),
        stage3=dict(num_modules=4, num_branches=3, block='BASIC', num_blocks=(4, 4, 4),
                    num_channels=(32, 64, 128)),
        stage4=dict(num_modules=3, num_branches=4, block='BASIC', num_blocks=(4, 4, 4, 4),
                    num_channels=(32, 64, 128, 256))))
HRFPN = dict(type='HRFPN', in_channels=[32, 64, 128, 256], out_channels=256)


def name_seeded_state(state):
    """{name: tensor} with the shapes of `state`, every entry drawn from a generator seeded by
    crc32(name): conv / linear weights ~ N(0, 1/fan_in), BN gamma ~ U(.2, .6) (activations stay
    O(1) through ~50 residual blocks), beta / running_mean ~ N(0, .1), running_var ~ U(.5, 1.5)."""
    out = {}
    for k, v in state.items():
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_var'):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean'):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() == 1 and k.endswith('weight'):
            out[k] = torch.rand(v.shape, generator=g) * 0.4 + 0.2
        elif v.dim() == 1:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.0 / fan_in) ** 0.5
    return out


def loss_of(neck_outs):
    return sum((o * o).mean() for o in neck_outs)


def main():
    from oracle import ref_env
    ref_env.activate()
    from mmdet.models import build_backbone, build_neck
    bb, neck = build_backbone(dict(HRNET)), build_neck(dict(HRFPN))
    for pre, mod in (('backbone.', bb), ('neck.', neck)):
        st = name_seeded_state({pre + k: v for k, v in mod.state_dict().items()})
        mod.load_state_dict({k[len(pre):]: v for k, v in st.items()})
    bb.train()
    neck.train()
    g = torch.Generator().manual_seed(7)
    img = torch.randn(2, 3, 64, 96, generator=g)
    feats = bb(img)
    outs = neck(feats)
    loss_of(outs).backward()
    d = dict(img=img.numpy())
    names, shapes, gnames, gnorms = [], [], [], []
    for pre, mod in (('backbone.', bb), ('neck.', neck)):
        for k, v in mod.state_dict().items():
            names.append(pre + k)
            shapes.append(','.join(str(s) for s in v.shape))
        for k, p in mod.named_parameters():
            gnames.append(pre + k)
            gnorms.append(float(p.grad.double().norm()) if p.grad is not None else -1.0)
    d.update(names=np.array(names), shapes=np.array(shapes), grad_names=np.array(gnames),
             grad_norms=np.array(gnorms))
    for i, f in enumerate(feats):
        d[f'feat_{i}'] = f.detach().numpy()
    for i, o in enumerate(outs):
        d[f'out_{i}'] = o.detach().numpy()
    d['loss'] = np.array([float(loss_of(outs))])
    path = os.path.join(ROOT, 'tests', 'golden', 'hrnet_w32.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB; loss', d['loss'],
          'feat rms', [float(f.pow(2).mean().sqrt()) for f in feats])


if __name__ == '__main__':
    main()
