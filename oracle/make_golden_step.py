"""Golden vectors of one whole LOFT training step from the CPU oracle (oracle/loft_cpu.py, itself
pinned against the unmodified reference by oracle/make_golden.py + tests/test_oracle_step.py).

    python oracle/make_golden_step.py [size n_img num_gt seed out.npz]

Default: BASELINE.json's own configuration -- 1024x1024 tiles, batch 2, G = 80 -- which takes the
oracle ~1 min per step, too long for the GPU box's test run; the GPU parity test
(tests/test_gpu_step.py::test_step_parity_1024_baseline_config_vs_golden) replays the recorded
sampler draws and proposals against this file instead.  Stored: every loss, the sampler draws (in
call order), the proposals, the L2 norm of every parameter gradient, the full gradient of every
1-D parameter (biases, BN affine), the offset targets / predictions of the FOA head, and checksums
of the inputs and weights (CPU RNG streams must reproduce them on the machine running the test).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loft_cpu as O  # noqa: E402


def checksums(p, img, gb, go):
    keys = sorted(p)
    w = torch.stack([p[k].double().sum() for k in keys]).sum()
    return np.array([float(img.double().sum()), float(torch.cat(gb).double().sum()),
                     float(torch.cat(go).double().sum()), float(w)], dtype=np.float64)


def main():
    a = sys.argv[1:]
    size, n_img, num_gt, seed = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 \
        else (1024, 2, 80, 0)
    out = a[4] if len(a) >= 5 else os.path.join(ROOT, 'tests', 'golden',
                                                f'loft_step_{size}x{n_img}_g{num_gt}.npz')
    torch.set_num_threads(os.cpu_count() or 1)
    p = O.randomize_bn(O.init_params(seed), seed)
    img, gb, gl, gm, go = O.make_inputs(seed, n_img, size, num_gt)
    tk = set(O.trainable_keys(p))
    po = {k: (v.clone().requires_grad_(True) if k in tk else v) for k, v in p.items()}
    torch.manual_seed(123)
    rec, aux = [], {}
    t0 = time.time()
    losses = O.forward_train(po, img, gb, gl, gm, go, record=rec, aux=aux, stable_sort=True)
    loss, logs = O.parse_losses(losses)
    loss.backward()
    print(f'oracle step: {time.time() - t0:.1f} s, losses {dict(logs)}')
    d = dict(meta=np.array([size, n_img, num_gt, seed], dtype=np.int64),
             checksums=checksums(p, img, gb, go),
             loss_names=np.array(list(logs.keys())),
             loss_values=np.array([float(v) for v in logs.values()], dtype=np.float64),
             n_draws=np.array([len(rec)], dtype=np.int64),
             offset_targets=aux['offset_targets'].detach().numpy(),
             offset_pred=aux['offset_pred'].detach().numpy(),
             num_pos=np.array([s['pos_bboxes'].shape[0] for s in aux['samples']], dtype=np.int64))
    for i, r in enumerate(rec):
        d[f'draw_{i}'] = r.numpy().astype(np.int64)
    for i, q in enumerate(aux['proposals']):
        d[f'proposals_{i}'] = q.detach().numpy().astype(np.float32)
    names, norms = [], []
    for k in sorted(tk):
        g = po[k].grad
        if g is None:
            continue
        names.append(k)
        norms.append(float(g.double().norm()))
        if g.dim() == 1:
            d['grad/' + k] = g.numpy().astype(np.float32)
    d['grad_names'] = np.array(names)
    d['grad_norms'] = np.array(norms, dtype=np.float64)
    np.savez_compressed(out, **d)
    print('wrote', out, os.path.getsize(out) // 1024, 'KiB')


if __name__ == '__main__':
    main()
