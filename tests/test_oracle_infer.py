"""Inference pin: the CPU restatement of simple_test (bbox / mask / offset results) reproduces the
reference's outputs (tests/golden/loft_infer_256.npz from oracle/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import loft_cpu as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_simple_test_matches_reference():
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'loft_infer_256.npz')))
    p = O.randomize_bn(O.init_params(0), 0)
    img, _, _, _, _ = O.make_inputs(0, 1, 256, 10)
    dets, masks, offsets = O.simple_test(p, img)
    assert dets.shape == g['dets'].shape == (2000, 5)            # max_per_img
    assert torch.allclose(dets, torch.from_numpy(g['dets']), rtol=1e-5, atol=1e-5)
    assert torch.allclose(offsets, torch.from_numpy(g['offsets']), rtol=1e-4, atol=1e-4)
    areas = masks.flatten(1).sum(1).numpy()
    assert (np.abs(areas - g['mask_areas']) <= 1).mean() > 0.999


def test_soft_nms_properties():
    from oracle import ops_cpu
    g = torch.Generator().manual_seed(0)
    c = torch.rand(200, 2, generator=g) * 100
    wh = torch.rand(200, 2, generator=g) * 30 + 2
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    scores = torch.rand(200, generator=g)
    dets, keep = ops_cpu.soft_nms_linear(boxes, scores, 0.5, 1e-3)
    assert bool((dets[1:, 4] <= dets[:-1, 4] + 1e-7).all())      # selection order = score order
    assert len(set(keep.tolist())) == keep.numel()
    assert bool((dets[:, 4] <= scores[keep] + 1e-7).all())       # scores only decay
    assert float(dets[0, 4]) == float(scores.max())
