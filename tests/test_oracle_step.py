"""Whole-step pin: the portable CPU restatement reproduces the reference's losses and gradients
(fixture tests/golden/loft_step_256.npz, produced by running the unmodified reference over the
import shim -- oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import loft_cpu as O


@pytest.fixture(scope='module')
def oracle_run():
    p = O.randomize_bn(O.init_params(0), 0)
    tk = set(O.trainable_keys(p))
    p = {k: (v.clone().requires_grad_(True) if k in tk else v) for k, v in p.items()}
    img, gb, gl, gm, go = O.make_inputs(0, 1, 256, 10)
    torch.manual_seed(123)
    losses = O.forward_train(p, img, gb, gl, gm, go)
    loss, log_vars = O.parse_losses(losses)
    loss.backward()
    return p, log_vars


def test_losses_match_reference(golden_step, oracle_run):
    _, log_vars = oracle_run
    for name, val in zip(golden_step['loss_names'], golden_step['loss_values']):
        got = float(log_vars[str(name)])
        assert abs(got - val) <= 1e-5 * max(1.0, abs(val)), (name, got, val)


def test_grads_match_reference(golden_step, oracle_run):
    p, _ = oracle_run
    names = [str(n) for n in golden_step['grad_names']]
    assert len(names) == 254                      # 254 trainable tensors
    assert set(names) == {k for k in p if p[k].requires_grad}
    worst = 0.0
    for n, gn in zip(names, golden_step['grad_norms']):
        got = float(p[n].grad.double().norm())
        rel = abs(got - gn) / max(gn, 1e-12)
        worst = max(worst, rel)
    assert worst < 2e-3, worst


def test_reference_live_if_present():
    """When /root/reference is mounted (build container) run the reference itself side by side."""
    from oracle import ref_env
    if not ref_env.available():
        pytest.skip('reference tree not present')
    from oracle.make_golden import build_reference_model, reference_step
    p = O.randomize_bn(O.init_params(1), 1)
    model, _ = build_reference_model(p)
    img, gb, gl, gm, go = O.make_inputs(5, 1, 256, 6)
    ref_logs, _ = reference_step(model, img, gb, gl, gm, go, seed=7)
    torch.manual_seed(7)
    with torch.no_grad():
        losses = O.forward_train(p, img, gb, gl, gm, go)
    _, logs = O.parse_losses(losses)
    for k, v in ref_logs.items():
        assert abs(float(logs[k]) - v) <= 1e-5 * max(1.0, abs(v)), (k, float(logs[k]), v)


def test_golden_1024_step_inputs_reproduce_here():
    """tests/golden/loft_step_1024x2_g80.npz (BASELINE config) replays sampler draws / proposals
    against weights and inputs regenerated from seeds: the CPU RNG streams must reproduce them."""
    import os
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'tools'))
    g = dict(np.load(os.path.join(root, 'tests', 'golden', 'loft_step_1024x2_g80.npz')))
    from oracle import loft_cpu as O
    from oracle.make_golden_step import checksums
    size, n_img, num_gt, seed = (int(v) for v in g['meta'])
    assert (size, n_img, num_gt) == (1024, 2, 80)
    p = O.randomize_bn(O.init_params(seed), seed)
    img, gb, gl, gm, go = O.make_inputs(seed, n_img, size, num_gt)
    assert np.allclose(checksums(p, img, gb, go), g['checksums'], rtol=1e-9, atol=1e-6)
    names = [str(n) for n in g['loss_names']]
    assert names[-1] == 'loss' and abs(float(g['loss_values'][-1]) - 13.4505) < 1e-3
    assert len(g['grad_names']) == len(g['grad_norms']) >= 250
    assert all(g[f'proposals_{i}'].shape == (3000, 5) for i in range(n_img))
