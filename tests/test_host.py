"""CPU tests of the host-side mirror of the reference interface: config surface, registries,
model construction (state_dict = reference manifest), C-ABI exports, LR schedule, and the
world_size-2 gradient exchange over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OWN_CFG = os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py')
REF_CFG = '/root/reference/configs/loft_foa/loft_foa_r50_fpn_2x_bonai.py'


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


def test_config_loads_and_merges():
    from bonai_b200 import Config
    cfg = Config.fromfile(OWN_CFG)
    assert cfg.model.type == 'LOFT'
    assert cfg.train_cfg.rcnn.sampler.num == 1024 and cfg['train_cfg']['rpn']['sampler']['num'] == 512
    assert cfg.get('nonexistent', 7) == 7
    assert cfg.model.roi_head.offset_head.loss_offset.loss_weight == 16.0
    cfg.merge_from_dict({'optimizer.lr': 0.1, 'model.backbone.depth': 101})
    assert cfg.optimizer.lr == 0.1 and cfg.model.backbone.depth == 101
    assert cfg.optimizer.momentum == 0.9


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason='reference tree not present')
def test_reference_config_loads_unchanged_and_equals_own():
    from bonai_b200 import Config
    ref = Config.fromfile(REF_CFG)            # _base_ x4, expressions (8*2.0, 0.02/4), loops
    own = Config.fromfile(OWN_CFG)
    assert ref.data.samples_per_gpu == 2 and ref.dataset_type == 'BONAI'
    for k in ['train_cfg', 'test_cfg', 'optimizer', 'optimizer_config', 'lr_config',
              'total_epochs', 'dist_params', 'workflow', 'log_config', 'checkpoint_config']:
        assert _plain(ref[k]) == _plain(own[k]), k
    m = _plain(ref.model)
    assert m.pop('pretrained') == 'torchvision://resnet50'
    o = _plain(own.model)
    o.pop('pretrained')
    assert m == o


def test_config_base_delete(tmp_path):
    from bonai_b200 import Config
    (tmp_path / 'base.py').write_text("a = dict(x=1, y=dict(p=1, q=2))\nb = [1, 2]\n")
    (tmp_path / 'child.py').write_text(
        "_base_ = './base.py'\na = dict(y=dict(_delete_=True, r=3), z=8*2.0)\n")
    cfg = Config.fromfile(str(tmp_path / 'child.py'))
    assert _plain(cfg.a) == dict(x=1, y=dict(r=3), z=16.0) and cfg.b == [1, 2]


def test_registry_both_decorator_styles():
    from bonai_b200 import Registry, build_from_cfg
    R = Registry('t')

    @R.register_module()
    class A:
        def __init__(self, v=1):
            self.v = v

    @R.register_module
    class B:
        pass

    assert build_from_cfg(dict(type='A', v=3), R).v == 3
    assert isinstance(build_from_cfg(dict(type='B'), R), B)
    assert build_from_cfg(dict(type='A'), R, dict(v=9)).v == 9
    with pytest.raises(KeyError):
        build_from_cfg(dict(type='C'), R)
    with pytest.raises(KeyError):
        R.register_module()(A)


@pytest.fixture(scope='module')
def model():
    from bonai_b200 import Config
    from bonai_b200.models import build_detector
    cfg = Config.fromfile(OWN_CFG)
    return build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)


def test_model_matches_reference_manifest(model):
    """446 state_dict entries, 81 472 088 params, 81 246 744 trainable, 254 trainable
    tensors; names and shapes equal to the (reference-validated) oracle initialiser."""
    from oracle import loft_cpu as O
    sd = model.state_dict()
    ref = O.init_params(0)
    assert len(sd) == 446 and set(sd) == set(ref)
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
    assert sum(p.numel() for p in model.parameters()) == 81472088
    tr = [n for n, p in model.named_parameters() if p.requires_grad]
    assert sum(dict(model.named_parameters())[n].numel() for n in tr) == 81246744
    assert len(tr) == 254 and set(tr) == set(O.trainable_keys(ref))
    model.load_state_dict(ref)                     # checkpoint-format compatible


def test_freeze_and_norm_eval_semantics(model):
    # tests/test_models/test_backbones.py:381-409 of the reference
    bb = model.backbone
    bb.train()
    assert not bb.bn1.training
    for p in list(bb.conv1.parameters()) + list(bb.bn1.parameters()) + list(bb.layer1.parameters()):
        assert not p.requires_grad
    for m in bb.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            assert not m.training
    assert all(p.requires_grad for p in bb.layer2.parameters())
    # zero_init_residual (resnet.py:601-619)
    assert float(bb.layer3[0].bn3.weight.abs().sum()) == 0.0


def test_roi_layer_lookup_by_name(model):
    ext = model.roi_head.bbox_roi_extractor
    assert ext.roi_layers[0].output_size == (7, 7) and ext.roi_layers[0].aligned
    assert model.roi_head.mask_roi_extractor.roi_layers[2].spatial_scale == 1 / 16
    assert model.roi_head.with_offset and model.roi_head.with_mask and model.with_rpn


def test_product_fails_loudly_without_cuda(model):
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from bonai_b200._lib import LoftError
    img = torch.zeros(1, 3, 64, 64)
    with pytest.raises(LoftError):
        model.forward_train(img, [dict(img_shape=(64, 64, 3), pad_shape=(64, 64, 3))],
                            [torch.zeros(0, 4)], [torch.zeros(0, dtype=torch.long)])


def test_trunk_eligibility_and_tape_replay(model, monkeypatch):
    """Host logic of bonai_b200/trunk.py that needs no GPU: the LOFT R50-FPN-RPN composition is
    recognised (and LOFT_TRUNK=0 turns the recorded path off); a Tape records what `_lib.call` is
    asked to launch and replays it with the CURRENT stream substituted for the recorded one."""
    from bonai_b200 import _lib as L
    from bonai_b200 import trunk as T
    assert T.Trunk.eligible(model)
    monkeypatch.setenv('LOFT_TRUNK', '0')
    assert not T.Trunk.eligible(model)
    monkeypatch.delenv('LOFT_TRUNK')
    model.backbone.style = 'caffe'
    try:
        assert not T.Trunk.eligible(model)
    finally:
        model.backbone.style = 'pytorch'

    seen = []

    class FakeLib:
        def __getattr__(self, name):
            def fn(*args):
                seen.append((name, args))
                return 0
            return fn

        def loft_last_error(self):
            return b''

    monkeypatch.setattr(L, '_lib', FakeLib())
    monkeypatch.setattr(L, 'stream', lambda: 'STREAM-NOW')
    tape = T.Tape('t')
    with tape:
        L.call('fill', 1, 2, 3.0, 'STREAM-AT-RECORD')
        L.call('add', 'a', 'b', 'c', 4, 0, 'STREAM-AT-RECORD')
    assert L.RECORD is None and [c[0] for c in tape.calls] == ['fill', 'add']
    n0 = len(seen)
    before = L.LAUNCHES[0]
    tape.launch()
    assert L.LAUNCHES[0] == before + 2
    replayed = seen[n0:]
    assert [r[0] for r in replayed] == ['loft_fill', 'loft_add']
    assert all(r[1][-1] == 'STREAM-NOW' for r in replayed)
    assert replayed[0][1][:-1] == (1, 2, 3.0)


def test_abi_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, 'include', 'loft_b200.h')).read()
    declared = set(re.findall(r'\b(loft_[a-z0-9_]+)\s*\(', hdr))
    declared.discard('loft_epilogue_t')
    lib = ctypes.CDLL(os.path.join(ROOT, 'bonai_b200', 'libloft_b200.so'))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert len(declared) >= 38
    lib.loft_abi_version.restype = ctypes.c_int
    assert lib.loft_abi_version() == 3
    # argument validation happens before any CUDA call
    lib.loft_last_error.restype = ctypes.c_char_p
    rc = lib.loft_copy2d(None, ctypes.c_longlong(1), None, ctypes.c_longlong(1),
                         ctypes.c_longlong(1), 1, 0, 0, None)
    assert rc == -1 and b'null' in lib.loft_last_error()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'bonai_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), (dirpath, f)


def test_step_lr_schedule():
    from bonai_b200.apis import step_lr
    kw = dict(step=[16, 22], warmup='linear', warmup_iters=300, warmup_ratio=0.001)
    assert step_lr(0.005, 0, 0, **kw) == pytest.approx(0.005 * 0.001)
    assert step_lr(0.005, 150, 0, **kw) == pytest.approx(0.005 * (1 - 0.5 * 0.999))
    assert step_lr(0.005, 300, 0, **kw) == pytest.approx(0.005)
    assert step_lr(0.005, 9999, 16, **kw) == pytest.approx(0.0005)
    assert step_lr(0.005, 9999, 23, **kw) == pytest.approx(0.00005)


def test_anchor_generator_cpu(golden_units):
    from bonai_b200.core import AnchorGenerator
    ag = AnchorGenerator(strides=[4, 8, 16, 32, 64], ratios=[0.5, 1.0, 2.0], scales=[8])
    sizes = [tuple(int(v) for v in s) for s in golden_units['anchor_sizes']]
    for i, a in enumerate(ag.grid_anchors(sizes, device='cpu')):
        assert torch.equal(a, torch.from_numpy(golden_units[f'anchors_l{i}']))
    assert ag.num_base_anchors == [3] * 5
    flags = ag.valid_flags([(4, 4)] + [(1, 1)] * 4, (12, 12, 3), device='cpu')
    assert int(flags[0].sum()) == 3 * 3 * 3


def test_offset_coder_and_fusion_cpu(golden_units):
    from bonai_b200.core import DeltaXYOffsetCoder, DeltaXYWHBBoxCoder
    from bonai_b200.models.roi_heads.attribute_heads import OffsetHeadExpandFeature
    g = golden_units
    props = torch.from_numpy(g['coder_props'])
    oc = DeltaXYOffsetCoder()
    assert torch.equal(oc.encode(props, torch.from_numpy(g['offset_gt'])),
                       torch.from_numpy(g['offset_encoded']))
    assert torch.equal(oc.decode(props, torch.from_numpy(g['offset_deltas']),
                                 max_shape=[1024, 1024]), torch.from_numpy(g['offset_decoded']))
    bc = DeltaXYWHBBoxCoder(target_stds=[0.1, 0.1, 0.2, 0.2])
    assert torch.equal(bc.decode(props, torch.from_numpy(g['coder_big_deltas']),
                                 max_shape=(100, 120)), torch.from_numpy(g['coder_decoded']))
    head = OffsetHeadExpandFeature(expand_feature_num=4, share_expand_fc=True, num_convs=1,
                                   loss_offset=dict(type='SmoothL1Loss', loss_weight=16.0))
    assert torch.equal(head.offset_fusion(torch.from_numpy(g['foa_pred'])),
                       torch.from_numpy(g['foa_fused']))
    with pytest.raises(NotImplementedError):
        OffsetHeadExpandFeature(rotations=[0, 45, 90, 135], share_expand_fc=True)


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bonai_b200.apis.train import allreduce_flat, bucket_views
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
views = bucket_views(g, bucket_bytes=1024)
assert sum(v.numel() for v in views) == 1000 and len(views) == 4
assert views[0].data_ptr() > views[-1].data_ptr()            # last parameters first
allreduce_flat(g, bucket_bytes=1024)
exp = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
assert torch.equal(g, exp), (g[:5], exp[:5])
# packed log-scalar mean (detectors/base.py:201-206 as one collective)
packed = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1) / world
dist.all_reduce(packed)
assert torch.allclose(packed, torch.tensor([1.5, 3.0, 4.5]))
dist.destroy_process_group()
print('ok', rank)
'''


def test_gradient_exchange_world2_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    port = 29500 + os.getpid() % 2000
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_bench_reference_arm_schema(monkeypatch, capsys):
    """`bench.py --impl reference` prints one JSON line with the contract's keys, the product arm's
    metric / unit / workload name, and only on rank 0 (the CPU step itself is mocked here; it is
    exercised for real by `python bench.py --impl reference`)."""
    import argparse
    import json
    import bench
    monkeypatch.setattr(bench, 'cpu_oracle_step', lambda n_img, threads: (lambda: 0.01))
    args = argparse.Namespace(gpus=1, steps=3, warmup=1)
    monkeypatch.setenv('RANK', '0')
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == bench.METRIC and line['unit'] == 'img/s'
    assert line['higher_is_better'] is True and line['steps'] == 3 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'img/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['config']['workload'].startswith('LOFT offset_rcnn R50-FPN 2x, 1024x1024')
    monkeypatch.setenv('RANK', '1')                     # other ranks exit without work or output
    bench.run_reference(args)
    assert capsys.readouterr().out.strip() == ''


def test_synthetic_inputs_equal_the_oracles_recipe():
    """bonai_b200.datasets.make_inputs (what bench.py feeds the GPU arm) and the oracle's own
    restatement of the SURVEY 8(d) recipe draw identical batches."""
    from bonai_b200.datasets import make_inputs
    from oracle import loft_cpu as O
    a = make_inputs(3, 2, (96, 128), [5, 2])
    b = O.make_inputs(3, 2, (96, 128), [5, 2])
    assert torch.equal(a[0], b[0])
    for x, y in zip(a[1:], b[1:]):
        assert all(torch.equal(u, v) for u, v in zip(x, y))


def test_batched_copies_collects_copy2d_into_one_call(monkeypatch):
    """_lib.batched_copies: every call('copy2d') inside the block becomes one job of a single
    loft_copy2d_multi launch at its end; other entry points run immediately; a recording launch
    program (trunk.Tape) and nesting pass straight through."""
    from bonai_b200 import _lib as L
    calls = []

    class Fake:
        def __getattr__(self, name):
            def fn(*a):
                calls.append((name, a))
                return 0
            return fn
    monkeypatch.setattr(L, 'lib', lambda: Fake())
    i32, ll, vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
    st = vp(7)

    def cp(src, dst, rows, cols, acc):
        L.call('copy2d', vp(src), ll(cols), vp(dst), ll(cols), ll(rows), i32(cols), i32(acc), i32(0), st)

    with L.batched_copies():
        cp(0x1000, 0x2000, 3, 256, 0)
        L.call('fill', vp(0x3000), ll(16), ctypes.c_float(0.0), st)       # not a copy: runs now
        with L.batched_copies():                                          # nested: same batch
            cp(0x1100, 0x2100, 1, 12, 1)
        assert [c[0] for c in calls] == ['loft_fill']
    assert [c[0] for c in calls] == ['loft_fill', 'loft_copy2d_multi']
    arr, n, stream = calls[-1][1]
    assert n.value == 2 and stream.value == 7
    assert (arr[0].src, arr[0].dst, arr[0].rows, arr[0].cols, arr[0].accumulate) == \
        (0x1000, 0x2000, 3, 256, 0)
    assert (arr[1].src, arr[1].dst, arr[1].lds, arr[1].ldd, arr[1].accumulate) == \
        (0x1100, 0x2100, 12, 12, 1)
    calls.clear()
    with L.batched_copies():                                              # nothing collected
        pass
    assert calls == []
    rec = []
    monkeypatch.setattr(L, 'RECORD', rec)
    with L.batched_copies():                                              # recording: not deferred
        cp(0x1000, 0x2000, 3, 256, 0)
        assert [c[0] for c in calls] == ['loft_copy2d'] and rec[0][0] == 'copy2d'


def test_offset_branch_shares_the_bbox_roi_features(model):
    """The LOFT config's offset extractor is the bbox extractor (same RoIAlign, strides, scale):
    LoftRoIHead feeds the offset head from the bbox features' rows; a different extractor or the
    LOFT_SHARE_OFFSET_ROI=0 switch keeps the separate RoIAlign."""
    rh = model.roi_head
    assert rh._shares_bbox_rois()
    os.environ['LOFT_SHARE_OFFSET_ROI'] = '0'
    try:
        assert not rh._shares_bbox_rois()
    finally:
        del os.environ['LOFT_SHARE_OFFSET_ROI']
    saved = rh.offset_roi_extractor.featmap_strides
    rh.offset_roi_extractor.featmap_strides = [4, 8, 16]
    try:
        assert not rh._shares_bbox_rois()
    finally:
        rh.offset_roi_extractor.featmap_strides = saved
    assert not rh.mask_roi_extractor.roi_layers[0].output_size == \
        rh.bbox_roi_extractor.roi_layers[0].output_size
