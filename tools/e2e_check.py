"""Whole-step parity on the GPU: product (CUDA kernels) vs the CPU oracle, teacher-forced
(identical weights, oracle's proposals and sampler draws injected).  Prints a JSON report."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loft_cpu as O  # noqa: E402
from bonai_b200 import Config  # noqa: E402
from bonai_b200.models import build_detector  # noqa: E402
from bonai_b200.core import BitmapMasks  # noqa: E402

CFG = os.path.join(ROOT, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py')


class teacher_force:
    """Context: make `model` draw the oracle's random samples and (optionally) use the oracle's
    proposals.  Nothing in the product knows about this: the sampler INSTANCES' `random_choice`
    and the RPN head INSTANCE's `get_bboxes` are replaced for the duration, and the proposal
    generation is kept out of the trunk's forward graph (LOFT_GRAPH_PROPOSALS=0) so that
    `get_bboxes` is what the step calls."""

    def __init__(self, model, draws=None, proposals=None):
        self.model, self.draws, self.proposals = model, draws, proposals
        self.samplers = [s for s in (getattr(model.rpn_head, 'sampler', None),
                                     getattr(model.roi_head, 'bbox_sampler', None))
                         if s is not None]

    def __enter__(self):
        if self.draws is not None:
            draws = list(self.draws)

            def forced_choice(gallery, num):
                return draws.pop(0).to(gallery.device)
            for s in self.samplers:
                s.random_choice = forced_choice
        self._env = os.environ.get('LOFT_GRAPH_PROPOSALS')
        if self.proposals is not None:
            props = self.proposals
            os.environ['LOFT_GRAPH_PROPOSALS'] = '0'
            self.model.rpn_head.get_bboxes = \
                lambda cls_scores, *a, **k: [p.to(cls_scores[0].device) for p in props]
        return self

    def __exit__(self, *a):
        for s in self.samplers:
            s.__dict__.pop('random_choice', None)
        self.model.rpn_head.__dict__.pop('get_bboxes', None)
        if self._env is None:
            os.environ.pop('LOFT_GRAPH_PROPOSALS', None)
        else:
            os.environ['LOFT_GRAPH_PROPOSALS'] = self._env
        return False


def _golden_inputs(golden):
    """Regenerate weights / inputs of a golden step (oracle/make_golden_step.py) and check that
    this machine's CPU RNG reproduced them."""
    import numpy as np
    from oracle.make_golden_step import checksums
    size, n_img, num_gt, seed = (int(v) for v in golden['meta'])
    p = O.randomize_bn(O.init_params(seed), seed)
    img, gb, gl, gm, go = O.make_inputs(seed, n_img, size, num_gt)
    cs = checksums(p, img, gb, go)
    if not np.allclose(cs, golden['checksums'], rtol=1e-9, atol=1e-6):
        raise RuntimeError(f'golden step inputs not reproduced on this machine: {cs} vs '
                           f'{golden["checksums"]}')
    return size, n_img, num_gt, seed, p, (img, gb, gl, gm, go)


def run(size=256, n_img=1, num_gt=10, seed=0, force_proposals=True, verbose=True, diag=False,
        golden=None, foa_emu=True):
    """One teacher-forced training step of the product on cuda:0 against the CPU oracle -- run
    live, or (golden = dict of oracle/make_golden_step.py arrays) replayed from a golden file."""
    if golden is not None:
        size, n_img, num_gt, seed, p, (img, gb, gl, gm, go) = _golden_inputs(golden)
        rec = [torch.from_numpy(golden[f'draw_{i}']) for i in range(int(golden['n_draws'][0]))]
        proposals = [torch.from_numpy(golden[f'proposals_{i}']) for i in range(n_img)]
        logs_o = dict(zip([str(n) for n in golden['loss_names']],
                          [float(v) for v in golden['loss_values']]))
        t_oracle = 0.0
        po = None
    else:
        p = O.randomize_bn(O.init_params(seed), seed)
        img, gb, gl, gm, go = O.make_inputs(seed, n_img, size, num_gt)
    H, W = (size, size) if isinstance(size, int) else size
    tk = set(O.trainable_keys(p))
    aux = {}
    if golden is None:
        po = {k: (v.clone().requires_grad_(True) if k in tk else v) for k, v in p.items()}
        torch.manual_seed(123)
        rec = []
        t0 = time.time()
        lo = O.forward_train(po, img, gb, gl, gm, go, record=rec, aux=aux, stable_sort=True)
        loss_o, logs_o = O.parse_losses(lo)
        loss_o.backward()
        t_oracle = time.time() - t0
        proposals = aux['proposals']

    cfg = Config.fromfile(CFG)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.load_state_dict(p)
    model.train()
    dev = torch.device('cuda:0')
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), ori_shape=(H, W, 3),
                  scale_factor=1.0, flip=False) for _ in range(n_img)]
    caps = {}

    def cap(name):
        def hook(mod, inp, out):
            caps[name] = out
        return hook
    if diag:
        model.neck.register_forward_hook(cap('feats'))
        model.roi_head.bbox_roi_extractor.register_forward_hook(cap('bbox_feats'))
        model.roi_head.mask_roi_extractor.register_forward_hook(cap('mask_feats'))
        model.roi_head.mask_head.register_forward_hook(cap('mask_pred'))
        model.roi_head.offset_head.register_forward_hook(cap('offset_pred'))
        model.roi_head.bbox_head.register_forward_hook(cap('bbox_out'))
    # inputs of the FOA head and its targets, for the TF32-emulated head oracle below
    # (the head's INPUT: in training the offset features are the positives' rows of the bbox
    # features, the offset extractor itself is not called)
    model.roi_head.offset_head.register_forward_pre_hook(
        lambda mod, inp: caps.__setitem__('offset_feats', inp[0]))
    oh = model.roi_head.offset_head
    get_targets = oh.get_targets

    def get_targets_cap(*a, **k):
        caps['offset_targets'] = get_targets(*a, **k)
        return caps['offset_targets']
    oh.get_targets = get_targets_cap
    # observe (not alter) the activations of the FOA head's layers: outputs of its grouped conv /
    # linear launches, for the layer-local (teacher-forced) check of its backward
    from bonai_b200.engine import get_store
    from bonai_b200.ops import dense as D
    get_store(model, dev)                    # builds the kernel-side specs (loft_prepare)
    foa_acts = {}
    own = {id(sp): ('conv', i) for i, sp in enumerate(getattr(oh, '_group_specs', None) or [])}
    own.update({id(sp): ('fc', i) for i, sp in enumerate(oh._fc_specs)})
    own[id(oh._head)] = ('out', 0)
    real_gconv, real_linear = D.grouped_conv3x3, D.linear

    def gconv_obs(x, spec, triggers=()):
        y = real_gconv(x, spec, triggers)
        if id(spec) in own:
            if own[id(spec)][1] == 0:
                foa_acts['in'] = x.detach()
            foa_acts[own[id(spec)]] = y.detach()
        return y

    def linear_obs(x, spec, triggers=()):
        y = real_linear(x, spec, triggers)
        if id(spec) in own:
            foa_acts[own[id(spec)]] = y.detach()
        return y
    D.grouped_conv3x3, D.linear = gconv_obs, linear_obs
    with teacher_force(model, [r.clone() for r in rec],
                       [q.clone() for q in proposals] if force_proposals else None):
        losses = model.forward_train(img.to(dev), metas, gb, gl,
                                     gt_masks=[BitmapMasks(m, H, W) for m in gm], gt_offsets=go)
        loss, logs = model._parse_losses(losses)
        loss.backward()
        torch.cuda.synchronize()
    D.grouped_conv3x3, D.linear = real_gconv, real_linear
    del oh.get_targets
    rep = {'losses': {}, 'grads': {}}
    if diag:
        import torch.nn.functional as F
        def st(name, a, b):
            a, b = a.detach().cpu().double(), b.detach().double()
            d = a - b
            print(f'{name:14s} rel_l2={float(d.norm() / b.norm()):.2e} '
                  f'signed_mean/abs_mean={float(d.mean() / b.abs().mean()):+.2e}')
        for i, (a, b) in enumerate(zip(caps['feats'], aux['feats'])):
            st(f'fpn{i}', a, b)
        st('bbox_feats', caps['bbox_feats'], aux['bbox_feats'])
        mo = O.roi_extract(aux['feats'][:4], aux['pos_rois'], 14)
        st('mask_feats', caps['mask_feats'], mo)
        st('mask_pred', caps['mask_pred'], aux['mask_pred'])
        st('offset_pred', caps['offset_pred'], aux['offset_pred'])
        st('cls_score', caps['bbox_out'][0], aux['cls_score'])
        mt = aux['mask_targets']
        lg = F.binary_cross_entropy_with_logits(caps['mask_pred'].detach().cpu()[:, 0], mt)
        lo_ = F.binary_cross_entropy_with_logits(aux['mask_pred'].detach()[:, 0], mt)
        print('mask loss from captured preds', float(lg), float(lo_), float((lg - lo_) / lo_))
        sr = model.roi_head._last_sampling_results
        print('mask targets equal:', bool(torch.equal(
            model.roi_head.mask_head.get_targets(sr, [BitmapMasks(m, H, W) for m in gm],
                                                 model.roi_head.train_cfg).cpu(), mt)))
    worst = 0.0
    for k, v in logs_o.items():
        a, b = float(logs[k]), float(v)
        rel = abs(a - b) / max(abs(b), 1e-6)
        rep['losses'][k] = [a, b, rel]
        if 'loss' in k:
            worst = max(worst, rel)
    rep['worst_loss_rel'] = worst
    # ---- parameter gradients vs the fp32 oracle: relative L2 of the difference (live oracle, and
    # the 1-D parameters of a golden), else relative difference of the norms (golden, big tensors)
    named = dict(model.named_parameters())
    if golden is not None:
        ref_norm = dict(zip([str(n) for n in golden['grad_names']],
                            [float(v) for v in golden['grad_norms']]))
    FOA = 'roi_head.offset_head.'
    gworst, gname, fworst, fname = 0.0, None, 0.0, None
    for k in sorted(tk):
        gp = named[k].grad
        if golden is None:
            go_ = po[k].grad
            if go_ is None or gp is None:
                rep['grads'][k] = 'missing'
                continue
            n = float(go_.double().norm())
            d = float((gp.detach().cpu().double() - go_.double()).norm())
        else:
            if k not in ref_norm or gp is None:
                rep['grads'][k] = 'missing'
                continue
            n = ref_norm[k]
            if 'grad/' + k in golden:
                d = float((gp.detach().cpu().double() -
                           torch.from_numpy(golden['grad/' + k]).double()).norm())
            else:
                d = abs(float(gp.detach().double().norm()) - n)
        rel = d / max(n, 1e-12)
        if n > 1e-8:
            if k.startswith(FOA):
                if rel > fworst:
                    fworst, fname = rel, k
            elif rel > gworst:
                gworst, gname = rel, k
        if verbose and n > 1e-8 and rel > 2e-2:
            rep['grads'][k] = [d, n, rel]
    rep['worst_grad_rel'] = [max(gworst, fworst), gname if gworst >= fworst else fname]
    rep['worst_grad_rel_non_foa'] = [gworst, gname]
    rep['worst_grad_rel_foa_vs_fp32'] = [fworst, fname]
    # ---- FOA head: the GPU's gradients against the TF32-EMULATED head oracle fed with the very
    # RoI features the GPU head consumed (oracle/tf32_emu.py), and the fp32 head on those same
    # inputs -- separates "TF32 operand rounding" from "a defect of the fused backward"
    if foa_emu and caps.get('offset_feats') is not None and caps['offset_feats'].shape[0] > 0:
        from oracle import tf32_emu as E
        x = caps['offset_feats'].detach().float().cpu().contiguous()
        tg = caps['offset_targets'].detach().float().cpu()
        _, g_emu = E.foa_head_forward_backward(x, p, tg, emulate=True)
        _, g_f32 = E.foa_head_forward_backward(x, p, tg, emulate=False)
        ew, en, dw, dn = 0.0, None, 0.0, None
        for k, ge in g_emu.items():
            gp = named[k].grad.detach().cpu().double()
            n = float(ge.double().norm())
            if n <= 1e-8:
                continue
            r_emu = float((gp - ge.double()).norm()) / n
            r_f32 = float((g_emu[k].double() - g_f32[k].double()).norm()) / \
                max(float(g_f32[k].double().norm()), 1e-12)
            if r_emu > ew:
                ew, en = r_emu, k
            if r_f32 > dw:
                dw, dn = r_f32, k
        rep['foa_grad_vs_tf32_emulated_oracle'] = [ew, en]
        rep['foa_tf32_emulated_vs_fp32_same_inputs'] = [dw, dn]
        # layer-local: fp32 back-propagation through the head with the GPU's own activations
        # (identical masks / saved operands) -- this is the check a backward defect cannot pass
        if ('conv', 0) in foa_acts:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            nconv = len(oh._group_specs)
            acts = [foa_acts['in'].float().contiguous()] + \
                [foa_acts[('conv', i)].float().contiguous() for i in range(nconv)]
            g_tf = E.foa_head_backward_teacher_forced(
                acts, foa_acts[('fc', 0)].float(), foa_acts[('fc', 1)].float(),
                foa_acts[('out', 0)][:, :2].float(), caps['offset_targets'].detach(), p)
            tw, tn = 0.0, None
            for k, gr in g_tf.items():
                n = float(gr.double().norm())
                if n <= 1e-8:
                    continue
                r_ = float((named[k].grad.detach().double() - gr.double()).norm()) / n
                if r_ > tw:
                    tw, tn = r_, k
            rep['foa_grad_teacher_forced_layerwise'] = [tw, tn]
    rep['oracle_seconds'] = t_oracle
    return rep


if __name__ == '__main__':
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    num_gt = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rep = run(size, n_img, num_gt, diag=len(sys.argv) > 4)
    print(json.dumps(rep, indent=1))
