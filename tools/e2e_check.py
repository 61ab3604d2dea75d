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


def run(size=256, n_img=1, num_gt=10, seed=0, force_proposals=True, verbose=True, diag=False):
    p = O.randomize_bn(O.init_params(seed), seed)
    img, gb, gl, gm, go = O.make_inputs(seed, n_img, size, num_gt)
    H, W = (size, size) if isinstance(size, int) else size
    tk = set(O.trainable_keys(p))
    po = {k: (v.clone().requires_grad_(True) if k in tk else v) for k, v in p.items()}
    torch.manual_seed(123)
    rec, aux = [], {}
    t0 = time.time()
    lo = O.forward_train(po, img, gb, gl, gm, go, record=rec, aux=aux, stable_sort=True)
    loss_o, logs_o = O.parse_losses(lo)
    loss_o.backward()
    t_oracle = time.time() - t0

    cfg = Config.fromfile(CFG)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.load_state_dict(p)
    model.train()
    dev = torch.device('cuda:0')
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), ori_shape=(H, W, 3),
                  scale_factor=1.0, flip=False) for _ in range(n_img)]
    caps = {}
    if diag:
        def cap(name):
            def hook(mod, inp, out):
                caps[name] = out
            return hook
        model.neck.register_forward_hook(cap('feats'))
        model.roi_head.bbox_roi_extractor.register_forward_hook(cap('bbox_feats'))
        model.roi_head.mask_roi_extractor.register_forward_hook(cap('mask_feats'))
        model.roi_head.mask_head.register_forward_hook(cap('mask_pred'))
        model.roi_head.offset_head.register_forward_hook(cap('offset_pred'))
        model.roi_head.bbox_head.register_forward_hook(cap('bbox_out'))
    with teacher_force(model, [r.clone() for r in rec],
                       [q.clone() for q in aux['proposals']] if force_proposals else None):
        losses = model.forward_train(img.to(dev), metas, gb, gl,
                                     gt_masks=[BitmapMasks(m, H, W) for m in gm], gt_offsets=go)
        loss, logs = model._parse_losses(losses)
        loss.backward()
        torch.cuda.synchronize()
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
        from bonai_b200.core import mask_target as MT
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
    gworst, gname = 0.0, None
    named = dict(model.named_parameters())
    for k in tk:
        go_ = po[k].grad
        gp = named[k].grad
        if go_ is None or gp is None:
            rep['grads'][k] = 'missing'
            continue
        d = float((gp.detach().cpu().double() - go_.double()).norm())
        n = float(go_.double().norm())
        rel = d / max(n, 1e-12)
        if n > 1e-8 and rel > gworst:
            gworst, gname = rel, k
        if verbose and n > 1e-8 and rel > 2e-2:
            rep['grads'][k] = [d, n, rel]
    rep['worst_grad_rel'] = [gworst, gname]
    rep['oracle_seconds'] = t_oracle
    return rep


if __name__ == '__main__':
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    num_gt = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rep = run(size, n_img, num_gt, diag=len(sys.argv) > 4)
    print(json.dumps(rep, indent=1))
