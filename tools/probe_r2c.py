"""Single launches of the kernels added in the second half of round 2, at the bench's sizes, for
`ncu --set full` (tools/gpu_r2c_evidence.sh): direct stem conv (pack + GEMM), RPN sampling +
targets, narrow mask-logits head forward / backward, FOA gather + rotate."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bonai_b200 import Config
from bonai_b200.engine import WeightRef, get_store
from bonai_b200.models import build_detector
from bonai_b200.ops import dense as D, misc as M
from bonai_b200.ops.roi import take_rows_rot

dev = torch.device('cuda:0')
torch.manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
reps = 3

if which in ('all', 'stem'):
    img = torch.randn(2, 3, 1024, 1024, device=dev)
    w = torch.randn(64, 7, 7, 3, device=dev) * 0.1
    wp = M.pack_stem_weight(w)
    shift = torch.zeros(64, device=dev)
    for _ in range(reps):
        y = M.stem_conv(img, wp, None, shift)
    torch.cuda.synchronize()

if which in ('all', 'rpn'):
    cfg = Config.fromfile(bench.CFG)
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    get_store(model, dev)
    head = model.rpn_head
    data = bench.to_model_inputs(bench.make_batch(0, device=dev))
    sizes = head.featmap_sizes_for((1024, 1024))
    for _ in range(reps):
        head._build_targets(sizes, data['gt_bboxes'], data['img_metas'], dev)
    torch.cuda.synchronize()

if which in ('all', 'narrow'):
    P, C = 203, 256
    x = torch.randn(P, C, 28, 28, device=dev).contiguous(memory_format=torch.channels_last)
    x.requires_grad_()
    w = torch.zeros(4, C, device=dev)
    w[0] = torch.randn(C, device=dev) * 0.05
    gw, b, gb, cs = torch.zeros_like(w), torch.zeros(4, device=dev), torch.zeros(4, device=dev), \
        torch.zeros(C, device=dev)
    spec = D.ConvSpec(WeightRef(w, gw), ksize=1, bias=b, bias_grad=gb, round_out=False, premask_in=True)
    spec.in_colsum = cs
    spec.n_out = 1                 # the class-agnostic mask head: one real output channel
    dy = torch.zeros(P, 4, 28, 28, device=dev).contiguous(memory_format=torch.channels_last)
    dy[:, 0] = torch.randn(P, 28, 28, device=dev)
    for _ in range(reps):
        y = D.narrow_head(x, spec)
        y.backward(dy)
    torch.cuda.synchronize()

if which in ('all', 'rot'):
    K, P = 2048, 203
    f = torch.randn(K, 256, 7, 7, device=dev).contiguous(memory_format=torch.channels_last)
    f.requires_grad_()
    rows = torch.arange(P, device=dev)
    for _ in range(reps):
        full, y = take_rows_rot(f, rows, (0, 1, 2, 3))
        (y.sum() + full.sum()).backward()
    torch.cuda.synchronize()
print('done', which)
