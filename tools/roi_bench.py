"""RoIAlign forward / backward alone on the step's shapes: time, algorithmic bytes (output written +
as much read), achieved GB/s."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200 import ops
from oracle import loft_cpu as O

dev = 'cuda'
g = torch.Generator().manual_seed(0)
feats = [torch.randn(2, 256, 1024 // s, 1024 // s, generator=g).to(dev).contiguous(
    memory_format=torch.channels_last).requires_grad_(True) for s in (4, 8, 16, 32)]
_, gb, _, _, _ = O.make_inputs(0, 2, 1024, 80)


def rois_like_step(K_per_img):
    rs = []
    for i in range(2):
        c = torch.rand(K_per_img, 2, generator=g) * 1024
        wh = torch.exp(torch.rand(K_per_img, 2, generator=g) * 2.3) * 16
        b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024)
        rs.append(torch.cat([torch.full((K_per_img, 1), float(i)), b], 1))
    return torch.cat(rs).to(dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for name, K, S in (('bbox  K=2048 7x7', 1024, 7), ('mask  P=216 14x14', 108, 14),
                   ('offset P=216 7x7', 108, 7)):
    rois = rois_like_step(K)
    out = ops.multilevel_roi_align(feats, rois, S, [4, 8, 16, 32], 56)
    dy = torch.randn_like(out)
    t_f = timeit(lambda: ops.multilevel_roi_align(feats, rois, S, [4, 8, 16, 32], 56))

    def bwd():
        o = ops.multilevel_roi_align(feats, rois, S, [4, 8, 16, 32], 56)
        o.backward(dy)
    t_fb = timeit(bwd)
    nbytes = out.numel() * 4
    print(f'{name}: fwd {t_f:7.1f} us  ({2 * nbytes / t_f / 1e3:6.0f} GB/s algorithmic: write '
          f'{nbytes / 1e6:.0f} MB + read as much)   fwd+bwd(incl. zero-filled grads) {t_fb:7.1f} us')
