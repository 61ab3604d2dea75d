"""Post-load pipeline (flip + normalize + pad + format) on one 1024^2 tile with 80 buildings:
device kernels (CUDA events, HBM GB/s) next to the CPU restatement of the reference transforms."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200.datasets import image_prep, mask_flip_pad
from oracle import pipeline_cpu as P
rng = np.random.RandomState(0)
H = W = 1024; G = 80
img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
masks = (rng.rand(G, H, W) > 0.5).astype(np.uint8)
mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
img_d, m_d = torch.from_numpy(img).cuda(), torch.from_numpy(masks).cuda()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
t_img = timeit(lambda: image_prep(img_d, mean, std, True, 'horizontal', 32))
t_msk = timeit(lambda: mask_flip_pad(m_d, 'horizontal', 32))
b_img = H * W * 3 + 3 * H * W * 4
b_msk = 2 * G * H * W
t0 = time.time()
for _ in range(3):
    P.train_pipeline(img, np.zeros((G, 4), np.float32), masks, np.zeros((G, 2), np.float32), True,
                     'horizontal', mean, std)
t_cpu = (time.time() - t0) / 3
print(json.dumps(dict(image_prep_us=round(t_img * 1e6, 1), image_prep_GBs=round(b_img / t_img / 1e9, 1),
                      mask_flip_pad_us=round(t_msk * 1e6, 1), mask_flip_pad_GBs=round(b_msk / t_msk / 1e9, 1),
                      tile_us=round((t_img + t_msk) * 1e6, 1), cpu_numpy_tile_ms=round(t_cpu * 1e3, 1),
                      algorithmic_bytes_per_tile=b_img + b_msk)))

# polygon -> bitmap (LoadAnnotations poly2mask): 80 rotated-rectangle / star buildings per tile
from bonai_b200 import _lib as L
import ctypes
from bonai_b200.datasets import polygons_to_bitmaps
from oracle.polygon_cpu import poly2mask
polys = []
for i in range(G):
    cx, cy = rng.uniform(50, W - 50), rng.uniform(50, H - 50)
    k = 4 if i % 2 == 0 else 12
    ang = np.sort(rng.uniform(0, 2 * np.pi, k)); r = rng.uniform(10, 60, k)
    polys.append([np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1).reshape(-1).tolist()])
for _ in range(3):
    polygons_to_bitmaps(polys, H, W, 'cuda')
torch.cuda.synchronize()
t0 = time.time()
for _ in range(10):
    polygons_to_bitmaps(polys, H, W, 'cuda')          # includes the wrapper's H2D copies + one sync
torch.cuda.synchronize()
t_poly = (time.time() - t0) / 10
t0 = time.time()
for p in polys[:20]:
    poly2mask(p, H, W)
t_poly_cpu = (time.time() - t0) / 20 * G
print(json.dumps(dict(poly_rasterize_tile_us_wall=round(t_poly * 1e6, 1),
                      poly_rasterize_GBs_written=round(G * H * W / t_poly / 1e9, 1),
                      cpu_numpy_restatement_tile_ms=round(t_poly_cpu * 1e3, 1))))
