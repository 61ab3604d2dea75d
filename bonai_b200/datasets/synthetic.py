"""Synthetic BONAI-like tile batches with the reference's input contract (what the train pipeline
of configs/_base_/datasets/bonai_instance.py:5-17 hands to LOFT.forward_train): a normalised fp32
image, G building boxes with log-uniform sides, an inscribed-ellipse roof bitmap per building
(uint8, BitmapMasks layout) and a roof-to-footprint offset vector in U(-40, 40) px.  The recipe is
SURVEY.md section 8(d); bench.py, tools/train.py and the tests draw their inputs from here (the CPU
oracle keeps an independent restatement, checked equal in tests/test_host.py)."""
import math

import torch


def make_inputs(seed, n_img, size, num_gt, device='cpu'):
    """Returns (img [n,3,H,W], gt_bboxes, gt_labels, gt_masks (uint8 [G,H,W]), gt_offsets), lists
    with one entry per image.  `size`: int or (H, W); `num_gt`: int or one count per image."""
    H, W = (size, size) if isinstance(size, int) else size
    counts = [num_gt] * n_img if isinstance(num_gt, int) else list(num_gt)
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(n_img, 3, H, W, generator=g)
    lim = torch.tensor([W, H, W, H], dtype=torch.float32)
    smin, smax = 16.0, min(160.0, min(H, W) / 2.0)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                            torch.arange(W, dtype=torch.float32), indexing='ij')
    boxes, labels, masks, offsets = [], [], [], []
    for n in counts:
        centre = torch.rand(n, 2, generator=g) * lim[:2]
        side = torch.exp(torch.rand(n, 2, generator=g) * math.log(smax / smin)) * smin
        b = torch.min(torch.cat([centre - side / 2, centre + side / 2], dim=1).clamp(min=0), lim)
        thin = (b[:, 2:] - b[:, :2]) < 2                      # keep every side >= 2 px
        b[:, 2:] = torch.where(thin, torch.min(b[:, :2] + 2, lim[2:]), b[:, 2:])
        b[:, :2] = torch.min(b[:, :2], b[:, 2:] - 2)
        cx, cy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
        rx, ry = (b[:, 2] - b[:, 0]) / 2, (b[:, 3] - b[:, 1]) / 2
        inside = (((xx[None] + 0.5 - cx[:, None, None]) / rx[:, None, None]) ** 2 +
                  ((yy[None] + 0.5 - cy[:, None, None]) / ry[:, None, None]) ** 2) <= 1.0
        boxes.append(b.to(device))
        labels.append(torch.zeros(n, dtype=torch.long, device=device))
        masks.append(inside.to(torch.uint8).to(device))
        offsets.append((torch.rand(n, 2, generator=g) * 80 - 40).to(device))
    return img.to(device), boxes, labels, masks, offsets
