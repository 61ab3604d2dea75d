"""soft-NMS (linear) kernel alone: time per call for n candidates (LOFT_SOFT_NMS_FAST=0|1)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonai_b200.ops import soft_nms
g = torch.Generator().manual_seed(0)
for n in (500, 1000, 2000, 3000):
    c = torch.rand(n, 2, generator=g) * 1024
    wh = torch.exp(torch.rand(n, 2, generator=g) * 2.5) * 8
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024).cuda()
    scores = (torch.rand(n, generator=g) * 0.9 + 0.06).cuda()
    idxs = torch.zeros(n, dtype=torch.long).cuda()
    for _ in range(2):
        d, k = soft_nms(boxes, scores, 0.5, min_score=0.05, idxs=idxs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d, k = soft_nms(boxes, scores, 0.5, min_score=0.05, idxs=idxs)
    torch.cuda.synchronize()
    print(n, 'kept', k.numel(), 'ms/call', round((time.perf_counter() - t0) / 5 * 1e3, 3), 'ptr%16', boxes.data_ptr() % 16)
