"""BitmapMasks container and mask-target sampling (mmdet/core/mask/structures.py,
mask_target.py:6-62).  Unlike the reference -- which ships proposals to the host, uploads every
GT bitmap as fp32 and downloads the bool result (4 PCIe crossings) -- the uint8 bitmaps stay
resident on the device and one kernel samples all positives of the batch."""
import numpy as np
import torch

from ..ops import mask_target_sample


class BitmapMasks:
    """uint8 instance bitmaps [G,H,W]; accepts numpy (reference convention) or a device tensor."""

    def __init__(self, masks, height, width):
        self.height = height
        self.width = width
        if isinstance(masks, torch.Tensor):
            self._t = masks.to(torch.uint8).reshape(-1, height, width)
            self._np = None
        else:
            if len(masks) == 0:
                self._np = np.empty((0, height, width), dtype=np.uint8)
            else:
                if isinstance(masks, list):
                    masks = np.stack(masks)
                self._np = np.asarray(masks, dtype=np.uint8).reshape(-1, height, width)
            self._t = None

    @property
    def masks(self):
        if self._np is None:
            self._np = self._t.cpu().numpy()
        return self._np

    def __len__(self):
        return self._t.shape[0] if self._t is not None else self._np.shape[0]

    def to_tensor(self, dtype=torch.uint8, device='cuda'):
        if self._t is None or self._t.device != torch.device(device):
            src = self._t if self._t is not None else torch.from_numpy(self._np)
            self._t = src.to(device=device, non_blocking=True).contiguous()
        return self._t if dtype == torch.uint8 else self._t.to(dtype)

    def to_ndarray(self):
        return self.masks


def mask_target(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, cfg):
    """Same signature / result as the reference's mask_target: [sum P, S, S] float {0,1}."""
    size = cfg.mask_size if hasattr(cfg, 'mask_size') else cfg['mask_size']
    if isinstance(size, (tuple, list)):
        assert size[0] == size[1]
        size = size[0]
    dev = pos_proposals_list[0].device
    if sum(p.size(0) for p in pos_proposals_list) == 0:
        return pos_proposals_list[0].new_zeros((0, size, size))
    outs = []
    for props, gi, gm in zip(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list):
        m = gm.to_tensor(device=dev) if isinstance(gm, BitmapMasks) else gm.to(dev).to(torch.uint8)
        outs.append(mask_target_sample(m.contiguous(), props, gi, int(size)))
    return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


def encode_mask_results(masks):
    """Run-length encode instance bitmaps on the device -- the packing step of the reference's
    test loop (mmdet/apis/test.py:53-74 `encode_mask_results` -> pycocotools `mask.encode`, COCO
    RLE: run lengths over the column-major flattening, starting with a run of zeros).  masks:
    bool / uint8 tensor [N, H, W] (what FCNMaskHead.get_seg_masks(to_numpy=False) returns).
    Transitions are found and compacted on the device; ONE device->host copy of the run
    boundaries replaces the N full-size bitmap copies.  Returns a list of
    {'size': [H, W], 'counts': [int, ...]} (COCO's uncompressed RLE form)."""
    N, H, W = masks.shape
    if N == 0:
        return []
    flat = masks.to(torch.uint8).transpose(1, 2).reshape(N, H * W)       # column-major per mask
    prev = torch.cat([flat.new_zeros((N, 1)), flat[:, :-1]], 1)
    change = flat != prev                                                # run starts (first: a 1-run)
    idx = torch.nonzero(change)                                          # sorted by (mask, position)
    per = torch.bincount(idx[:, 0], minlength=N)
    pos = idx[:, 1].cpu().tolist()
    per = per.cpu().tolist()
    out, o = [], 0
    for n in range(N):
        b = [0] + pos[o:o + per[n]] + [H * W]
        o += per[n]
        out.append({'size': [H, W], 'counts': [b[i + 1] - b[i] for i in range(len(b) - 1)]})
    return out
