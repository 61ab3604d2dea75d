"""Test-time post-processing kernels (csrc/infer.cu) against the oracle's restatement of the
reference: mask paste (fcn_mask_head.py:151-308), FOA offset fusion + decode
(offset_head_expand_feature.py:346-448, delta_xy_offset_coder.py:67-88), device RLE packing
(apis/test.py:53-74)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dets(n, size, seed):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * size
    wh = torch.exp(torch.rand(n, 2, generator=g) * 2.2) * 9
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    b[0] = torch.tensor([-7.3, -2.2, 31.7, 40.1])                  # sticks out at the top-left
    b[1] = torch.tensor([size - 20.5, size - 33.25, size + 12.0, size + 3.0])   # bottom-right
    b[2] = torch.tensor([50.0, 60.0, 50.0, 90.0])                  # zero width (inf grid -> 0)
    b[3] = torch.tensor([10.25, 10.75, 11.0, 12.5])                # sub-pixel box
    return torch.cat([b, torch.rand(n, 1, generator=g)], 1), g


@pytest.mark.parametrize('size,n', [(256, 37), (1024, 64)])
def test_paste_masks_vs_oracle(size, n):
    from bonai_b200.ops import paste_masks
    from oracle import loft_cpu as O
    dets, g = _dets(n, size, 0)
    logits = torch.randn(n, 1, 28, 28, generator=g) * 3
    want = O.paste_masks(logits, dets, size, size, 0.5)             # bool [n, H, W]
    got = paste_masks(logits[:, 0].cuda(), dets.cuda(), size, size, 0.5).cpu()
    assert got.dtype == torch.bool and got.shape == want.shape
    # every pixel equal except where the interpolated probability is within float round-off of
    # the threshold (CPU and GPU sigmoid / FMA contraction differ in the last ulp)
    diff = int((got != want).sum())
    assert diff <= max(2, int(1e-6 * want.numel())), diff
    assert int(got.sum()) > 0
    # a strided view of a fused NHWC head output [n, 4, 28, 28] (what the mask head hands over)
    fused = torch.zeros(n, 28, 28, 4)
    fused[..., 0] = logits[:, 0]
    view = fused.cuda()[..., 0]
    got2 = paste_masks(view, dets.cuda(), size, size, 0.5).cpu()
    assert torch.equal(got2, got)


def test_offset_fusion_decode_vs_oracle():
    from bonai_b200.ops import offset_fusion_decode
    from oracle import loft_cpu as O
    n = 333
    dets, g = _dets(n, 1024, 1)
    pred = torch.randn(4 * n, 2, generator=g) * 2
    pred[5] = 0.0                                                   # polarity of an exact zero
    want = O.delta2offset(dets, O.offset_fusion_max(pred), (0.5, 0.5), max_shape=[1024, 1024])
    fused = torch.zeros(4 * n, 4)
    fused[:, :2] = pred
    got = offset_fusion_decode(fused.cuda()[:, :2], dets.cuda(), (0.5, 0.5), [1024, 1024]).cpu()
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)


def test_rle_of_pasted_masks():
    from bonai_b200.core import encode_mask_results
    from bonai_b200.ops import paste_masks
    dets, g = _dets(9, 128, 2)
    logits = torch.randn(9, 28, 28, generator=g) * 3
    m = paste_masks(logits.cuda(), dets.cuda(), 128, 128, 0.5)
    rle = encode_mask_results(m)
    host = m.cpu().numpy()
    for i, r in enumerate(rle):
        assert r['size'] == [128, 128] and sum(r['counts']) == 128 * 128
        flat = host[i].T.reshape(-1)                                # column-major
        dec = np.zeros_like(flat)
        o, v = 0, 0
        for c in r['counts']:
            dec[o:o + c] = v
            o += c
            v ^= 1
        assert np.array_equal(dec.astype(bool), flat)


def test_batched_device_inference_equals_per_tile_api():
    """model.simple_test_batch (dense work batched over tiles, results on the device) gives what
    the reference-format simple_test gives tile by tile."""
    import os
    from bonai_b200 import Config
    from bonai_b200.models import build_detector
    from oracle import loft_cpu as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, 'configs', 'loft', 'loft_foa_r50_fpn_2x_b200.py'))
    model = build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    model.load_state_dict(O.randomize_bn(O.init_params(0), 0))
    model.eval()
    S = 256
    g = torch.Generator().manual_seed(0)
    imgs = torch.randn(2, 3, S, S, generator=g).cuda()
    meta = dict(img_shape=(S, S, 3), ori_shape=(S, S, 3), pad_shape=(S, S, 3), scale_factor=1.0,
                flip=False)
    batch = model.simple_test_batch(imgs, [meta, meta])
    for i in range(2):
        bbox_r, segm_r, off_r = model.simple_test(imgs[i:i + 1], [meta])
        dets, labels, masks, offs = batch[i]
        want = torch.from_numpy(bbox_r[0])
        assert dets.shape == want.shape
        # same proposals / scores up to the batch-size dependence of GEMM tiling (fp32 sums in a
        # different order): boxes to 1e-3 px, scores to 1e-5
        assert torch.allclose(dets.cpu(), want, rtol=1e-4, atol=2e-3)
        assert torch.allclose(offs.cpu(), torch.from_numpy(np.asarray(off_r)), rtol=1e-3, atol=2e-2)
        host = masks.cpu().numpy()
        mism = sum(int((host[j] != segm_r[0][j]).sum()) for j in range(len(segm_r[0])))
        assert mism <= 1e-4 * host.size, mism
