"""CPU restatement of the polygon -> bitmap step of LoadAnnotations(poly2mask=True)
(mmdet/datasets/pipelines/loading.py:301-326: ``maskUtils.decode(maskUtils.merge(
maskUtils.frPyObjects(mask_ann, img_h, img_w)))``).  TEST INFRASTRUCTURE ONLY: imported by tests/
(and nothing under bonai_b200/).

The algorithm lives in a third-party dependency that is NOT vendored in the reference tree and is
not installed in this image: pycocotools 2.0.x (requirements/runtime.txt pins no version;
mmdet 2.x was released against pycocotools 2.0.0-2.0.2), file common/maskApi.c, functions
rleFrPoly, rleFrBbox, rleMerge(intersect=0) and rleDecode, bound by _mask.pyx frPyObjects / merge /
decode.  This file restates that published algorithm in numpy float64 / int arithmetic (C `int`
casts are truncations, `double` expressions are evaluated without fused multiply-add, as gcc does
for x86-64).

PARITY UNPINNED: there is no pycocotools here to generate golden bitmaps from and the reference has
no test for this step, so the pin is the algorithm text plus hand-checkable known answers
(tests/test_pipeline.py: axis-aligned rectangles, a triangle, clipping at the image border).
"""
import numpy as np


def _up(c):
    """(int)(scale * c + .5) with scale = 5 (C truncation toward zero)."""
    return np.trunc(5.0 * np.asarray(c, dtype=np.float64) + 0.5).astype(np.int64)


def fr_poly_boundaries(xy, h, w):
    """rleFrPoly up to the sorted run boundaries a[j] = x*h + y (column-major), without the final
    h*w sentinel.  `xy` = flat [x0, y0, x1, y1, ...]."""
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    k = xy.shape[0]
    if k == 0:
        return np.zeros((0,), dtype=np.int64)
    x = _up(xy[:, 0])
    y = _up(xy[:, 1])
    x = np.concatenate([x, x[:1]])
    y = np.concatenate([y, y[:1]])
    us, vs = [], []
    for j in range(k):                      # dense boundary, one unit step along the longer axis
        xs, xe, ys, ye = int(x[j]), int(x[j + 1]), int(y[j]), int(y[j + 1])
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            d = np.arange(dx + 1, dtype=np.int64)
            t = dx - d if flip else d
            s = (ye - ys) / dx if dx > 0 else 0.0      # 0/0 in C: the point is never used
            us.append(t + xs)
            vs.append(np.trunc(ys + s * t.astype(np.float64) + 0.5).astype(np.int64))
        else:
            d = np.arange(dy + 1, dtype=np.int64)
            t = dy - d if flip else d
            s = (xe - xs) / dy
            vs.append(t + ys)
            us.append(np.trunc(xs + s * t.astype(np.float64) + 0.5).astype(np.int64))
    u = np.concatenate(us)
    v = np.concatenate(vs)
    # points along the y boundary, downsampled
    ch = np.nonzero(u[1:] != u[:-1])[0] + 1
    uj, up, vj, vp = u[ch], u[ch - 1], v[ch], v[ch - 1]
    xd = np.where(uj < up, uj, uj - 1).astype(np.float64)
    xd = (xd + 0.5) / 5.0 - 0.5
    keep = (np.floor(xd) == xd) & (xd >= 0) & (xd <= w - 1)
    yd = np.minimum(vj, vp).astype(np.float64)
    yd = (yd + 0.5) / 5.0 - 0.5
    yd = np.ceil(np.clip(yd, 0.0, float(h)))
    a = (xd[keep].astype(np.int64) * h + yd[keep].astype(np.int64))
    return np.sort(a)


def decode_boundaries(a, h, w):
    """rleFrPoly's run construction + rleDecode: pixel i (column-major) is 1 iff an odd number of
    boundaries are <= i (zero-length runs merge, which leaves the parity unchanged)."""
    cnt = np.zeros(h * w + 1, dtype=np.int64)
    a = a[a < h * w]
    np.add.at(cnt, a, 1)
    par = (np.cumsum(cnt[:h * w]) & 1).astype(np.uint8)
    return par.reshape(w, h).T.copy()          # column-major -> [h, w]


def _bbox_poly(bb):
    """rleFrBbox: [x, y, w, h] -> the 4-vertex polygon it rasterises."""
    xs, ys = float(bb[0]), float(bb[1])
    xe, ye = xs + float(bb[2]), ys + float(bb[3])
    return [xs, ys, xs, ye, xe, ye, xe, ys]


def poly2mask(mask_ann, h, w):
    """`_poly2mask` for list input (loading.py:314-318): frPyObjects -> merge (union) -> decode.
    As in _mask.pyx frPyObjects, a list whose FIRST element has exactly 4 numbers is a list of
    bounding boxes."""
    out = np.zeros((h, w), dtype=np.uint8)
    if len(mask_ann) == 0:
        return out
    as_bbox = len(mask_ann[0]) == 4
    for part in mask_ann:
        xy = _bbox_poly(part) if as_bbox else part
        out |= decode_boundaries(fr_poly_boundaries(xy, h, w), h, w)
    return out
