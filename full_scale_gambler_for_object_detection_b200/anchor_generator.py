"""Anchor grids with ``DefaultAnchorGenerator`` semantics (detectron2/modeling/anchor_generator.py:121-168).

Host-side closed form (torch ops on whatever device is asked for): cell anchors per level from
(sizes, aspect_ratios), shifted over the (H, W) grid with the level stride; order (h, w, a), XYXY fp32.
Generating them inside K1/K3 is the "next" row (f1) of SURVEY.md section 8.
"""
import math

import torch

RETINANET_STRIDES = (8, 16, 32, 64, 128)
# ImbalanceDetection/configs/Base-RetinaNet.yaml:8 -- three octave scales per level, aspect ratio 1.0 (A = 3)
RETINANET_SIZES = tuple(tuple(x * 2 ** (i / 3.0) for i in range(3)) for x in (32, 64, 128, 256, 512))


def generate_cell_anchors(sizes, aspect_ratios):
    """anchor_generator.py:131-168."""
    out = []
    for size in sizes:
        area = size ** 2.0
        for ar in aspect_ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def grid_anchors(grid_sizes, strides, sizes, aspect_ratios, device="cpu"):
    """-> list[(H*W*A, 4)] per level (anchor_generator.py:121-129)."""
    res = []
    for (H, W), stride, sz, ar in zip(grid_sizes, strides, sizes, aspect_ratios):
        base = generate_cell_anchors(sz, ar).to(device)
        sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32, device=device)
        sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32, device=device)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        res.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return res


def retinanet_grid_sizes(height, width, size_divisibility=32):
    """FPN P3..P7 grids for a padded image (fpn.py:100,189-190): P3..P5 = ceil(H/stride); P6, P7 are
    stride-2 3x3 convs with padding 1 -> ceil(in/2)."""
    H = (height + size_divisibility - 1) // size_divisibility * size_divisibility
    W = (width + size_divisibility - 1) // size_divisibility * size_divisibility
    g = [((H + s - 1) // s, (W + s - 1) // s) for s in (8, 16, 32)]
    for _ in range(2):
        h, w = g[-1]
        g.append(((h + 1) // 2, (w + 1) // 2))
    return g


def retinanet_anchors(height, width, device="cpu", sizes=RETINANET_SIZES, aspect_ratios=((1.0,),) * 5):
    """-> (anchors (R,4), level_offsets [L+1], grid_sizes) for the reference's RetinaNet+gambler configs."""
    grids = retinanet_grid_sizes(height, width)
    per_level = grid_anchors(grids, RETINANET_STRIDES, sizes, aspect_ratios, device)
    offs = [0]
    for a in per_level:
        offs.append(offs[-1] + a.shape[0])
    return torch.cat(per_level).contiguous(), offs, grids
