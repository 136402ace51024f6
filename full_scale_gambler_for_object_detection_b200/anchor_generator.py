"""Anchor grids with ``DefaultAnchorGenerator`` semantics (detectron2/modeling/anchor_generator.py:121-168).

``DefaultAnchorGenerator`` produces the grid on the device with one kernel launch (``fsg_grid_anchors``) and
caches it per grid shape; every image of a batch shares that one (R,4) tensor (1 MB at 800x1333: it stays
L2-resident across the images of a batch, which is why the matching / loss / decode kernels read it instead of
re-deriving (level, y, x, a) per anchor with integer divisions).  ``grid_anchors`` below is the host-side closed
form used to build synthetic inputs on the CPU.  Order (h, w, a), XYXY fp32.
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


class DefaultAnchorGenerator:
    """anchor_generator.py:53-191 with the reference's call shape: ``gen(features) -> list[list[Boxes]]``
    (#images x #levels).  ``sizes`` / ``aspect_ratios`` broadcast over levels like the reference (:83-90).
    The per-image deep copies of the reference (:188) are replaced by references to one cached set of
    ``Boxes`` (nothing on this path mutates anchors); ``flat(features)`` gives the (R,4) tensor + level offsets
    the fused entry points take."""

    def __init__(self, sizes, aspect_ratios, strides, device="cuda"):
        self.strides = [int(s) for s in strides]
        n = len(self.strides)
        sizes = [list(s) for s in sizes] * (n if len(sizes) == 1 else 1)
        aspect_ratios = [list(a) for a in aspect_ratios] * (n if len(aspect_ratios) == 1 else 1)
        assert n == len(sizes) and n == len(aspect_ratios)
        self.cell_anchors = [generate_cell_anchors(s, a) for s, a in zip(sizes, aspect_ratios)]
        self.device = device
        self._cache = {}

    @classmethod
    def from_config(cls, cfg, input_shape, device="cuda"):
        return cls(cfg.MODEL.ANCHOR_GENERATOR.SIZES, cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS,
                   [x.stride for x in input_shape], device)

    @property
    def num_cell_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @property
    def box_dim(self):
        return 4

    def flat_for_grids(self, grid_sizes):
        from . import ops
        key = tuple((int(h), int(w)) for h, w in grid_sizes)
        if key not in self._cache:
            self._cache[key] = ops.grid_anchors(key, self.strides, self.cell_anchors, self.device)
        return self._cache[key]

    def flat(self, features):
        return self.flat_for_grids([f.shape[-2:] for f in features])

    def grid_anchors(self, grid_sizes):
        flat, offs = self.flat_for_grids(grid_sizes)
        return [flat[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]

    def __call__(self, features):
        from .structures import make_boxes
        per_level = [make_boxes(t) for t in self.grid_anchors([f.shape[-2:] for f in features])]
        return [per_level for _ in range(len(features[0]))]
