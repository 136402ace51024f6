"""``Box2BoxTransform`` (detectron2/modeling/box_regression.py:15-107) on the GPU library."""
import math

import torch

from . import ops

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)  # box_regression.py:8
STRICT = False  # True: keep the reference's host-synchronising asserts (box_regression.py:66,79)

__all__ = ["Box2BoxTransform"]


class Box2BoxTransform(object):
    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        """(n,4),(n,4) -> (n,4) (dx,dy,dw,dh); box_regression.py:34-67."""
        assert isinstance(src_boxes, torch.Tensor), type(src_boxes)
        assert isinstance(target_boxes, torch.Tensor), type(target_boxes)
        deltas = ops.get_deltas(src_boxes, target_boxes, self.weights)
        if STRICT:
            assert ((src_boxes[:, 2] - src_boxes[:, 0]) > 0).all().item(), \
                "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        """deltas (n,4k), boxes (n,4) -> (n,4k); box_regression.py:69-107."""
        if STRICT:
            assert torch.isfinite(deltas).all().item(), "Box regression deltas become infinite or NaN!"
        return ops.apply_deltas(deltas, boxes, self.weights, self.scale_clamp)
