"""``nms`` / ``batched_nms`` with the reference's signatures (detectron2/layers/nms.py:6,9-26).

torchvision semantics: stable score-descending greedy NMS, suppress when IoU > threshold; ``batched_nms``
is the per-class (un-offset) algorithm; the result is sorted by score descending.  The whole thing runs in
one CUDA kernel; the only host synchronisation is reading the number of kept boxes to size the result
(the reference has the same variable-length return)."""

from . import ops


def nms(boxes, scores, iou_threshold):
    """boxes (n,4), scores (n) -> int64 keep indices, score-descending."""
    keep, num = ops.nms_raw(boxes, scores, None, iou_threshold)
    return keep[: int(num.item())]


def batched_nms(boxes, scores, idxs, iou_threshold):
    """Per-class NMS; idxs (n) int64 class ids in [0, 2^19)."""
    assert boxes.shape[-1] == 4
    keep, num = ops.nms_raw(boxes, scores, idxs, iou_threshold)
    return keep[: int(num.item())]
