"""``nms`` / ``batched_nms`` with the reference's signatures (detectron2/layers/nms.py:6,9-26).

torchvision semantics: stable score-descending greedy NMS, suppress when IoU > threshold; the result is sorted by
score descending.  ``batched_nms`` is the per-class (un-offset) algorithm -- torchvision's ``_batched_nms_vanilla`` and
the reference's own >= 40 000-box branch (layers/nms.py:20-26).  For smaller inputs the reference calls torchvision's
coordinate-offset form (boxes shifted by ``idxs * (max_coordinate + 1)``, one NMS over everything): the same
suppression decisions except where a pair's fp32 IoU sits within rounding of the threshold, because the shift changes
the rounding of the box areas.  The keep indices here are bit-exact against the per-class algorithm (and against the
reference on every committed fixture); a borderline pair may differ from the offset form.

The whole thing runs on the device; the only host synchronisation is reading the number of kept boxes to size the
result (the reference has the same variable-length return)."""

from . import ops


def _keep(keep, num):
    n = int(num.item())
    if n < 0:
        raise ValueError("batched_nms: class ids must lie in [0, 2^18) for inputs of up to %d boxes "
                         "(the shared-memory path packs them into the sort key)" % ops.NMS_SMEM_BOXES)
    return keep[:n]


def nms(boxes, scores, iou_threshold):
    """boxes (n,4), scores (n) -> int64 keep indices, score-descending."""
    return _keep(*ops.nms_raw(boxes, scores, None, iou_threshold))


def batched_nms(boxes, scores, idxs, iou_threshold):
    """Per-class NMS; idxs (n) int64 class ids in [0, 2^18) = [0, 262144) (checked on the device for n <= 8192, any
    int64 beyond that)."""
    assert boxes.shape[-1] == 4
    return _keep(*ops.nms_raw(boxes, scores, idxs, iou_threshold))
