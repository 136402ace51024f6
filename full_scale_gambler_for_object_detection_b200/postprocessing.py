"""``detector_postprocess`` with the reference's signature (detectron2/modeling/postprocessing.py:8-52) for the
box fields of an ``Instances`` (pred_boxes or proposal_boxes; masks and keypoints are outside this path).

``RetinaNetDensePath.inference(..., output_sizes=...)`` does the same work inside the NMS kernel's epilogue
(no extra launch); this stand-alone form serves any other caller that already holds an ``Instances``."""
from . import ops
from .structures import as_tensor, make_boxes


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    if results.has("pred_masks") or results.has("pred_keypoints"):
        raise NotImplementedError("masks / keypoints are not on the dense-detection path")
    fields = dict(results.get_fields())
    name = "pred_boxes" if "pred_boxes" in fields else "proposal_boxes"
    boxes, keep = ops.postprocess_boxes(as_tensor(fields[name]), results.image_size, output_height, output_width)
    fields[name] = make_boxes(boxes)
    out = type(results)((output_height, output_width), **fields)   # the caller's own Instances class comes back
    return out[keep]
