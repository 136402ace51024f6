"""The two-stage callers of the same kernels, with the reference's signatures:

    find_top_rpn_proposals            proposal_generator/rpn_outputs.py:52-151
    rpn_ground_truth                  RPNOutputs._get_ground_truth, rpn_outputs.py:250-295
    subsample_labels                  modeling/sampling.py:7-50
    label_proposals                   ROIHeads.label_and_sample_proposals, roi_heads/roi_heads.py:233-246 (matching part)
    fast_rcnn_inference[_single_image] roi_heads/fast_rcnn.py:40-118

Everything numeric runs in ``libfsg_dense.so``; torch is used for allocation, random permutations (the
reference's own ``torch.randperm``) and the final variable-length indexing.
"""
import torch

from . import ops
from .nms import batched_nms
from .structures import as_tensor, make_boxes, make_instances


def _image_sizes(images):
    return list(images.image_sizes) if hasattr(images, "image_sizes") else list(images)


def find_top_rpn_proposals(proposals, pred_objectness_logits, images, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_side_len, training):
    """proposals list[(N, Hi*Wi*A, 4)], pred_objectness_logits list[(N, Hi*Wi*A)], images: ImageList or a list of
    (h, w) -> list[Instances] with ``proposal_boxes`` / ``objectness_logits``.  The whole batch runs in two kernel
    launches (the reference loops over images with a host sync per image, rpn_outputs.py:123-131).  ``training``
    only matters to the reference's NaN filter (:106-116), which is a host-synchronising debug aid not kept here."""
    sizes = _image_sizes(images)
    res = ops.rpn_proposals(proposals, pred_objectness_logits, sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                            min_box_side_len)
    counts = res["count"].tolist()
    out = []
    for n, size in enumerate(sizes):
        out.append(make_instances(tuple(size), proposal_boxes=make_boxes(res["boxes"][n, :counts[n]]),
                                  objectness_logits=res["logits"][n, :counts[n]]))
    return out


def rpn_ground_truth(anchors, gt_boxes, iou_thresholds=(0.3, 0.7), iou_labels=(0, -1, 1),
                     box_weights=(1.0, 1.0, 1.0, 1.0), boundary_threshold=-1, image_sizes=None):
    """RPNOutputs._get_ground_truth.  anchors: (R,4) tensor shared by the images (or (N,R,4)); gt_boxes: list of
    Boxes / (M_i,4) tensors.  -> (gt_objectness_logits list[(R) int8 in {-1,0,1}], gt_anchor_deltas list[(R,4)]).
    Matcher(allow_low_quality_matches=True) as rpn.py builds it; images without GT: labels 0, deltas 0."""
    boxes = [as_tensor(b) for b in gt_boxes]
    dev = anchors.device
    gt = ops.PackedGT.from_lists(boxes, [torch.zeros(b.shape[0], dtype=torch.int64) for b in boxes], dev)
    out = ops.match_anchors(anchors, gt, 1, iou_thresholds, iou_labels, None, None, box_weights,
                            want=("match_labels", "gt_deltas"))
    labels, deltas = out["match_labels"], out["gt_deltas"]
    if boundary_threshold >= 0:   # legacy option, off by default (rpn_outputs.py:276-280)
        a = anchors if anchors.dim() == 3 else anchors[None].expand(len(boxes), -1, -1)
        for n, (h, w) in enumerate(image_sizes):
            t = boundary_threshold
            inside = (a[n, :, 0] >= -t) & (a[n, :, 1] >= -t) & (a[n, :, 2] < w + t) & (a[n, :, 3] < h + t)
            labels[n][~inside] = -1
    return [labels[n] for n in range(len(boxes))], [deltas[n] for n in range(len(boxes))]


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    """modeling/sampling.py:7-50 -> (pos_idx, neg_idx): random subsets (torch.randperm, like the reference)."""
    pos = torch.nonzero((labels != -1) & (labels != bg_label)).squeeze(1)
    neg = torch.nonzero(labels == bg_label).squeeze(1)
    num_pos = min(pos.numel(), int(num_samples * positive_fraction))
    num_neg = min(neg.numel(), num_samples - num_pos)
    return (pos[torch.randperm(pos.numel(), device=pos.device)[:num_pos]],
            neg[torch.randperm(neg.numel(), device=neg.device)[:num_neg]])


def label_proposals(proposal_boxes, gt_boxes, gt_classes, num_classes, iou_thresholds=(0.5,), iou_labels=(0, 1)):
    """The matching part of ROIHeads.label_and_sample_proposals for one image: pairwise_iou + Matcher (no
    low-quality pass) + the relabelling of _sample_proposals (roi_heads.py:178-186), without the (M, P) matrix.
    -> (matched_idxs (P) int64, matched_labels (P) int8, gt_classes (P) int64 with background = num_classes and
    ignored = -1)."""
    p, g = as_tensor(proposal_boxes), as_tensor(gt_boxes)
    gt = ops.PackedGT.from_lists([g], [gt_classes], p.device)
    out = ops.match_anchors(p.contiguous(), gt, num_classes, iou_thresholds, iou_labels, None, None,
                            want=("matches", "match_labels", "gt_classes"), allow_low_quality_matches=False)
    return out["matches"][0], out["match_labels"][0], out["gt_classes"][0]


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """fast_rcnn.py:76-118: boxes (R, K*4) or (R, 4), scores (R, K+1) -> (Instances, kept row indices)."""
    cb, cs, cc, cr = ops.score_filter(boxes, scores, image_shape, score_thresh)
    keep = batched_nms(cb, cs, cc, nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    result = make_instances(tuple(image_shape), pred_boxes=make_boxes(cb[keep]), scores=cs[keep],
                            pred_classes=cc[keep])
    return result, cr[keep]


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, topk_per_image):
    """fast_rcnn.py:40-73."""
    per_image = [fast_rcnn_inference_single_image(b, s, shp, score_thresh, nms_thresh, topk_per_image)
                 for s, b, shp in zip(scores, boxes, image_shapes)]
    return tuple(list(x) for x in zip(*per_image))
