"""The per-anchor method bodies of the reference's RetinaNet meta-architecture
(detectron2/modeling/meta_arch/retinanet.py), with the same names, arguments and return values:

    get_ground_truth(anchors, targets)           :309-368
    get_picky_ground_truth(anchors, targets)     :370-429   (fork)
    losses(gt_classes, gt_anchors_deltas, pred_class_logits, pred_anchor_deltas)   :201-248
    inference(box_cls, box_delta, anchors, image_sizes) / inference_single_image   :431-520

Backbone, head convolutions and the gambler U-Net are out of scope (cuDNN territory); this class is what a
maintainer binds those methods to.  All arithmetic runs in ``libfsg_dense.so``.
"""
from typing import List

import torch

from . import _lib, ops
from .box_regression import Box2BoxTransform
from .matcher import Matcher
from .structures import as_tensor, cat_tensors, make_boxes, make_instances


class _LevelsToFlat(torch.autograd.Function):
    """permute_to_N_HWA_K + cat over levels (retinanet.py:24-54) as transpose kernels, differentiable."""

    @staticmethod
    def forward(ctx, K, *levels):
        ctx.K = K
        ctx.shapes = [tuple(x.shape[1:]) for x in levels]
        return ops.levels_to_flat([x.detach() for x in levels], K)

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(ops.flat_to_levels(g.contiguous(), ctx.shapes))


def levels_to_flat(levels, K):
    return _LevelsToFlat.apply(K, *levels)


def _as_f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


class _RetinaLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, deltas, gt_classes, gt_deltas, params):
        x = logits.detach().to(torch.float32).contiguous()
        d = deltas.detach().to(torch.float32).contiguous()
        stats = ops.loss_prepass(gt_classes, None, None, params.num_classes, 0.0)
        out = ops.loss_main(x, gt_classes, params, stats, pred_deltas=d, gt_deltas=gt_deltas)
        ctx.save_for_backward(out["grad_logits"], out["grad_deltas"])
        s = out["scalars"]
        return s[5].to(torch.float32), s[6].to(torch.float32)

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        ops.single_use(ctx)
        gl, gd = ctx.saved_tensors
        # loss_cls depends only on the logits and loss_box_reg only on the deltas, so the gradients saved
        # for unit coefficients just need scaling by the upstream scalars (on the device, no sync)
        ops.scale_(gl, g_cls)
        ops.scale_(gd, g_reg)
        return gl, gd, None, None, None


class _RetinaLossesLevels(torch.autograd.Function):
    """RetinaNet.losses on the head's native per-level layout: no permute/cat copy of the logits or deltas in
    either direction (fsg_loss_main_levels).  Inputs: L logit levels then L delta levels."""

    @staticmethod
    def forward(ctx, gt_classes, gt_deltas, params, L, *levels):
        xs = [t.detach() for t in levels[:L]]
        ds = [t.detach() for t in levels[L:]]
        stats = ops.loss_prepass(gt_classes, None, None, params.num_classes, 0.0)
        out = ops.loss_main_levels(xs, gt_classes, params, stats, delta_levels=ds, gt_deltas=gt_deltas)
        ctx.L = L
        ctx.save_for_backward(*(out["grad_logits"] + out["grad_deltas"]))
        s = out["scalars"]
        return s[5].to(torch.float32), s[6].to(torch.float32)

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        ops.single_use(ctx)
        saved = ctx.saved_tensors
        for t in saved[:ctx.L]:
            ops.scale_(t, g_cls)
        for t in saved[ctx.L:]:
            ops.scale_(t, g_reg)
        return (None, None, None, None) + tuple(saved)


class RetinaNetDensePath:
    def __init__(self, num_classes=80, focal_loss_alpha=0.25, focal_loss_gamma=2.0, smooth_l1_loss_beta=0.1,
                 score_threshold=0.05, topk_candidates=1000, nms_threshold=0.5, max_detections_per_image=100,
                 iou_thresholds=(0.4, 0.5), iou_labels=(0, -1, 1), bbox_reg_weights=(1.0, 1.0, 1.0, 1.0),
                 native_layout=True):
        # retinanet.py:69-100
        self.native_layout = native_layout   # losses() reads the (N, A*K, H, W) head outputs in place
        self.num_classes = num_classes
        self.focal_loss_alpha = focal_loss_alpha
        self.focal_loss_gamma = focal_loss_gamma
        self.smooth_l1_loss_beta = smooth_l1_loss_beta
        self.score_threshold = score_threshold
        self.topk_candidates = topk_candidates
        self.nms_threshold = nms_threshold
        self.max_detections_per_image = max_detections_per_image
        self.box2box_transform = Box2BoxTransform(weights=tuple(bbox_reg_weights))
        self.matcher = Matcher(list(iou_thresholds), list(iou_labels), allow_low_quality_matches=True)
        self.picky_matcher = Matcher([0.4, 0.9], list(iou_labels), allow_low_quality_matches=True)
        self._match_cache = None

    @classmethod
    def from_config(cfg_cls, cfg):
        r = cfg.MODEL.RETINANET
        return cfg_cls(r.NUM_CLASSES, r.FOCAL_LOSS_ALPHA, r.FOCAL_LOSS_GAMMA, r.SMOOTH_L1_LOSS_BETA,
                       r.SCORE_THRESH_TEST, r.TOPK_CANDIDATES_TEST, r.NMS_THRESH_TEST,
                       cfg.TEST.DETECTIONS_PER_IMAGE, r.IOU_THRESHOLDS, r.IOU_LABELS,
                       cfg.MODEL.RPN.BBOX_REG_WEIGHTS)

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _anchor_tensor(anchors):
        """list[list[Boxes]] (per image, per level) -> (R,4) if every image carries the SAME anchor objects or
        storage (what ``DefaultAnchorGenerator`` of this package hands out), else (N,R,4).  The reference deep-copies
        one anchor set per image (anchor_generator.py:188): equal contents in different storage -- telling that apart
        would take a device-side compare and a host sync per image, so copies are simply stacked (17 MB at config 2)."""
        first_list = anchors[0]
        if all(a is first_list for a in anchors[1:]):
            return cat_tensors(first_list).contiguous()
        per_image = [cat_tensors(a) for a in anchors]  # retinanet.py:339
        first = per_image[0]
        if all(t.shape == first.shape and t.data_ptr() == first.data_ptr() for t in per_image[1:]):
            return first.contiguous()
        return torch.stack(per_image).contiguous()

    def _match(self, anchors, targets):
        """One K1 run yields gt_classes, gt_anchors_deltas AND the picky mask; the reference calls get_ground_truth and
        get_picky_ground_truth back to back on the same (anchors, targets) (train_net.py:1142-1150), so the result is
        kept for the second call.  The key holds storage pointers and tensor versions: an in-place edit of the ground
        truth or the anchors between the two calls invalidates it."""
        a = self._anchor_tensor(anchors)
        gtb = [as_tensor(t.gt_boxes) for t in targets]
        gtc = [t.gt_classes for t in targets]
        key = ((a.data_ptr(), a._version, tuple(a.shape)),
               tuple((b.data_ptr(), b._version, c.data_ptr(), c._version, b.shape[0]) for b, c in zip(gtb, gtc)))
        if self._match_cache is not None and self._match_cache[0] == key:
            return self._match_cache[1]
        gt = ops.PackedGT.from_lists(gtb, gtc, a.device)
        out = ops.match_anchors(a, gt, self.num_classes, self.matcher.thresholds[1:-1], self.matcher.labels,
                                self.picky_matcher.thresholds[1:-1], self.picky_matcher.labels,
                                self.box2box_transform.weights, want=("gt_classes", "gt_deltas", "mask"))
        self._match_cache = (key, out)
        return out

    @torch.no_grad()
    def get_ground_truth(self, anchors, targets):
        """-> gt_classes (N,R) int64 in {-1, 0..K-1, K}, gt_anchors_deltas (N,R,4)."""
        out = self._match(anchors, targets)
        return out["gt_classes"], out["gt_deltas"]

    @torch.no_grad()
    def get_picky_ground_truth(self, anchors, targets):
        """-> mask (N,R) int64: 1 where the [0.4,0.9] matcher labels the anchor positive (IoU >= 0.9 or a
        ground truth's best anchor), else 0; all K for an image without ground truth (retinanet.py:425)."""
        return self._match(anchors, targets)["mask"]

    def losses(self, gt_classes, gt_anchors_deltas, pred_class_logits, pred_anchor_deltas):
        """-> {"loss_cls", "loss_box_reg"} (differentiable wrt the per-level head outputs)."""
        params = ops.make_loss_params(self.num_classes, self.focal_loss_alpha, self.focal_loss_gamma,
                                      self.smooth_l1_loss_beta, 0.0, 1.0, "focal", _lib.NORM_NONE, 1.0, 1.0, 0.0,
                                      self.box2box_transform.weights)
        gt_classes = gt_classes.contiguous()
        gt_anchors_deltas = gt_anchors_deltas.to(torch.float32).contiguous()
        if self.native_layout:
            levels = [_as_f32c(t) for t in list(pred_class_logits) + list(pred_anchor_deltas)]
            loss_cls, loss_box_reg = _RetinaLossesLevels.apply(gt_classes, gt_anchors_deltas, params,
                                                               len(pred_class_logits), *levels)
        else:   # the reference's own data flow: permute + cat, then the (N, R, K) kernel
            x = levels_to_flat(list(pred_class_logits), self.num_classes)
            d = levels_to_flat(list(pred_anchor_deltas), 4)
            loss_cls, loss_box_reg = _RetinaLosses.apply(x, d, gt_classes, gt_anchors_deltas, params)
        return {"loss_cls": loss_cls, "loss_box_reg": loss_box_reg}

    # ---------------------------------------------------------------------------------------------
    def _detect(self, logits_flat, deltas_flat, anchor_tensor, level_offsets, want_candidates=False,
                postprocess=None):
        return ops.detect(logits_flat, deltas_flat, anchor_tensor, level_offsets, self.score_threshold,
                          self.topk_candidates, self.nms_threshold, self.max_detections_per_image,
                          self.box2box_transform.weights, self.box2box_transform.scale_clamp, want_candidates,
                          postprocess)

    @staticmethod
    def _to_instances(res, n, image_size, count=None):
        c = int(res["count"][n].item()) if count is None else count
        return make_instances(tuple(image_size), pred_boxes=make_boxes(res["boxes"][n, :c]),
                              scores=res["scores"][n, :c], pred_classes=res["classes"][n, :c])

    @staticmethod
    def _to_instances_batch(res, image_sizes):
        """One device->host read of the N detection counts for the whole batch (the reference's per-image
        variable-length Instances need the lengths on the host), not one ``.item()`` per image."""
        counts = res["count"].tolist()
        return [RetinaNetDensePath._to_instances(res, n, image_sizes[n], counts[n]) for n in range(len(counts))]

    @torch.no_grad()
    def inference(self, box_cls, box_delta, anchors, images, *, output_sizes=None):
        """retinanet.py:431-458.  box_cls/box_delta: list over levels of (N, A*K|A*4, H, W); anchors:
        list[list[Boxes]]; images: an ``ImageList`` (anything with ``.image_sizes``) or a list of (h, w).  All images
        go through one batched launch sequence.
        output_sizes (extension, keyword only): list of (height, width) -- ``detector_postprocess``
        (postprocessing.py:8-52, what RetinaNet.forward applies to every result, retinanet.py:150-157) is then fused
        into the NMS epilogue and the returned Instances are at the output resolution."""
        image_sizes = list(images.image_sizes) if hasattr(images, "image_sizes") else list(images)
        assert len(anchors) == len(image_sizes)
        post = None
        if output_sizes is not None:
            assert len(output_sizes) == len(image_sizes)
            post = ops.postprocess_rows(image_sizes, output_sizes, box_cls[0].device)
            image_sizes = [tuple(s) for s in output_sizes]
        a = self._anchor_tensor(anchors)
        if self.native_layout:   # the (N, A*K, H, W) / (N, A*4, H, W) head outputs are read in place
            res = ops.detect_levels([t.detach() for t in box_cls], [t.detach() for t in box_delta], a,
                                    self.num_classes, self.score_threshold, self.topk_candidates, self.nms_threshold,
                                    self.max_detections_per_image, self.box2box_transform.weights,
                                    self.box2box_transform.scale_clamp, postprocess=post)
        else:                    # the reference's data flow: permute + cat per level (retinanet.py:444-447)
            x = ops.levels_to_flat([t.detach() for t in box_cls], self.num_classes)
            d = ops.levels_to_flat([t.detach() for t in box_delta], 4)
            offs = [0]
            for lvl in anchors[0]:
                offs.append(offs[-1] + len(lvl))
            res = self._detect(x, d, a, offs, postprocess=post)
        return self._to_instances_batch(res, image_sizes)

    @torch.no_grad()
    def inference_single_image(self, box_cls, box_delta, anchors, image_size):
        """box_cls: list over levels of (HWA, K); box_delta: (HWA, 4); anchors: list[Boxes]."""
        x = torch.cat([t.reshape(-1, self.num_classes) for t in box_cls]).to(torch.float32).contiguous()[None]
        d = torch.cat([t.reshape(-1, 4) for t in box_delta]).to(torch.float32).contiguous()[None]
        offs = [0]
        for a in anchors:
            offs.append(offs[-1] + len(a))
        res = self._detect(x, d, cat_tensors(list(anchors)).contiguous(), offs)
        return self._to_instances(res, 0, image_size)
