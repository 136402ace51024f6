"""B200-native (sm_100a) per-anchor dense-detection hot path of the Full-Scale-Gambler detector.

Drop-in callables with the reference's names and signatures, backed by hand-written CUDA kernels behind the
C ABI in ``include/fsg_dense.h`` (``libfsg_dense.so``).  There is no CPU fallback: every op raises if the
library is not built or a tensor is not on a CUDA device.

    structures     Boxes, Instances, pairwise_iou          (detectron2/structures)
    matcher        Matcher                                 (detectron2/modeling/matcher.py)
    box_regression Box2BoxTransform                        (detectron2/modeling/box_regression.py)
    nms            nms, batched_nms                        (detectron2/layers/nms.py)
    anchor_generator DefaultAnchorGenerator                (detectron2/modeling/anchor_generator.py)
    postprocessing detector_postprocess                    (detectron2/modeling/postprocessing.py)
    proposals      find_top_rpn_proposals, rpn_ground_truth, subsample_labels, label_proposals,
                   fast_rcnn_inference[_single_image]      (rpn_outputs.py, sampling.py, roi_heads.py, fast_rcnn.py)
    retinanet      RetinaNetDensePath                      (meta_arch/retinanet.py: GT, losses, inference)
    gambler        GamblerLoss, get_loss_upper_bound       (imbalancedetection/gambler_heads.py)
    fused          dense_train_step, DenseLossConfig       (the fused K1+K2 step)
    ops            one function per C-ABI entry point
"""
from . import _lib, ops  # noqa: F401
from .anchor_generator import DefaultAnchorGenerator  # noqa: F401
from .box_regression import Box2BoxTransform  # noqa: F401
from .fused import (DenseLossConfig, DenseStepPlan, DenseStepPlanLevels, StepResult,  # noqa: F401
                    dense_train_step, dense_train_step_levels)
from .gambler import GamblerLoss, get_loss_upper_bound  # noqa: F401
from .matcher import Matcher  # noqa: F401
from .nms import batched_nms, nms  # noqa: F401
from .postprocessing import detector_postprocess  # noqa: F401
from .proposals import (fast_rcnn_inference, fast_rcnn_inference_single_image,  # noqa: F401
                        find_top_rpn_proposals, label_proposals, rpn_ground_truth, subsample_labels)
from .retinanet import RetinaNetDensePath  # noqa: F401
from .structures import Boxes, Instances, pairwise_iou  # noqa: F401

__all__ = [
    "Boxes", "Instances", "pairwise_iou", "Matcher", "Box2BoxTransform", "nms", "batched_nms",
    "RetinaNetDensePath", "GamblerLoss", "get_loss_upper_bound", "dense_train_step", "DenseLossConfig",
    "StepResult", "DenseStepPlan", "DenseStepPlanLevels", "dense_train_step_levels", "ops", "DefaultAnchorGenerator", "detector_postprocess",
    "find_top_rpn_proposals", "rpn_ground_truth", "subsample_labels", "label_proposals", "fast_rcnn_inference",
    "fast_rcnn_inference_single_image",
]
