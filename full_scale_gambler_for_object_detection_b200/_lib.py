"""ctypes binding of ``libfsg_dense.so`` (C ABI: ``include/fsg_dense.h``).

The shared library is the product; this module only marshals ``tensor.data_ptr()`` values, sizes and the
current CUDA stream across the boundary.  There is **no CPU fallback**: if the library is missing or a
tensor is not a CUDA tensor the call raises.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSG_DENSE_LIB") or os.path.join(_HERE, "libfsg_dense.so")  # override: kernel experiments
ABI_VERSION = 4
STATS_HEADER = 2
SCALARS_HEADER = 10

c_i32, c_i64, c_f32, c_f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double
c_ptr, c_size = ctypes.c_void_p, ctypes.c_size_t


class LossParams(ctypes.Structure):
    """``struct fsg_loss_params``."""

    _fields_ = [
        ("num_classes", c_i32),
        ("gambler_mode", c_i32),
        ("norm_mode", c_i32),
        ("reserved", c_i32),
        ("focal_alpha", c_f32),
        ("focal_gamma", c_f32),
        ("smooth_l1_beta", c_f32),
        ("temperature", c_f32),
        ("gambler_gamma", c_f32),
        ("c_cls", c_f32),
        ("c_reg", c_f32),
        ("c_gam", c_f32),
        ("box_weights", c_f32 * 4),
    ]


class HeadLevel(ctypes.Structure):
    """``struct fsg_head_level``."""

    _fields_ = [("logits", c_ptr), ("grad_logits", c_ptr), ("pred_deltas", c_ptr), ("grad_deltas", c_ptr),
                ("bets", c_ptr), ("per_anchor_loss", c_ptr), ("H", c_i32), ("W", c_i32)]


class BetLevels(ctypes.Structure):
    """``struct fsg_bet_levels``."""

    _fields_ = [("bets", c_ptr * 8), ("H", c_i32 * 8), ("W", c_i32 * 8), ("num_levels", c_i32), ("A", c_i32)]


class PostLevel(ctypes.Structure):
    """``struct fsg_post_level``."""

    _fields_ = [("bets", c_ptr), ("per_anchor_loss", c_ptr), ("grad_bets", c_ptr), ("H", c_i32), ("W", c_i32)]


class AnchorLevel(ctypes.Structure):
    """``struct fsg_anchor_level``."""

    _fields_ = [("H", c_i32), ("W", c_i32), ("stride", c_i32), ("A", c_i32), ("cell", (c_f32 * 4) * 16)]


class DetectLevel(ctypes.Structure):
    """``struct fsg_detect_level``."""

    _fields_ = [("logits", c_ptr), ("deltas", c_ptr), ("H", c_i32), ("W", c_i32)]


class PeerCtx(ctypes.Structure):
    """``struct fsg_peer_ctx``."""

    _fields_ = [("mailbox", ctypes.c_uint64 * 8), ("epoch", ctypes.c_uint64), ("error", ctypes.c_uint64),
                ("rank", c_i32), ("world", c_i32), ("timeout_cycles", ctypes.c_uint64)]


class StepIO(ctypes.Structure):
    """``struct fsg_step_io``."""

    _fields_ = [("logits", c_ptr), ("pred_deltas", c_ptr), ("bets", c_ptr), ("anchors", c_ptr),
                ("anchor_image_stride", c_i64), ("gt_boxes", c_ptr), ("gt_class_ids", c_ptr), ("gt_offsets", c_ptr),
                ("sum_M", c_i64), ("gt_classes", c_ptr), ("mask", c_ptr), ("matched_idx32", c_ptr), ("stats", c_ptr),
                ("scalars", c_ptr), ("grad_logits", c_ptr), ("grad_deltas", c_ptr), ("grad_bets", c_ptr),
                ("per_anchor_loss", c_ptr), ("weights_out", c_ptr)]


class StepLevelsIO(ctypes.Structure):
    """``struct fsg_step_levels_io``."""

    _fields_ = [("anchors", c_ptr), ("anchor_image_stride", c_i64), ("gt_boxes", c_ptr), ("gt_class_ids", c_ptr),
                ("gt_offsets", c_ptr), ("sum_M", c_i64), ("gt_classes", c_ptr), ("mask", c_ptr),
                ("matched_idx32", c_ptr), ("stats", c_ptr), ("scalars", c_ptr), ("weights_out", c_ptr)]


class MatchConfig(ctypes.Structure):
    """``struct fsg_match_config``."""

    _fields_ = [("thresholds", c_f32 * 4), ("picky_thresholds", c_f32 * 4), ("labels", ctypes.c_int8 * 8),
                ("picky_labels", ctypes.c_int8 * 8), ("num_thresholds", c_i32), ("num_picky_thresholds", c_i32),
                ("allow_low_quality_matches", c_i32), ("workspace_is_clean", c_i32)]


CLS_MODES = {"focal": 0, "sigmoid": 1}
NORM_NONE, NORM_IMAGE, NORM_BATCH = 0, 1, 2

# name -> (restype, argtypes); must list every symbol include/fsg_dense.h declares
PROTOTYPES = {
    "fsg_abi_version": (c_i32, []),
    "fsg_status_string": (ctypes.c_char_p, [c_i32]),
    "fsg_pairwise_iou": (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr]),
    "fsg_matcher": (c_i32, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    "fsg_match_workspace_bytes": (c_size, [c_i32, c_i64, c_i64]),
    "fsg_match_anchors": (
        c_i32,
        [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i32, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr,
         c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.POINTER(BetLevels), c_f32, c_ptr,
         ctypes.POINTER(PeerCtx), c_ptr, c_size, c_ptr],
    ),
    "fsg_match_anchors_ex": (
        c_i32,
        [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i32, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr,
         c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.POINTER(BetLevels), c_f32, c_ptr,
         ctypes.POINTER(PeerCtx), c_ptr, c_size, c_i32, c_ptr],
    ),
    "fsg_match_gt_max_offset": (c_size, [c_i32, c_i64, c_i64]),
    "fsg_box2box_get_deltas": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "fsg_box2box_apply_deltas": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_f32, c_ptr, c_ptr]),
    "fsg_loss_prepass_workspace_bytes": (c_size, [c_i32, c_i64]),
    "fsg_loss_prepass": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i32, c_f32, c_ptr, c_ptr, c_size, c_ptr]),
    "fsg_loss_main_workspace_bytes": (c_size, [c_i32, c_i64, c_i32]),
    "fsg_loss_main": (
        c_i32,
        [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i64,
         ctypes.POINTER(LossParams), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr],
    ),
    "fsg_loss_main_levels_workspace_bytes": (c_size, [c_i32, ctypes.POINTER(HeadLevel), c_i32, c_i32]),
    "fsg_loss_main_levels": (
        c_i32,
        [ctypes.POINTER(HeadLevel), c_i32, c_i32, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
         c_i32, c_i64, ctypes.POINTER(LossParams), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr],
    ),
    "fsg_loss_post_levels": (c_i32, [ctypes.POINTER(PostLevel), c_i32, c_i32, c_ptr, c_i32, c_i64,
                                     ctypes.POINTER(LossParams), c_ptr, c_ptr, c_ptr]),
    "fsg_loss_post": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i64, ctypes.POINTER(LossParams), c_ptr, c_ptr, c_ptr, c_ptr]),
    "fsg_dense_step_workspace_bytes": (c_size, [c_i32, c_i64, c_i32, c_i64]),
    "fsg_dense_step": (c_i32, [ctypes.POINTER(StepIO), c_i32, c_i64, ctypes.POINTER(MatchConfig),
                               ctypes.POINTER(LossParams), ctypes.POINTER(PeerCtx), c_ptr, c_size, c_ptr]),
    "fsg_dense_step_levels_workspace_bytes": (c_size, [c_i32, ctypes.POINTER(HeadLevel), c_i32, c_i32, c_i64]),
    "fsg_dense_step_levels": (c_i32, [ctypes.POINTER(StepLevelsIO), ctypes.POINTER(HeadLevel),
                                      ctypes.POINTER(PostLevel), c_i32, c_i32, c_i32, c_i64,
                                      ctypes.POINTER(MatchConfig), ctypes.POINTER(LossParams),
                                      ctypes.POINTER(PeerCtx), c_ptr, c_size, c_ptr]),
    "fsg_bet_stats_workspace_bytes": (c_size, []),
    "fsg_bet_stats": (c_i32, [c_ptr, ctypes.POINTER(BetLevels), c_ptr, c_i32, c_i64, ctypes.POINTER(LossParams), c_ptr,
                              c_ptr, c_ptr, c_size, c_ptr]),
    "fsg_scale_inplace": (c_i32, [c_ptr, c_i64, c_ptr, c_f32, c_ptr]),
    "fsg_nms_workspace_bytes": (c_size, [c_i64]),
    "fsg_nms": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_f64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "fsg_detect_workspace_bytes": (c_size, [c_i32, c_i64, c_i32, c_i32, c_i32]),
    "fsg_detect_status_offset": (c_size, [c_i32, c_i32, c_i32]),
    "fsg_detect": (
        c_i32,
        [c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_i64, c_i32, c_ptr, c_i32, c_f32, c_i32, c_f64, c_i32, c_ptr, c_f32,
         c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr],
    ),
    "fsg_detect_levels": (
        c_i32,
        [ctypes.POINTER(DetectLevel), c_i32, c_i32, c_i32, c_ptr, c_i64, c_i32, c_i64, c_f32, c_i32, c_f64, c_i32,
         c_ptr, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr],
    ),
    "fsg_postprocess_boxes": (c_i32, [c_ptr, c_i64, c_f32, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr]),
    "fsg_grid_anchors": (c_i32, [ctypes.POINTER(AnchorLevel), c_i32, c_ptr, c_i64, c_ptr]),
    "fsg_rpn_proposals_workspace_bytes": (c_size, [c_i32, c_ptr, c_i32, c_i32, c_i32]),
    "fsg_rpn_proposals": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_i32, c_i32, c_f64, c_f32, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "fsg_score_filter": (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_i32, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr,
                                 c_ptr, c_ptr]),
    "fsg_anchor_maps": (c_i32, [c_ptr, c_ptr, c_i32, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "fsg_permute_level": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i64, c_i64, c_i64, c_i32, c_ptr]),
}

_LIB = None
LAUNCHES = 0  # kernels enqueued through this binding (bench.py reports it as gpu_launches)
_CALL_DEVICE = None  # device of the tensor arguments marshalled for the call being assembled (see ptr / stream)


def build(verbose=False):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = [os.path.join(_HERE, "csrc", "build.sh")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libfsg_dense.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def lib():
    """The loaded library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libfsg_dense.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or csrc/build.sh. There is no CPU fallback for this path." % LIB_PATH
            )
        cdll = ctypes.CDLL(LIB_PATH)
        L = _Guarded()
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(cdll, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
            setattr(L, name, _guard(fn) if (args and args[-1] is c_ptr and res is c_i32) else fn)
        if L.fsg_abi_version() != ABI_VERSION:
            raise RuntimeError("libfsg_dense.so ABI %d != binding ABI %d" % (L.fsg_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


class _Guarded:
    """Namespace of the library's entry points; the launching ones run under a device guard."""


def _guard(fn):
    """The library never changes the current device (include/fsg_dense.h) and launches on the stream it is given, so
    the binding makes the tensors' device current for the duration of the call when it is not already: tensors on
    cuda:1 with cuda:0 current would otherwise be launched on a stream of the wrong device.  ``ptr`` records the
    device of every tensor argument (and refuses a mix), ``stream`` picks that device's current stream."""

    def call(*args):
        global _CALL_DEVICE
        dev, _CALL_DEVICE = _CALL_DEVICE, None
        if dev is None or dev == torch.cuda.current_device():
            return fn(*args)
        with torch.cuda.device(dev):
            return fn(*args)

    call.__name__ = fn.__name__
    return call


def check(status):
    if status != 0:
        msg = lib().fsg_status_string(status).decode()
        raise RuntimeError("fsg_dense: %s (status %d)" % (msg, status))


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL).  All tensors of one call must live on one device."""
    global _CALL_DEVICE
    if t is None:
        return None
    if not t.is_cuda:
        _CALL_DEVICE = None
        raise RuntimeError("fsg_dense kernels take CUDA tensors only (no CPU path); got %s" % t.device)
    if not t.is_contiguous():
        _CALL_DEVICE = None
        raise RuntimeError("fsg_dense kernels take contiguous tensors")
    idx = t.device.index
    if _CALL_DEVICE is None:
        _CALL_DEVICE = idx
    elif _CALL_DEVICE != idx:
        was, _CALL_DEVICE = _CALL_DEVICE, None
        raise RuntimeError("fsg_dense: tensor arguments on different devices (cuda:%d and cuda:%d)" % (was, idx))
    return t.data_ptr()


def stream():
    """Current stream of the device the call's tensors live on (the current device when there are none)."""
    if _CALL_DEVICE is None:
        return torch.cuda.current_stream().cuda_stream
    return torch.cuda.current_stream(_CALL_DEVICE).cuda_stream


def host_f32(vals):
    return (c_f32 * len(vals))(*[float(v) for v in vals])


def host_i8(vals):
    return (ctypes.c_int8 * len(vals))(*[int(v) for v in vals])


def host_i64(vals):
    return (c_i64 * len(vals))(*[int(v) for v in vals])


def count_launches(n):
    global LAUNCHES
    LAUNCHES += n
