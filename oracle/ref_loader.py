"""Container-only loader for the reference's own hot-path Python (TEST INFRASTRUCTURE).

Loads the handful of pure-Python/torch files of the reference that make up the
per-anchor dense-detection path *by file path* from ``/root/reference`` after
registering thin stand-ins for the packages that are not installed here
(fvcore, yacs, the compiled ``detectron2._C``).  Recipe: SURVEY.md App. C.

It exists to (1) pin ``oracle/dense_oracle.py`` against the real reference and
(2) generate the committed fixtures under ``tests/golden/`` (see
``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box:
nothing imported by ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this
module.  ``available()`` says whether the reference tree is present.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("FSG_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "detectron2", "modeling", "matcher.py"))


def _load(dotted, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    """cfg-node stand-in: attribute access over nested dicts."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(**gambler):
    """Minimal ``global_cfg`` carrying the keys the loss code reads
    (gambler_heads.py:201,213,282,304,308,561; imbalancedetection/config.py:10-76)."""
    g = dict(
        GAMBLER_TEMPERATURE=0.1,
        GAMBLER_OUTPUT="L_BAHW",
        IN_LAYERS=[64, 32, 16, 8, 4],
        GAMBLER_KAPPA=1.0,
        NUM_CLASSES=80,
        GAMBLER_IN_CHANNELS=240,
        GAMBLER_OUT_CHANNELS=3,
        BILINEAR_UPSAMPLING=True,
        GAMBLER_LOSS_MODE="focal",
        NORMALIZE=True,
        GAMBLER_GAMMA=1.0,
    )
    g.update(gambler)
    cfg = _AttrDict(
        MODEL=_AttrDict(
            DEVICE="cpu",
            ANCHOR_GENERATOR=_AttrDict(SIZES=[[32, 40.31747359663594, 50.79683366298238]]),
            GAMBLER_HEAD=_AttrDict(g),
            RETINANET=_AttrDict(
                NUM_CLASSES=g["NUM_CLASSES"],
                FOCAL_LOSS_ALPHA=0.25,
                FOCAL_LOSS_GAMMA=2.0,
                IN_FEATURES=["p3", "p4", "p5", "p6", "p7"],
            ),
        ),
        OUTPUT_DIR="/tmp",
    )
    return cfg


class _Storage:
    """EventStorage stand-in (gambler_heads.py:583-587 only calls put_scalar)."""

    def __init__(self):
        self.scalars = {}
        self.iter = 0

    def put_scalar(self, k, v):
        self.scalars[k] = v


_LOADED = None


def load_reference():
    """Return a namespace with the reference's hot-path symbols, executed from
    the reference's own source files."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    # ---- third-party arithmetic not under /root/reference: fvcore.nn.
    # The focal formula is restated in-tree (retinanet.py:283-307, gambler_heads.py:104-128);
    # smooth_l1_loss follows fvcore's published semantics.
    import torch.nn.functional as F

    def sigmoid_focal_loss(inputs, targets, alpha: float = -1, gamma: float = 2, reduction: str = "none"):
        p = torch.sigmoid(inputs)
        ce_loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
        p_t = p * targets + (1 - p) * (1 - targets)
        loss = ce_loss * ((1 - p_t) ** gamma)
        if alpha >= 0:
            alpha_t = alpha * targets + (1 - alpha) * (1 - targets)
            loss = alpha_t * loss
        if reduction == "mean":
            loss = loss.mean()
        elif reduction == "sum":
            loss = loss.sum()
        return loss

    def smooth_l1_loss(input, target, beta: float, reduction: str = "none"):
        if beta < 1e-5:
            loss = torch.abs(input - target)
        else:
            n = torch.abs(input - target)
            cond = n < beta
            loss = torch.where(cond, 0.5 * n ** 2 / beta, n - 0.5 * beta)
        if reduction == "mean":
            loss = loss.mean()
        elif reduction == "sum":
            loss = loss.sum()
        return loss

    fv = _pkg("fvcore")
    fvnn = _pkg("fvcore.nn")
    fvnn.sigmoid_focal_loss = sigmoid_focal_loss
    fvnn.sigmoid_focal_loss_jit = sigmoid_focal_loss
    fvnn.smooth_l1_loss = smooth_l1_loss
    fv.nn = fvnn

    # ---- detectron2 package skeleton
    d2 = _pkg("detectron2")
    layers = _pkg("detectron2.layers")

    def cat(tensors, dim=0):  # layers/wrappers.py:15-22
        assert isinstance(tensors, (list, tuple))
        if len(tensors) == 1:
            return tensors[0]
        return torch.cat(tensors, dim)

    layers.cat = cat
    layers.ShapeSpec = lambda **kw: types.SimpleNamespace(**kw)
    nms_mod = _load("detectron2.layers.nms", "detectron2/layers/nms.py")
    layers.batched_nms = nms_mod.batched_nms
    layers.nms = nms_mod.nms
    layers.paste_masks_in_image = None
    d2.layers = layers

    structures = _pkg("detectron2.structures")
    boxes_mod = _load("detectron2.structures.boxes", "detectron2/structures/boxes.py")
    inst_mod = _load("detectron2.structures.instances", "detectron2/structures/instances.py")
    il_mod = _load("detectron2.structures.image_list", "detectron2/structures/image_list.py")
    structures.Boxes = boxes_mod.Boxes
    structures.BoxMode = boxes_mod.BoxMode
    structures.pairwise_iou = boxes_mod.pairwise_iou
    structures.Instances = inst_mod.Instances
    structures.ImageList = il_mod.ImageList
    structures.RotatedBoxes = None
    d2.structures = structures

    utils = _pkg("detectron2.utils")
    reg = _pkg("detectron2.utils.registry")

    class Registry:
        def __init__(self, name):
            self._name = name
            self._map = {}

        def register(self, obj=None):
            if obj is None:
                def deco(o):
                    self._map[o.__name__] = o
                    return o
                return deco
            self._map[obj.__name__] = obj
            return obj

        def get(self, name):
            return self._map[name]

    reg.Registry = Registry
    events = _pkg("detectron2.utils.events")
    storage = _Storage()
    events.get_event_storage = lambda: storage
    logger = _pkg("detectron2.utils.logger")
    logger.log_first_n = lambda *a, **k: None
    utils.registry, utils.events, utils.logger = reg, events, logger

    config = _pkg("detectron2.config")
    config.global_cfg = make_cfg()
    d2.config = config

    modeling = _pkg("detectron2.modeling")
    for name, attrs in (
        ("detectron2.modeling.anchor_generator", {"build_anchor_generator": None}),
        ("detectron2.modeling.backbone", {"build_backbone": None}),
        ("detectron2.modeling.postprocessing", {"detector_postprocess": None}),
    ):
        m = _pkg(name)
        for k, v in attrs.items():
            setattr(m, k, v)
    meta = _pkg("detectron2.modeling.meta_arch")
    build = _pkg("detectron2.modeling.meta_arch.build")
    build.META_ARCH_REGISTRY = Registry("META_ARCH")
    matcher_mod = _load("detectron2.modeling.matcher", "detectron2/modeling/matcher.py")
    b2b_mod = _load("detectron2.modeling.box_regression", "detectron2/modeling/box_regression.py")
    retina_mod = _load(
        "detectron2.modeling.meta_arch.retinanet", "detectron2/modeling/meta_arch/retinanet.py"
    )

    # ---- the fork's project package
    imb = _pkg("imbalancedetection")
    imb.__path__ = [os.path.join(REF_ROOT, "ImbalanceDetection", "imbalancedetection")]
    mdl = _pkg("imbalancedetection.modelling")
    mdl.__path__ = [os.path.join(imb.__path__[0], "modelling")]
    _load("imbalancedetection.build", "ImbalanceDetection/imbalancedetection/build.py")
    _load("imbalancedetection.modelling.unet", "ImbalanceDetection/imbalancedetection/modelling/unet.py")
    _load(
        "imbalancedetection.modelling.pre_post_models",
        "ImbalanceDetection/imbalancedetection/modelling/pre_post_models.py",
    )
    gh_mod = _load("imbalancedetection.gambler_heads", "ImbalanceDetection/imbalancedetection/gambler_heads.py")

    # the real post-processing and anchor-generator modules (their stand-ins above only satisfy
    # retinanet.py's imports), loaded under private names
    post_mod = _load("detectron2.modeling._postprocessing_real", "detectron2/modeling/postprocessing.py")
    ag_mod = _load("detectron2.modeling._anchor_generator_real", "detectron2/modeling/anchor_generator.py")

    # two-stage callers: rpn_outputs.py and fast_rcnn.py are pure Python + torch
    mem = _pkg("detectron2.utils.memory")
    mem.retry_if_cuda_oom = lambda f: f
    utils.memory = mem
    _load("detectron2.modeling.sampling", "detectron2/modeling/sampling.py")
    _pkg("detectron2.modeling.proposal_generator")
    rpn_mod = _load("detectron2.modeling.proposal_generator.rpn_outputs",
                    "detectron2/modeling/proposal_generator/rpn_outputs.py")
    _pkg("detectron2.modeling.roi_heads")
    frcnn_mod = _load("detectron2.modeling.roi_heads.fast_rcnn", "detectron2/modeling/roi_heads/fast_rcnn.py")

    ns = types.SimpleNamespace(
        rpn_outputs=rpn_mod,
        fast_rcnn=frcnn_mod,
        detector_postprocess=post_mod.detector_postprocess,
        anchor_generator=ag_mod,
        Boxes=boxes_mod.Boxes,
        pairwise_iou=boxes_mod.pairwise_iou,
        Instances=inst_mod.Instances,
        Matcher=matcher_mod.Matcher,
        Box2BoxTransform=b2b_mod.Box2BoxTransform,
        batched_nms=nms_mod.batched_nms,
        nms=nms_mod.nms,
        RetinaNet=retina_mod.RetinaNet,
        retinanet=retina_mod,
        gambler_heads=gh_mod,
        LayeredUnetGambler=gh_mod.LayeredUnetGambler,
        config_module=config,
        storage=storage,
        make_cfg=make_cfg,
    )
    _LOADED = ns
    return ns


def retinanet_self(ref, num_classes=80, alpha=0.25, gamma=2.0, beta=0.1, score_thresh=0.05,
                   topk=1000, nms_thresh=0.5, max_det=100, iou_thresholds=(0.4, 0.5),
                   iou_labels=(0, -1, 1), bbox_weights=(1.0, 1.0, 1.0, 1.0)):
    """``self`` stand-in for calling RetinaNet methods unbound (retinanet.py:69-100)."""
    return types.SimpleNamespace(
        num_classes=num_classes,
        focal_loss_alpha=alpha,
        focal_loss_gamma=gamma,
        smooth_l1_loss_beta=beta,
        score_threshold=score_thresh,
        topk_candidates=topk,
        nms_threshold=nms_thresh,
        max_detections_per_image=max_det,
        box2box_transform=ref.Box2BoxTransform(weights=bbox_weights),
        matcher=ref.Matcher(list(iou_thresholds), list(iou_labels), allow_low_quality_matches=True),
        picky_matcher=ref.Matcher([0.4, 0.9], list(iou_labels), allow_low_quality_matches=True),
    )


def gambler_self(ref, cfg):
    """LayeredUnetGambler without building the U-Net (gambler_heads.py:436-443)."""
    g = ref.LayeredUnetGambler.__new__(ref.LayeredUnetGambler)
    torch.nn.Module.__init__(g)
    g.cfg = cfg
    g.mode = cfg.MODEL.GAMBLER_HEAD.GAMBLER_LOSS_MODE
    g.alpha = cfg.MODEL.RETINANET.FOCAL_LOSS_ALPHA
    g.focal_gamma = cfg.MODEL.RETINANET.FOCAL_LOSS_GAMMA
    g.normalize_w = cfg.MODEL.GAMBLER_HEAD.NORMALIZE
    g.gambler_output = cfg.MODEL.GAMBLER_HEAD.GAMBLER_OUTPUT
    g.in_layers = cfg.MODEL.GAMBLER_HEAD.IN_LAYERS
    g.gamma = cfg.MODEL.GAMBLER_HEAD.GAMBLER_GAMMA
    return g


def set_global_cfg(ref, cfg):
    """gambler_heads.py reads a process-global cfg (`from detectron2.config import global_cfg`)."""
    ref.config_module.global_cfg = cfg
    ref.gambler_heads.global_cfg = cfg
