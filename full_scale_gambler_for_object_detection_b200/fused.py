"""The fused per-anchor training step: GT assignment (both matchers) + RetinaNet losses + gambler loss,
forward and backward, in four kernel launches.

This is what the reference does across ``RetinaNet.get_ground_truth`` / ``get_picky_ground_truth`` /
``losses`` (detectron2/modeling/meta_arch/retinanet.py:201-248, 309-429),
``LayeredUnetGambler.gambler_loss`` (ImbalanceDetection/imbalancedetection/gambler_heads.py:502-602) and
the loss combination in ``GANTrainer.calc_log_metrics`` (ImbalanceDetection/train_net.py:1089-1098),
followed by ``loss.backward()`` through all of it.

    launch 1  K1 pass A   IoU, per-anchor argmax, per-GT max            (anchors streamed once)
    launch 2  K1 pass B   low-quality pass, labels, gt_classes, picky mask, num_foreground, S[n]
    [sharded] all-reduce(SUM) of [num_foreground, S_batch] over the process group
    launch 3  K2 main     focal + smooth-L1 + gambler weighting, fwd + bwd (logits read once, grads written once)
    launch 4  K2 post     d/d bets

Two layouts: ``dense_train_step`` / ``DenseStepPlan`` take the reference's flattened (N, sum HWA, K) tensors;
``dense_train_step_levels`` / ``DenseStepPlanLevels`` take what the heads and the gambler actually produce -- per-level
(N, A*K, H, W), (N, A*4, H, W) and (N, A, H, W) tensors -- and run the same four launches on them in place
(SURVEY.md section 8f row 2), returning gradients and NAKHW_loss in that layout.
"""
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib, ops, sharded


@dataclass
class DenseLossConfig:
    """Scalars the reference reads from cfg on this path (config/defaults.py:400-435, imbalancedetection/config.py)."""

    num_classes: int = 80
    iou_thresholds: Sequence[float] = (0.4, 0.5)        # MODEL.RETINANET.IOU_THRESHOLDS
    iou_labels: Sequence[int] = (0, -1, 1)              # MODEL.RETINANET.IOU_LABELS
    picky_thresholds: Sequence[float] = (0.4, 0.9)      # retinanet.py:96-100
    focal_alpha: float = 0.25
    focal_gamma: float = 2.0
    smooth_l1_beta: float = 0.1
    bbox_reg_weights: Sequence[float] = (1.0, 1.0, 1.0, 1.0)
    gambler_temperature: float = 0.1                    # GAMBLER_TEMPERATURE
    gambler_gamma: float = 1.0                          # GAMBLER_GAMMA
    gambler_loss_mode: str = "focal"                    # GAMBLER_LOSS_MODE: focal | sigmoid
    gambler_output: str = "L_BAHW"                      # L_BAHW | L_BAHW_extendtobatch
    normalize: bool = True                              # NORMALIZE
    gambler_kappa: float = 1.0                          # GAMBLER_KAPPA (logging bound only)

    @property
    def norm_mode(self):
        if not self.normalize:
            return _lib.NORM_NONE
        return _lib.NORM_BATCH if self.gambler_output == "L_BAHW_extendtobatch" else _lib.NORM_IMAGE

    def loss_params(self, c_cls, c_reg, c_gam):
        if self.gambler_output not in ("L_BAHW", "L_BAHW_extendtobatch"):
            # gambler_heads.py:518-520 admits L_B1HW too, but its mask broadcast (:568-569) makes the
            # shapes disagree in calc_gambler_loss, so the reference raises for it as well.
            raise ValueError("unsupported GAMBLER_OUTPUT %r" % (self.gambler_output,))
        return ops.make_loss_params(self.num_classes, self.focal_alpha, self.focal_gamma, self.smooth_l1_beta,
                                    self.gambler_temperature, self.gambler_gamma, self.gambler_loss_mode,
                                    self.norm_mode, c_cls, c_reg, c_gam, self.bbox_reg_weights)


@dataclass
class StepResult:
    total: torch.Tensor                 # c_cls*loss_cls + c_reg*loss_box_reg + c_gam*gambler_loss (differentiable)
    scalars: torch.Tensor               # raw double scalars (include/fsg_dense.h)
    stats: torch.Tensor                 # [num_foreground, S_batch, S[n]...]
    per_anchor_loss: torch.Tensor       # (N,R)  NAKHW_loss values
    gt_classes: torch.Tensor            # (N,R) int64
    mask: torch.Tensor                  # (N,R) int64
    weights: Optional[torch.Tensor] = None  # (N,R) normalised bets (detached)
    extras: dict = field(default_factory=dict)

    # device scalars, no host sync
    @property
    def loss_cls(self):
        return self.scalars[5]

    @property
    def loss_box_reg(self):
        return self.scalars[6]

    @property
    def gambler_loss(self):
        return self.scalars[7]

    @property
    def num_foreground(self):
        return self.stats[0]

    def loss_before_weighting(self, mode="focal"):
        """gambler_heads.py:589-594."""
        if mode == "focal":
            return self.scalars[3] / torch.clamp(self.stats[0], min=1.0)
        return self.scalars[3] / self.gt_classes.numel()

    def lower_bound(self, temperature, kappa=1.0):
        """-get_loss_upper_bound (gambler_heads.py:17-31, 583-587)."""
        N, R = self.per_anchor_loss.shape
        w_max = (1 + temperature) / (R * temperature + 1)
        return -(kappa * w_max * N) * self.scalars[4]

    def log_metrics(self, bets, cfg, lambda_reg=1.0, kappa=1.0, lambda_out=1.0, mode="cls+reg-gambler"):
        """GANTrainer.calc_log_metrics (train_net.py:1089-1124) as a dict of device scalars, no host sync: the loss
        combination plus the bet / weight statistics (sum / max / mean of the masked betting maps, sum / max / mean /
        median of the normalised weights).  ``bets``: the (N,R) tensor or the list of per-level (N, A, H, W) maps the
        step was given (unmasked; the step's picky mask is applied here as gambler_loss applies it in place)."""
        params = cfg.loss_params(1.0, 1.0, 1.0)
        if isinstance(bets, (list, tuple)):
            st = ops.bet_stats(None, self.mask, params, self.stats, bet_levels=[b.detach() for b in bets])
        else:
            st = ops.bet_stats(bets.detach().to(torch.float32).contiguous(), self.mask, params, self.stats)
        d = {"loss_cls": self.loss_cls, "loss_box_reg": self.loss_box_reg * lambda_reg,
             "loss_gambler": self.gambler_loss * kappa,
             "loss_before_weighting": self.loss_before_weighting(cfg.gambler_loss_mode)}
        if mode == "cls+reg-gambler":
            d["loss_detector"] = d["loss_box_reg"] + d["loss_cls"] - lambda_out * d["loss_gambler"]
        elif mode == "weighted_cls_with_gambler+reg":
            d["loss_detector"] = d["loss_box_reg"] - lambda_out * d["loss_gambler"]
        else:
            raise ValueError("unknown DETECTOR_LOSS_MODE %r" % (mode,))
        for i, name in enumerate(ops.BET_STAT_NAMES):
            d[name] = st[i]
        return d

    def per_level_loss(self, grids, A):
        """(N,R) -> list[(N,A,H,W)] views, the NAKHW_loss layout (gambler_heads.py:91-101, :218)."""
        N = self.per_anchor_loss.shape[0]
        out, off = [], 0
        for H, W in grids:
            n = H * W * A
            out.append(self.per_anchor_loss[:, off:off + n].reshape(N, H, W, A).permute(0, 3, 1, 2))
            off += n
        return out


def _match_config(cfg):
    mc = _lib.MatchConfig()
    mc.num_thresholds, mc.num_picky_thresholds = len(cfg.iou_thresholds), len(cfg.picky_thresholds)
    mc.allow_low_quality_matches = 1
    for i, v in enumerate(cfg.iou_thresholds):
        mc.thresholds[i] = float(v)
    for i, v in enumerate(cfg.picky_thresholds):
        mc.picky_thresholds[i] = float(v)
    for i, v in enumerate(cfg.iou_labels):
        mc.labels[i] = int(v)
        mc.picky_labels[i] = int(v)
    return mc


class DenseStepPlan:
    """Pre-allocated buffers, cached host-side arguments and (optionally) a CUDA graph for the fused step
    at fixed shapes (N images, R anchors, K classes).  ``run`` enqueues the four kernels with no allocation
    and ~4 ctypes calls; ``capture``/``replay`` turn that into one graph launch, which is how a B200 keeps up
    with a step that is only ~200 us of device time.

    Buffers returned by ``run`` are owned by the plan and overwritten by the next ``run``/``replay``.
    """

    def __init__(self, N, R, K, cfg, device, coeffs=(1.0, 1.0, -1.0), detach_pred=False, want_weights=False,
                 max_total_gt=4096, group=None, peer=None):
        assert K == cfg.num_classes
        self.N, self.R, self.K, self.cfg, self.device = N, R, K, cfg, torch.device(device)
        self.coeffs = tuple(float(c) for c in coeffs)
        self.detach_pred, self.group, self.max_total_gt = bool(detach_pred), group, int(max_total_gt)
        # peer: sharded.PeerExchange -- the [num_foreground, S_batch] all-reduce then happens inside K1's
        # second kernel over NVLink peer memory instead of an NCCL launch (and the step is one graph again)
        self.peer = peer if (peer is not None and group is not None) else None
        self.params = cfg.loss_params(*self.coeffs)
        L = self.L = ops.lib()
        dev = self.device
        f32, i64 = torch.float32, torch.int64
        self.gt_classes = torch.empty((N, R), dtype=i64, device=dev)
        self.mask = torch.empty((N, R), dtype=i64, device=dev)
        self.matched = torch.empty((N, R), dtype=torch.int32, device=dev)
        self.stats = torch.zeros(_lib.STATS_HEADER + N, dtype=torch.float64, device=dev)
        self.scalars = torch.zeros(_lib.SCALARS_HEADER + N, dtype=torch.float64, device=dev)
        self.ell = torch.empty((N, R), dtype=f32, device=dev)
        self.weights = torch.empty((N, R), dtype=f32, device=dev) if want_weights else None
        self.need_gl = not self.detach_pred
        self.need_gd = self.coeffs[1] != 0.0
        self.grad_logits = torch.empty((N, R, K), dtype=f32, device=dev) if self.need_gl else None
        self.grad_deltas = torch.empty((N, R, 4), dtype=f32, device=dev) if self.need_gd else None
        self.grad_bets = torch.empty((N, R), dtype=f32, device=dev)
        self.total = torch.zeros((), dtype=f32, device=dev)
        self.ws_match = torch.empty(max(16, L.fsg_match_workspace_bytes(N, R, self.max_total_gt)), dtype=torch.uint8,
                                    device=dev)
        self.ws_loss = torch.empty(max(16, L.fsg_loss_main_workspace_bytes(N, R, K)), dtype=torch.uint8, device=dev)
        self._thr = _lib.host_f32(cfg.iou_thresholds)
        self._lab = _lib.host_i8(cfg.iou_labels)
        self._pthr = _lib.host_f32(cfg.picky_thresholds)
        self._bw = _lib.host_f32(cfg.bbox_reg_weights)
        self.graph = None
        self._static = None
        # one-call form of the step (fsg_dense_step: persistent K1, K2 under programmatic dependent launch, the peer
        # exchange polled by K2): everything except a sharded run that still needs an NCCL collective in the middle
        self.one_call = (os.environ.get("FSG_STEP_STAGED", "0") != "1") and (
            group is None or (self.peer is not None and cfg.norm_mode != _lib.NORM_BATCH))
        # zero-filled once and owned by this plan: every fsg_dense_step call leaves it clean again, so no memset node
        # is enqueued (fsg_match_config.workspace_is_clean; a memset node costs ~4 us of the step inside a graph)
        self.ws_step = torch.zeros(max(16, L.fsg_dense_step_workspace_bytes(N, R, K, self.max_total_gt)),
                                   dtype=torch.uint8, device=dev) if self.one_call else None
        self._mc = _match_config(cfg)
        self._mc.workspace_is_clean = 1

    def _check(self, logits, deltas, bets, anchors, gt):
        N, R, K = self.N, self.R, self.K
        for t, shape in ((logits, (N, R, K)), (deltas, (N, R, 4)), (bets, (N, R))):
            if tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise RuntimeError("DenseStepPlan: expected contiguous CUDA fp32 tensor of shape %s" % (shape,))
        if gt.num_images != N or gt.total > self.max_total_gt:
            raise RuntimeError("DenseStepPlan: GT does not fit the plan (images %d, boxes %d > %d)"
                               % (gt.num_images, gt.total, self.max_total_gt))
        if anchors.dtype != torch.float32 or not anchors.is_contiguous() or anchors.shape[-2:] != (R, 4):
            raise RuntimeError("DenseStepPlan: anchors must be contiguous fp32 (R,4) or (N,R,4)")

    # ---- the three stages (K1 = two launches, K2 main, K2 post) ---------------------------------------
    def stage_match(self, bets, anchors, gt):
        L, N, R, cfg, P = self.L, self.N, self.R, self.cfg, _lib.ptr
        a_stride = R * 4 if anchors.dim() == 3 else 0
        _lib.check(L.fsg_match_anchors(
            P(anchors), R, a_stride, P(gt.boxes), P(gt.classes), P(gt.offsets), N, gt.total, cfg.num_classes,
            self._thr, self._lab, len(cfg.iou_thresholds), 1, self._pthr, self._lab, len(cfg.picky_thresholds),
            self._bw, None, None, None, P(self.gt_classes), P(self.mask), None, P(self.matched), P(bets), None,
            float(cfg.gambler_temperature), P(self.stats), self.peer.ctx if self.peer is not None else None,
            P(self.ws_match), self.ws_match.numel(), _lib.stream()))
        _lib.count_launches(3)

    def stage_main(self, logits, deltas, bets, anchors, gt):
        L, N, R, P = self.L, self.N, self.R, _lib.ptr
        a_stride = R * 4 if anchors.dim() == 3 else 0
        _lib.check(L.fsg_loss_main(
            P(logits), P(deltas), None, P(anchors), a_stride, P(gt.boxes), P(gt.offsets), P(self.matched),
            P(self.gt_classes), P(self.mask), P(bets), N, R, self.params, P(self.stats), P(self.grad_logits),
            P(self.grad_deltas), P(self.ell), P(self.weights), P(self.scalars), P(self.ws_loss),
            self.ws_loss.numel(), _lib.stream()))
        _lib.count_launches(1)

    def stage_post(self, bets):
        P = _lib.ptr
        _lib.check(self.L.fsg_loss_post(P(bets), P(self.mask), P(self.ell), self.N, self.R, self.params,
                                        P(self.stats), P(self.scalars), P(self.grad_bets), _lib.stream()))
        _lib.count_launches(1)

    def step_one_call(self, logits, deltas, bets, anchors, gt):
        """The whole step through ``fsg_dense_step`` (one ctypes call, five kernels, no memset)."""
        P = _lib.ptr
        io = _lib.StepIO()
        io.logits, io.pred_deltas, io.bets, io.anchors = P(logits), P(deltas), P(bets), P(anchors)
        io.anchor_image_stride = self.R * 4 if anchors.dim() == 3 else 0
        io.gt_boxes, io.gt_class_ids, io.gt_offsets, io.sum_M = P(gt.boxes), P(gt.classes), P(gt.offsets), gt.total
        io.gt_classes, io.mask, io.matched_idx32 = P(self.gt_classes), P(self.mask), P(self.matched)
        io.stats, io.scalars = P(self.stats), P(self.scalars)
        io.grad_logits, io.grad_deltas, io.grad_bets = P(self.grad_logits), P(self.grad_deltas), P(self.grad_bets)
        io.per_anchor_loss, io.weights_out = P(self.ell), P(self.weights)
        _lib.check(self.L.fsg_dense_step(io, self.N, self.R, self._mc, self.params,
                                         self.peer.ctx if self.peer is not None else None, P(self.ws_step),
                                         self.ws_step.numel(), _lib.stream()))
        _lib.count_launches(5)

    def run(self, logits, deltas, bets, anchors, gt):
        """Enqueue the step on the current stream.  Inputs: detached contiguous CUDA fp32 tensors."""
        self._check(logits, deltas, bets, anchors, gt)
        if self.one_call:
            self.step_one_call(logits, deltas, bets, anchors, gt)
            return self.result()
        self.stage_match(bets, anchors, gt)
        if self.group is not None and self.peer is None:
            sharded.all_reduce_stats(self.stats, self.group)
        self.stage_main(logits, deltas, bets, anchors, gt)
        if self.group is not None and self.cfg.norm_mode == _lib.NORM_BATCH:
            sharded.all_reduce_batch_weighted_sum(self.scalars, self.group)
        self.stage_post(bets)
        return self.result()

    def result(self):
        return StepResult(total=self.scalars[8], scalars=self.scalars, stats=self.stats, per_anchor_loss=self.ell,
                          gt_classes=self.gt_classes, mask=self.mask, weights=self.weights,
                          extras={"grad_logits": self.grad_logits, "grad_deltas": self.grad_deltas,
                                  "grad_bets": self.grad_bets})

    # ---- CUDA graph -----------------------------------------------------------------------------
    def capture(self, logits, deltas, bets, anchors, gt, warmup=2):
        """Capture the step reading from exactly these tensors (their storage must stay alive and is
        refreshed in place by the caller between replays).

        Single process: one graph (the five kernels of the step).  Sharded (``group`` set): the collectives stay
        *outside* the graphs -- graph 1 = K1, eager all-reduce of the two scalars, graph 2 = K2 main (+ post);
        for ``L_BAHW_extendtobatch`` the post pass is a third graph after the second all-reduce."""
        self._check(logits, deltas, bets, anchors, gt)
        self._static = (logits, deltas, bets, anchors, gt)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.run(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)

        def cap(fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g

        if self.one_call or self.group is None or (self.peer is not None and self.cfg.norm_mode != _lib.NORM_BATCH):
            self.graph = [cap(lambda: self.run(*self._static))]
        else:
            batch_norm = self.cfg.norm_mode == _lib.NORM_BATCH
            g1 = cap(lambda: self.stage_match(bets, anchors, gt))
            if batch_norm:
                g2 = cap(lambda: self.stage_main(logits, deltas, bets, anchors, gt))
                g3 = cap(lambda: self.stage_post(bets))
                self.graph = [g1, g2, g3]
            else:
                g2 = cap(lambda: (self.stage_main(logits, deltas, bets, anchors, gt), self.stage_post(bets)))
                self.graph = [g1, g2]
        return self

    def replay(self):
        """Graph launch(es) of the whole step; see ``capture``."""
        g = self.graph
        g[0].replay()
        if len(g) > 1:
            if self.peer is None:
                sharded.all_reduce_stats(self.stats, self.group)
            g[1].replay()
            if len(g) == 3:
                sharded.all_reduce_batch_weighted_sum(self.scalars, self.group)
                g[2].replay()
        _lib.count_launches(5)
        return self.result()

    def release_graphs(self):
        self.graph = None
        self._static = None


_single_use = ops.single_use


def _grad_mul(group, grad_reduction):
    """Gradients of a sharded step are those of the WHOLE-batch loss (global num_foreground) w.r.t. this rank's
    inputs.  Under DistributedDataParallel the parameter gradients are then AVERAGED over ranks, which would leave
    1/world of the single-process whole-batch gradient (and of the reference's own DDP result); 'mean' therefore
    multiplies the outgoing gradients by world_size.  'sum': leave them for a SUM reduction."""
    if grad_reduction not in ("mean", "sum"):
        raise ValueError("grad_reduction must be 'mean' (DDP averages gradients) or 'sum'")
    if group is None or grad_reduction == "sum":
        return 1.0
    import torch.distributed as dist
    return float(dist.get_world_size(group))


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, pred_deltas, bets, anchors, gt, cfg, coeffs, detach_pred, group, want_weights,
                plan=None, grad_mul=1.0):
        c_cls, c_reg, c_gam = coeffs
        ctx.grad_mul, ctx.used = float(grad_mul), False
        if plan is not None:
            x = logits.detach()
            x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
            r = plan.run(x, pred_deltas.detach().to(torch.float32).contiguous(),
                         bets.detach().to(torch.float32).contiguous(), anchors, gt)
            ctx.save_for_backward(plan.grad_logits, plan.grad_deltas, plan.grad_bets)
            ctx.has_gl = plan.need_gl
            total = r.scalars[8].to(torch.float32)
            nd = [r.scalars, r.stats, r.per_anchor_loss, r.gt_classes, r.mask]
            if r.weights is not None:
                nd.append(r.weights)
            ctx.mark_non_differentiable(*nd)
            return (total,) + tuple(nd)
        params = cfg.loss_params(c_cls, c_reg, c_gam)
        logits_c = logits.detach()
        if logits_c.dtype != torch.float32 or not logits_c.is_contiguous():
            logits_c = logits_c.to(torch.float32).contiguous()
        deltas_c = pred_deltas.detach().to(torch.float32).contiguous()
        bets_c = bets.detach().to(torch.float32).contiguous()
        N, R, K = logits_c.shape
        assert K == cfg.num_classes and deltas_c.shape == (N, R, 4) and bets_c.shape == (N, R)

        m = ops.match_anchors(anchors, gt, cfg.num_classes, cfg.iou_thresholds, cfg.iou_labels,
                              cfg.picky_thresholds, None, cfg.bbox_reg_weights,
                              want=("gt_classes", "mask", "matched_idx32"), bets=bets_c,
                              temperature=cfg.gambler_temperature)
        stats = m["stats"]
        if group is not None:
            # the path's only exchange step before the main pass: global foreground count (+ batch normaliser)
            sharded.all_reduce_stats(stats, group)
        need_gl = not detach_pred
        need_gd = c_reg != 0.0
        out = ops.loss_main(logits_c, m["gt_classes"], params, stats, pred_deltas=deltas_c, anchors=anchors, gt=gt,
                            matched_idx32=m["matched_idx32"], mask=m["mask"], bets=bets_c,
                            want_grad_logits=need_gl, want_grad_deltas=need_gd, want_weights=want_weights)
        scalars = out["scalars"]
        if group is not None and cfg.norm_mode == _lib.NORM_BATCH:
            sharded.all_reduce_batch_weighted_sum(scalars, group)
        grad_bets = ops.loss_post(bets_c, m["mask"], out["per_anchor_loss"], params, stats, scalars)
        ctx.save_for_backward(out.get("grad_logits"), out.get("grad_deltas"), grad_bets)
        ctx.has_gl = need_gl
        total = scalars[8].to(torch.float32)
        weights = out.get("weights")
        nd = [scalars, stats, out["per_anchor_loss"], m["gt_classes"], m["mask"]]
        if weights is not None:
            nd.append(weights)
        ctx.mark_non_differentiable(*nd)
        if weights is None:
            return total, scalars, stats, out["per_anchor_loss"], m["gt_classes"], m["mask"]
        return total, scalars, stats, out["per_anchor_loss"], m["gt_classes"], m["mask"], weights

    @staticmethod
    def backward(ctx, g_total, *unused):
        _single_use(ctx)
        gl, gd, gb = ctx.saved_tensors
        # grads were produced for a unit upstream gradient; rescale on the device (no-op kernel when 1.0)
        if gl is not None:
            ops.scale_(gl, g_total, ctx.grad_mul)
        if gd is not None:
            ops.scale_(gd, g_total, ctx.grad_mul)
        ops.scale_(gb, g_total, ctx.grad_mul)
        return gl if ctx.has_gl else None, gd, gb, None, None, None, None, None, None, None, None, None


def dense_train_step(logits, pred_deltas, bets, anchors, gt, cfg, coeffs=(1.0, 1.0, -1.0), detach_pred=False,
                     group=None, want_weights=False, plan=None, grad_reduction="mean"):
    """Fused match + loss step.

    logits (N,R,K), pred_deltas (N,R,4), bets (N,R): CUDA fp32, the flattened (N, sum HWA, .) layout;
    anchors (R,4) or (N,R,4); gt: ops.PackedGT; cfg: DenseLossConfig.
    coeffs = (c_cls, c_reg, c_gam): the scalar returned (and differentiated) is
    ``c_cls*loss_cls + c_reg*loss_box_reg + c_gam*gambler_loss``; the detector phase of the reference is
    (1, lambda_reg, -lambda_out*kappa), the gambler phase (detach_pred=True) is (0, 0, kappa).
    group: a torch.distributed process group when the batch is sharded by image over ranks.
    grad_reduction (with ``group``): how the caller reduces parameter gradients over ranks -- 'mean'
    (DistributedDataParallel, the reference's deployment: the gradients leaving this node are multiplied by
    world_size so that the averaged parameter gradient equals the single-process whole-batch one) or 'sum'.
    The loss values are unaffected.
    plan: a DenseStepPlan built for these shapes/coefficients: no allocation, outputs live in the plan.
    The autograd node is single-use (see ``_single_use``).
    """
    if plan is not None:
        assert plan.coeffs == tuple(float(c) for c in coeffs) and plan.detach_pred == bool(detach_pred)
    if detach_pred:
        logits = logits.detach()
    res = _FusedStep.apply(logits, pred_deltas, bets, anchors, gt, cfg, tuple(float(c) for c in coeffs),
                           bool(detach_pred), group, bool(want_weights), plan, _grad_mul(group, grad_reduction))
    total, scalars, stats, ell, gtc, mask = res[:6]
    return StepResult(total=total, scalars=scalars, stats=stats, per_anchor_loss=ell, gt_classes=gtc, mask=mask,
                      weights=res[6] if len(res) > 6 else None)


class _FusedStepLevels(torch.autograd.Function):
    """The fused step on the head's native layout: every tensor argument is a per-level conv output."""

    @staticmethod
    def forward(ctx, anchors, gt, cfg, coeffs, detach_pred, group, L, grad_mul, *levels):
        c_cls, c_reg, c_gam = coeffs
        ctx.grad_mul, ctx.used = float(grad_mul), False
        f32c = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()
        xs = [f32c(t.detach()) for t in levels[:L]]
        ds = [f32c(t.detach()) for t in levels[L:2 * L]]
        bs = [f32c(t.detach()) for t in levels[2 * L:]]
        params = cfg.loss_params(c_cls, c_reg, c_gam)
        m = ops.match_anchors(anchors, gt, cfg.num_classes, cfg.iou_thresholds, cfg.iou_labels,
                              cfg.picky_thresholds, None, cfg.bbox_reg_weights,
                              want=("gt_classes", "mask", "matched_idx32"), bet_levels=bs,
                              temperature=cfg.gambler_temperature)
        stats = m["stats"]
        if group is not None:
            sharded.all_reduce_stats(stats, group)
        need_gl, need_gd = not detach_pred, c_reg != 0.0
        ell_levels = [torch.empty_like(b) for b in bs]
        out = ops.loss_main_levels(xs, m["gt_classes"], params, stats, delta_levels=ds, anchors=anchors, gt=gt,
                                   matched_idx32=m["matched_idx32"], mask=m["mask"], bet_levels=bs,
                                   ell_levels_out=ell_levels, want_grad_logits=need_gl, want_grad_deltas=need_gd)
        scalars = out["scalars"]
        if group is not None and cfg.norm_mode == _lib.NORM_BATCH:
            sharded.all_reduce_batch_weighted_sum(scalars, group)
        gb_levels = ops.loss_post_levels(bs, m["mask"], ell_levels, params, stats, scalars)
        saved = (out["grad_logits"] if need_gl else []) + (out["grad_deltas"] if need_gd else []) + gb_levels
        ctx.save_for_backward(*saved)
        ctx.cfg_ = (L, need_gl, need_gd)
        nd = [scalars, stats, m["gt_classes"], m["mask"]] + ell_levels
        ctx.mark_non_differentiable(*nd)
        return (scalars[8].to(torch.float32),) + tuple(nd)

    @staticmethod
    def backward(ctx, g_total, *unused):
        _single_use(ctx)
        L, need_gl, need_gd = ctx.cfg_
        saved = list(ctx.saved_tensors)
        for t in saved:
            ops.scale_(t, g_total, ctx.grad_mul)
        gl = saved[:L] if need_gl else [None] * L
        saved = saved[L:] if need_gl else saved
        gd = saved[:L] if need_gd else [None] * L
        saved = saved[L:] if need_gd else saved
        return (None,) * 8 + tuple(gl) + tuple(gd) + tuple(saved)


def dense_train_step_levels(logit_levels, delta_levels, bet_levels, anchors, gt, cfg, coeffs=(1.0, 1.0, -1.0),
                            detach_pred=False, group=None, grad_reduction="mean"):
    """The fused match + loss step straight from the head outputs, no permute/cat copies in either direction.

    logit_levels list[(N, A*K, H, W)], delta_levels list[(N, A*4, H, W)], bet_levels list[(N, A, H, W)] (the
    gambler's betting maps); anchors (R,4) / (N,R,4) and gt as in :func:`dense_train_step`.
    Returns a StepResult whose ``per_anchor_loss`` is the list of (N, A, H, W) NAKHW_loss maps
    (gambler_heads.py:218); gradients arrive on the per-level inputs in their own layout.
    Launches: K1 (2), K2 on the native layout, K2 post on the native layout -- no layout adapter of any size."""
    L = len(logit_levels)
    assert len(delta_levels) == L and len(bet_levels) == L
    xs = [t.detach() for t in logit_levels] if detach_pred else list(logit_levels)
    res = _FusedStepLevels.apply(anchors, gt, cfg, tuple(float(c) for c in coeffs), bool(detach_pred), group, L,
                                 _grad_mul(group, grad_reduction), *(xs + list(delta_levels) + list(bet_levels)))
    total, scalars, stats, gtc, mask = res[:5]
    return StepResult(total=total, scalars=scalars, stats=stats, per_anchor_loss=list(res[5:]), gt_classes=gtc,
                      mask=mask)


class DenseStepPlanLevels:
    """``DenseStepPlan`` for the head's native layout: pre-allocated outputs and (optionally) CUDA graph(s) for
    K1 (2 launches, betting maps read in place) -> K2 on the per-level conv outputs -> K2 post on the per-level
    maps.  ``group`` / ``peer`` as in ``DenseStepPlan``: a batch sharded by image over ranks exchanges
    [num_foreground, S_batch] between K1 and K2 -- inside K1's second kernel over NVLink peer memory (``peer``, one
    graph), else with an eager NCCL all-reduce between two graphs.

    ``level_shapes``: [(H, W)] per level; the head outputs are (N, A*K, H, W), (N, A*4, H, W), (N, A, H, W).
    Outputs (owned by the plan, overwritten by the next run/replay): ``grad_logits`` / ``grad_deltas`` /
    ``grad_bets`` / ``nakhw_loss`` (lists, per-level layout), ``gt_classes``, ``mask``, ``stats``, ``scalars``."""

    def __init__(self, N, level_shapes, A, K, cfg, device, coeffs=(1.0, 1.0, -1.0), detach_pred=False, group=None,
                 peer=None, max_total_gt=4096):
        assert K == cfg.num_classes
        self.N, self.A, self.K, self.cfg, self.device = N, A, K, cfg, torch.device(device)
        self.shapes = [(int(h), int(w)) for h, w in level_shapes]
        self.R = sum(h * w * A for h, w in self.shapes)
        self.coeffs = tuple(float(c) for c in coeffs)
        self.params = cfg.loss_params(*self.coeffs)
        self.need_gl, self.need_gd = not detach_pred, self.coeffs[1] != 0.0
        self.group = group
        self.peer = peer if (peer is not None and group is not None) else None
        self.max_total_gt = int(max_total_gt)
        dev, f32, i64 = self.device, torch.float32, torch.int64
        mk = lambda c: [torch.empty((N, c, h, w), dtype=f32, device=dev) for h, w in self.shapes]
        self.grad_logits = mk(A * K) if self.need_gl else None
        self.grad_deltas = mk(A * 4) if self.need_gd else None
        self.grad_bets, self.nakhw_loss = mk(A), mk(A)
        R = self.R
        self.m = {"gt_classes": torch.empty((N, R), dtype=i64, device=dev),
                  "mask": torch.empty((N, R), dtype=i64, device=dev),
                  "matched_idx32": torch.empty((N, R), dtype=torch.int32, device=dev),
                  "stats": torch.zeros(_lib.STATS_HEADER + N, dtype=torch.float64, device=dev)}
        self.scalars = torch.zeros(_lib.SCALARS_HEADER + N, dtype=torch.float64, device=dev)
        L = ops.lib()
        self.ws_match = torch.empty(max(16, L.fsg_match_workspace_bytes(N, R, self.max_total_gt)), dtype=torch.uint8,
                                    device=dev)
        self.graph, self._static, self._last = None, None, None
        # one-call form (fsg_dense_step_levels: K1 -> K2 native -> post under programmatic dependent launch, the peer
        # exchange polled by K2); a sharded run that needs an NCCL collective in the middle keeps the staged form
        self.one_call = (os.environ.get("FSG_STEP_STAGED", "0") != "1") and (
            group is None or (self.peer is not None and cfg.norm_mode != _lib.NORM_BATCH))
        self._mc = _match_config(cfg)
        self.ws_step = None
        self.L = L

    def step_one_call(self, logit_levels, delta_levels, bet_levels, anchors, gt):
        P = _lib.ptr
        nl, N, A, K = len(self.shapes), self.N, self.A, self.K
        hl = (_lib.HeadLevel * nl)()
        pl = (_lib.PostLevel * nl)()
        for i, (h, w) in enumerate(self.shapes):
            x, d, b = logit_levels[i], delta_levels[i], bet_levels[i]
            for t, c in ((x, A * K), (d, A * 4), (b, A)):
                if tuple(t.shape) != (N, c, h, w) or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError("DenseStepPlanLevels: level %d expects contiguous fp32 (N=%d, %d, %d, %d)"
                                       % (i, N, c, h, w))
            hl[i].logits, hl[i].pred_deltas, hl[i].bets = P(x), P(d), P(b)
            hl[i].grad_logits = P(self.grad_logits[i]) if self.need_gl else None
            hl[i].grad_deltas = P(self.grad_deltas[i]) if self.need_gd else None
            hl[i].per_anchor_loss = P(self.nakhw_loss[i])
            hl[i].H, hl[i].W = h, w
            pl[i].bets, pl[i].per_anchor_loss, pl[i].grad_bets = P(b), P(self.nakhw_loss[i]), P(self.grad_bets[i])
            pl[i].H, pl[i].W = h, w
        if gt.total > self.max_total_gt:
            raise RuntimeError("DenseStepPlanLevels: %d GT boxes > max_total_gt %d" % (gt.total, self.max_total_gt))
        if self.ws_step is None:
            need = self.L.fsg_dense_step_levels_workspace_bytes(N, hl, nl, A, self.max_total_gt)
            self.ws_step = torch.zeros(max(16, need), dtype=torch.uint8, device=self.device)   # kept clean, see above
            self._mc.workspace_is_clean = 1
        io = _lib.StepLevelsIO()
        io.anchors = P(anchors)
        io.anchor_image_stride = self.R * 4 if anchors.dim() == 3 else 0
        io.gt_boxes, io.gt_class_ids, io.gt_offsets, io.sum_M = P(gt.boxes), P(gt.classes), P(gt.offsets), gt.total
        m = self.m
        io.gt_classes, io.mask, io.matched_idx32 = P(m["gt_classes"]), P(m["mask"]), P(m["matched_idx32"])
        io.stats, io.scalars, io.weights_out = P(m["stats"]), P(self.scalars), None
        _lib.check(self.L.fsg_dense_step_levels(io, hl, pl, nl, A, N, self.R, self._mc, self.params,
                                                self.peer.ctx if self.peer is not None else None, P(self.ws_step),
                                                self.ws_step.numel(), _lib.stream()))
        _lib.count_launches(5)

    # ---- stages ---------------------------------------------------------------------------------------
    def stage_match(self, bet_levels, anchors, gt):
        cfg = self.cfg
        if gt.total > self.max_total_gt:
            raise RuntimeError("DenseStepPlanLevels: %d GT boxes > max_total_gt %d" % (gt.total, self.max_total_gt))
        ops.match_anchors(anchors, gt, self.K, cfg.iou_thresholds, cfg.iou_labels, cfg.picky_thresholds, None,
                          cfg.bbox_reg_weights, bet_levels=bet_levels, temperature=cfg.gambler_temperature,
                          workspace=self.ws_match, out=self.m, peer=self.peer)

    def stage_main(self, logit_levels, delta_levels, bet_levels, anchors, gt):
        m = self.m
        ops.loss_main_levels(logit_levels, m["gt_classes"], self.params, m["stats"], delta_levels=delta_levels,
                             anchors=anchors, gt=gt, matched_idx32=m["matched_idx32"], mask=m["mask"],
                             bet_levels=bet_levels, ell_levels_out=self.nakhw_loss,
                             want_grad_logits=self.need_gl, want_grad_deltas=self.need_gd,
                             grad_logits_out=self.grad_logits, grad_deltas_out=self.grad_deltas,
                             scalars_out=self.scalars)

    def stage_post(self, bet_levels):
        ops.loss_post_levels(bet_levels, self.m["mask"], self.nakhw_loss, self.params, self.m["stats"], self.scalars,
                             out=self.grad_bets)

    def result(self):
        m = self.m
        self._last = StepResult(total=self.scalars[8], scalars=self.scalars, stats=m["stats"],
                                per_anchor_loss=self.nakhw_loss, gt_classes=m["gt_classes"], mask=m["mask"],
                                extras={"grad_logits": self.grad_logits, "grad_deltas": self.grad_deltas,
                                        "grad_bets": self.grad_bets})
        return self._last

    def run(self, logit_levels, delta_levels, bet_levels, anchors, gt):
        if self.one_call:
            self.step_one_call(logit_levels, delta_levels, bet_levels, anchors, gt)
            return self.result()
        self.stage_match(bet_levels, anchors, gt)
        if self.group is not None and self.peer is None:
            sharded.all_reduce_stats(self.m["stats"], self.group)
        self.stage_main(logit_levels, delta_levels, bet_levels, anchors, gt)
        if self.group is not None and self.cfg.norm_mode == _lib.NORM_BATCH:
            sharded.all_reduce_batch_weighted_sum(self.scalars, self.group)
        self.stage_post(bet_levels)
        return self.result()

    def capture(self, logit_levels, delta_levels, bet_levels, anchors, gt, warmup=2):
        """Capture the step reading from exactly these tensors (refresh them in place between replays)."""
        xs, ds, bs = list(logit_levels), list(delta_levels), list(bet_levels)
        self._static = (xs, ds, bs, anchors, gt)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.run(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)

        def cap(fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g

        batch_norm = self.cfg.norm_mode == _lib.NORM_BATCH
        if self.one_call or self.group is None or (self.peer is not None and not batch_norm):
            self.graph = [cap(lambda: self.run(*self._static))]
        else:
            g1 = cap(lambda: self.stage_match(bs, anchors, gt))
            if batch_norm:
                self.graph = [g1, cap(lambda: self.stage_main(xs, ds, bs, anchors, gt)), cap(lambda: self.stage_post(bs))]
            else:
                self.graph = [g1, cap(lambda: (self.stage_main(xs, ds, bs, anchors, gt), self.stage_post(bs)))]
        self.result()
        return self

    def replay(self):
        g = self.graph
        g[0].replay()
        if len(g) > 1:
            if self.peer is None:
                sharded.all_reduce_stats(self.m["stats"], self.group)
            g[1].replay()
            if len(g) == 3:
                sharded.all_reduce_batch_weighted_sum(self.scalars, self.group)
                g[2].replay()
        _lib.count_launches(5)
        return self._last

    def release_graphs(self):
        self.graph = None
        self._static = None
