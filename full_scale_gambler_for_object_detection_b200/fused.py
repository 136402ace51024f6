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
"""
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
        return self.scalars[3] / self.per_anchor_loss.numel()

    def lower_bound(self, temperature, kappa=1.0):
        """-get_loss_upper_bound (gambler_heads.py:17-31, 583-587)."""
        N, R = self.per_anchor_loss.shape
        w_max = (1 + temperature) / (R * temperature + 1)
        return -(kappa * w_max * N) * self.scalars[4]

    def per_level_loss(self, grids, A):
        """(N,R) -> list[(N,A,H,W)] views, the NAKHW_loss layout (gambler_heads.py:91-101, :218)."""
        N = self.per_anchor_loss.shape[0]
        out, off = [], 0
        for H, W in grids:
            n = H * W * A
            out.append(self.per_anchor_loss[:, off:off + n].reshape(N, H, W, A).permute(0, 3, 1, 2))
            off += n
        return out


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, pred_deltas, bets, anchors, gt, cfg, coeffs, detach_pred, group, want_weights):
        c_cls, c_reg, c_gam = coeffs
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
        gl, gd, gb = ctx.saved_tensors
        # grads were produced for a unit upstream gradient; rescale on the device (no-op kernel when 1.0)
        if gl is not None:
            ops.scale_(gl, g_total)
        if gd is not None:
            ops.scale_(gd, g_total)
        ops.scale_(gb, g_total)
        return gl if ctx.has_gl else None, gd, gb, None, None, None, None, None, None, None


def dense_train_step(logits, pred_deltas, bets, anchors, gt, cfg, coeffs=(1.0, 1.0, -1.0), detach_pred=False,
                     group=None, want_weights=False):
    """Fused match + loss step.

    logits (N,R,K), pred_deltas (N,R,4), bets (N,R): CUDA fp32, the flattened (N, sum HWA, .) layout;
    anchors (R,4) or (N,R,4); gt: ops.PackedGT; cfg: DenseLossConfig.
    coeffs = (c_cls, c_reg, c_gam): the scalar returned (and differentiated) is
    ``c_cls*loss_cls + c_reg*loss_box_reg + c_gam*gambler_loss``; the detector phase of the reference is
    (1, lambda_reg, -lambda_out*kappa), the gambler phase (detach_pred=True) is (0, 0, kappa).
    group: a torch.distributed process group when the batch is sharded by image over ranks.
    """
    if detach_pred:
        logits = logits.detach()
    res = _FusedStep.apply(logits, pred_deltas, bets, anchors, gt, cfg, tuple(float(c) for c in coeffs),
                           bool(detach_pred), group, bool(want_weights))
    total, scalars, stats, ell, gtc, mask = res[:6]
    return StepResult(total=total, scalars=scalars, stats=stats, per_anchor_loss=ell, gt_classes=gtc, mask=mask,
                      weights=res[6] if len(res) > 6 else None)
