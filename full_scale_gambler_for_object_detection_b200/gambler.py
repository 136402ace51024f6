"""The gambler's loss arithmetic (ImbalanceDetection/imbalancedetection/gambler_heads.py) with the
reference's call signature: ``gambler_loss(pred_class_logits, weights, gt_classes, mask, detach_pred)``
(:502-602) -> ``(loss_dict, weights)``; ``get_loss_upper_bound`` (:17-31).

Reachable output modes: ``L_BAHW`` and ``L_BAHW_extendtobatch`` (:518-520).  ``L_B1HW`` passes the assert
but the mask broadcast at :568-569 turns the (N,1,H,W) bets into (N,3,H,W) and the shapes no longer agree in
``calc_gambler_loss``; it raises in the reference and is rejected here.
"""
import torch

from . import _lib, ops
from .retinanet import levels_to_flat


def get_loss_upper_bound(nakhw, N, smoothing, kappa):
    """gambler_heads.py:17-31 (device tensor in, device scalar out -- no CPU staging)."""
    assert len(nakhw) == 5, "only works with 5 fpn layers"
    total = 0
    per_level = []
    for layer in nakhw:
        total += layer.shape[1] * layer.shape[2] * layer.shape[3]
        per_level.append(layer.reshape(N, -1).max(dim=1).values)
    max_loss = torch.stack(per_level, dim=1).max(dim=1).values
    w_max = (1 + smoothing) / (total * smoothing + 1)
    return kappa * w_max * N * max_loss.sum()


class _GamblerLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, bets, gt_classes, params, need_logit_grad):
        x = logits.detach().to(torch.float32).contiguous()
        b = bets.detach().to(torch.float32).contiguous()
        stats = ops.loss_prepass(gt_classes, None, b, params.num_classes, params.temperature)
        out = ops.loss_main(x, gt_classes, params, stats, bets=b, want_grad_logits=need_logit_grad,
                            want_weights=True)
        gb = ops.loss_post(b, None, out["per_anchor_loss"], params, stats, out["scalars"])
        ctx.save_for_backward(out.get("grad_logits"), gb)
        ctx.mark_non_differentiable(out["per_anchor_loss"], out["weights"], out["scalars"], stats)
        return out["scalars"][7].to(torch.float32), out["per_anchor_loss"], out["weights"], out["scalars"], stats

    @staticmethod
    def backward(ctx, g, *unused):
        ops.single_use(ctx)
        gl, gb = ctx.saved_tensors
        if gl is not None:
            ops.scale_(gl, g)
        ops.scale_(gb, g)
        return gl, gb, None, None, None


class _GamblerLossLevelsFn(torch.autograd.Function):
    """gambler_loss with the logits left in the head's native per-level layout (fsg_loss_main_levels); the
    (N,R)-sized bets are flat.  Inputs: bets (N,R) then the L logit levels."""

    @staticmethod
    def forward(ctx, bets, gt_classes, params, need_logit_grad, *levels):
        xs = [t.detach() for t in levels]
        b = bets.detach().to(torch.float32).contiguous()
        stats = ops.loss_prepass(gt_classes, None, b, params.num_classes, params.temperature)
        out = ops.loss_main_levels(xs, gt_classes, params, stats, bets=b, want_grad_logits=need_logit_grad,
                                   want_weights=True)
        gb = ops.loss_post(b, None, out["per_anchor_loss"], params, stats, out["scalars"])
        ctx.n_levels = len(xs) if need_logit_grad else 0
        ctx.total_levels = len(xs)
        ctx.save_for_backward(gb, *(out["grad_logits"] if need_logit_grad else []))
        ctx.mark_non_differentiable(out["per_anchor_loss"], out["weights"], out["scalars"], stats)
        return out["scalars"][7].to(torch.float32), out["per_anchor_loss"], out["weights"], out["scalars"], stats

    @staticmethod
    def backward(ctx, g, *unused):
        ops.single_use(ctx)
        saved = ctx.saved_tensors
        gb, gls = saved[0], saved[1:]
        for t in gls:
            ops.scale_(t, g)
        ops.scale_(gb, g)
        gl_out = tuple(gls) if ctx.n_levels else (None,) * ctx.total_levels
        return (gb, None, None, None) + gl_out


class GamblerLoss:
    """Stand-in for the loss side of ``LayeredUnetGambler`` (the U-Net itself is out of scope)."""

    def __init__(self, num_classes=80, mode="focal", alpha=0.25, focal_gamma=2.0, normalize_w=True,
                 gambler_output="L_BAHW", gamma=1.0, temperature=0.1, kappa=1.0, num_scale=3,
                 event_storage=None, native_layout=True):
        self.native_layout = native_layout   # read the (N, A*K, H, W) logits in place (no permute/cat copy)
        assert gambler_output in ("L_BAHW", "L_B1HW", "L_BAHW_extendtobatch"), "does not support other shapes!"
        if gambler_output == "L_B1HW":
            raise NotImplementedError("L_B1HW raises in the reference as well (mask broadcast, gambler_heads.py:568)")
        self.num_classes, self.mode, self.alpha, self.focal_gamma = num_classes, mode, alpha, focal_gamma
        self.normalize_w, self.gambler_output, self.gamma = normalize_w, gambler_output, gamma
        self.temperature, self.kappa, self.num_scale = temperature, kappa, num_scale
        self.event_storage = event_storage

    @classmethod
    def from_config(cls, cfg, event_storage=None):
        g = cfg.MODEL.GAMBLER_HEAD
        return cls(g.NUM_CLASSES, g.GAMBLER_LOSS_MODE, cfg.MODEL.RETINANET.FOCAL_LOSS_ALPHA,
                   cfg.MODEL.RETINANET.FOCAL_LOSS_GAMMA, g.NORMALIZE, g.GAMBLER_OUTPUT, g.GAMBLER_GAMMA,
                   g.GAMBLER_TEMPERATURE, g.GAMBLER_KAPPA, len(cfg.MODEL.ANCHOR_GENERATOR.SIZES[0]), event_storage)

    def _norm_mode(self):
        if not self.normalize_w:
            return _lib.NORM_NONE
        return _lib.NORM_BATCH if self.gambler_output == "L_BAHW_extendtobatch" else _lib.NORM_IMAGE

    def gambler_loss(self, pred_class_logits, weights, gt_classes, mask, detach_pred=False):
        """pred_class_logits: list[(N, A*K, H, W)]; weights: list[(N, A, H, W)] betting maps (mutated in
        place to ``bets * mask`` like the reference, :560-569); gt_classes, mask: (N, R).
        -> ({"NAKHW_loss", "loss_before_weighting", "gambler_loss"}, normalised weights (N*R, 1))."""
        N = pred_class_logits[0].shape[0]
        hw = [(p.shape[2], p.shape[3]) for p in pred_class_logits]
        A = self.num_scale
        if detach_pred:
            pred_class_logits = [p.detach() for p in pred_class_logits]
        # mask -> per-level (N,A,H,W) and multiply into the caller's list (reference quirk kept)
        off = 0
        for i, (H, W) in enumerate(hw):
            n = H * W * A
            m_l = mask[:, off:off + n].reshape(N, H, W, A).permute(0, 3, 1, 2)
            weights[i] = weights[i] * m_l
            off += n
        b = levels_to_flat(list(weights), 1).reshape(N, -1)
        params = ops.make_loss_params(self.num_classes, self.alpha, self.focal_gamma, 0.1, self.temperature,
                                      self.gamma, self.mode, self._norm_mode(), 0.0, 0.0, 1.0)
        if self.native_layout:
            from .retinanet import _as_f32c
            G, ell, w_hat, scalars, stats = _GamblerLossLevelsFn.apply(
                b, gt_classes.contiguous(), params, not detach_pred, *[_as_f32c(p) for p in pred_class_logits])
        else:
            x = levels_to_flat(list(pred_class_logits), self.num_classes)
            G, ell, w_hat, scalars, stats = _GamblerLossFn.apply(x, b, gt_classes.contiguous(), params,
                                                                 not detach_pred)
        nakhw, off = [], 0
        for (H, W) in hw:
            n = H * W * A
            nakhw.append(ell[:, off:off + n].reshape(N, H, W, A).permute(0, 3, 1, 2))
            off += n
        R = ell.shape[1]
        w_max = (1 + self.temperature) / (R * self.temperature + 1)
        self.last_lower_bound = -(self.kappa * w_max * N) * scalars[4]
        if self.event_storage is not None:
            self.event_storage.put_scalar("loss_gambler/lower_bound", self.last_lower_bound)
        if self.mode == "focal":
            lbw = scalars[3] / torch.clamp(stats[0], min=1.0)
        elif self.mode == "sigmoid":
            lbw = scalars[3] / float(N * R)
        else:
            raise Exception("No mode it selected for the retinanet loss!!")
        loss_dict = {"NAKHW_loss": nakhw, "loss_before_weighting": lbw.to(torch.float32), "gambler_loss": G}
        return loss_dict, w_hat.reshape(-1, 1)
