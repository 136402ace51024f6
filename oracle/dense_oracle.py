"""CPU oracle for the per-anchor dense-detection hot path.  TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch-CPU fp32 tensor arithmetic (the same host library the
reference itself computes with, so integer outputs are bit-identical and float outputs
share the reference's rounding sequence), of the reference algorithm on the path that
``BASELINE.json:north_star`` names.  Every function cites the reference lines it follows
(paths relative to ``/root/reference``).

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or the timed
CPU baseline, never as the product.  The product package
(``full_scale_gambler_for_object_detection_b200``) never imports it and has no CPU path.

Parity pin: PINNED.  ``oracle/make_golden.py`` runs the *reference's own source files*
(through ``oracle/ref_loader.py``, in the build container) and this oracle on the same
seeded inputs and stores the reference's outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures everywhere, and
replays the reference's known-answer tests (tests/test_boxes.py:36-59,
tests/test_box2box_transform.py:16-30, tests/test_anchor_generator.py:14-43).
Third-party arithmetic that is not under /root/reference: ``fvcore.nn.smooth_l1_loss`` /
``sigmoid_focal_loss_jit`` (fvcore, unpinned git HEAD, INSTALL.md:15) -- the focal formula
is restated in-tree (retinanet.py:283-307, gambler_heads.py:104-128) and followed here;
smooth-L1 follows fvcore's published definition; ``torchvision.ops.nms`` (unpinned,
0.26.0 installed) -- greedy NMS restated below and pinned against the installed op.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SCALE_CLAMP = math.log(1000.0 / 16)  # box_regression.py:8


# ----------------------------------------------------------------------------------------
# geometry / matching
# ----------------------------------------------------------------------------------------
def box_area(b):
    """boxes.py:111-120."""
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def pairwise_iou(b1, b2):
    """boxes.py:243-275.  b1 (N,4), b2 (M,4) fp32 XYXY -> (N,M) fp32; 0 where inter <= 0."""
    b1 = b1.to(torch.float32).reshape(-1, 4)
    b2 = b2.to(torch.float32).reshape(-1, 4)
    a1, a2 = box_area(b1), box_area(b2)
    hi = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    lo = torch.max(b1[:, None, :2], b2[None, :, :2])
    wh = (hi - lo).clamp_(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[:, None] + a2[None, :] - inter  # (a1 + a2) - inter, boxes.py:272
    return torch.where(inter > 0, inter / union, torch.zeros((), dtype=inter.dtype))


def matcher(mqm, thresholds, labels, allow_low_quality_matches):
    """matcher.py:55-132.  mqm (M,N) -> matches int64 (N), match_labels int8 (N).

    Bands are [-inf, t0, .., +inf) with ``low <= v < high`` (:90-92), compared in fp32.
    Low-quality pass (:99-132) sets label 1 for every prediction whose quality equals a
    ground-truth row's maximum (ties included); it never changes ``matches``.
    """
    assert mqm.dim() == 2
    M, N = mqm.shape
    if mqm.numel() == 0:  # :70-80
        return (torch.zeros(N, dtype=torch.int64), torch.full((N,), labels[0], dtype=torch.int8))
    assert bool((mqm >= 0).all())  # :82
    vals, idx = mqm.max(dim=0)
    out = torch.ones(N, dtype=torch.int8)
    edges = [-float("inf")] + list(thresholds) + [float("inf")]
    for lab, lo, hi in zip(labels, edges[:-1], edges[1:]):
        out[(vals >= lo) & (vals < hi)] = lab
    if allow_low_quality_matches:
        row_best = mqm.max(dim=1).values
        hit = (mqm == row_best[:, None]).any(dim=0)
        out[hit] = 1
    return idx, out


def get_deltas(src, tgt, weights=(1.0, 1.0, 1.0, 1.0)):
    """box_regression.py:34-67."""
    sw = src[:, 2] - src[:, 0]
    sh = src[:, 3] - src[:, 1]
    sx = src[:, 0] + 0.5 * sw
    sy = src[:, 1] + 0.5 * sh
    tw = tgt[:, 2] - tgt[:, 0]
    th = tgt[:, 3] - tgt[:, 1]
    tx = tgt[:, 0] + 0.5 * tw
    ty = tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    return torch.stack(
        (wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)), dim=1
    )


def apply_deltas(deltas, boxes, weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=SCALE_CLAMP):
    """box_regression.py:69-107.  deltas (n,4k), boxes (n,4) -> (n,4k)."""
    boxes = boxes.to(deltas.dtype)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=scale_clamp)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=scale_clamp)
    pcx = dx * w[:, None] + cx[:, None]
    pcy = dy * h[:, None] + cy[:, None]
    pw = torch.exp(dw) * w[:, None]
    ph = torch.exp(dh) * h[:, None]
    out = torch.zeros_like(deltas)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw
    out[:, 3::4] = pcy + 0.5 * ph
    return out


def ground_truth(anchors, gt_boxes, gt_classes, num_classes, thresholds=(0.4, 0.5), labels=(0, -1, 1),
                 picky_thresholds=(0.4, 0.9), weights=(1.0, 1.0, 1.0, 1.0)):
    """retinanet.py:309-429 (get_ground_truth + get_picky_ground_truth), per image.

    anchors: (R,4) shared or list of (R,4) per image; gt_boxes: list[(M_i,4)];
    gt_classes: list[int64 (M_i)].  Returns dict with
      gt_classes (N,R) int64 in {-1, 0..K-1, K}; gt_deltas (N,R,4) fp32;
      mask (N,R) int64 (1 iff picky label == 1; all K for an image without GT, :425);
      matches (N,R) int64; match_labels (N,R) int8; picky_labels (N,R) int8.
    """
    out = {k: [] for k in ("gt_classes", "gt_deltas", "mask", "matches", "match_labels", "picky_labels")}
    for i, (gb, gc) in enumerate(zip(gt_boxes, gt_classes)):
        a = anchors[i] if isinstance(anchors, (list, tuple)) else anchors
        gb = gb.to(torch.float32).reshape(-1, 4)
        q = pairwise_iou(gb, a)
        m, lab = matcher(q, list(thresholds), list(labels), True)
        _, plab = matcher(q, list(picky_thresholds), list(labels), True)
        if gb.shape[0] > 0:
            d = get_deltas(a, gb[m], weights)
            c = gc[m].clone()
            c[lab == 0] = num_classes
            c[lab == -1] = -1
            pm = gc[m].clone()  # :414-423 -- relabelled in this order
            pm[plab == 0] = 0
            pm[plab == 1] = 1
            pm[plab == -1] = 0
        else:
            c = torch.zeros_like(m) + num_classes
            d = torch.zeros_like(a)
            pm = torch.zeros_like(m) + num_classes  # quirk kept: retinanet.py:425
        for k, v in zip(out, (c, d, pm, m, lab, plab)):
            out[k].append(v)
    return {k: torch.stack(v) for k, v in out.items()}


# ----------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------
def cls_loss_elementwise(x, t, mode="focal", alpha=0.25, gamma=2.0):
    """gambler_heads.py:104-128 (== retinanet.py:283-307 == fvcore sigmoid_focal_loss)."""
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    if mode == "sigmoid":
        return ce
    p_t = p * t + (1 - p) * (1 - t)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    return loss


def smooth_l1(x, y, beta):
    """fvcore.nn.smooth_l1_loss (third party; call site retinanet.py:241-246), reduction none."""
    n = torch.abs(x - y)
    if beta < 1e-5:
        return n
    return torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)


def one_hot_targets(gt_classes_flat, num_classes, like):
    """retinanet.py:224-229."""
    fg = (gt_classes_flat >= 0) & (gt_classes_flat != num_classes)
    t = torch.zeros_like(like)
    t[fg, gt_classes_flat[fg]] = 1
    return t, fg


def retinanet_losses(gt_classes, gt_deltas, logits, pred_deltas, num_classes, alpha=0.25, gamma=2.0,
                     beta=0.1):
    """retinanet.py:201-248 on already-flattened predictions.

    logits (N,R,K), pred_deltas (N,R,4) (the (N, sum HWA, K) layout of :24-54).
    Returns (loss_cls, loss_box_reg, num_foreground)."""
    x = logits.reshape(-1, num_classes)
    d = pred_deltas.reshape(-1, 4)
    g = gt_classes.flatten()
    gd = gt_deltas.reshape(-1, 4)
    valid = g >= 0
    t, fg = one_hot_targets(g, num_classes, x)
    nf = fg.sum()
    loss_cls = cls_loss_elementwise(x[valid], t[valid], "focal", alpha, gamma).sum() / max(1, nf)
    loss_reg = smooth_l1(d[fg], gd[fg], beta).sum() / max(1, nf)
    return loss_cls, loss_reg, nf


def gambler_loss(logits, bets, gt_classes, mask, num_classes, temperature=0.1, normalize=True,
                 mode="focal", alpha=0.25, focal_gamma=2.0, gambler_gamma=1.0, output="L_BAHW",
                 kappa=1.0):
    """gambler_heads.py:502-602 + :131-253 (branches L_BAHW / L_BAHW_extendtobatch) + :291-318,
    on flattened layouts: logits (N,R,K); bets, gt_classes, mask (N,R) with anchor index
    r = level_offset + (h*W + w)*A + a (the flattening of :34-48).

    Returns dict: gambler_loss (scalar, differentiable wrt logits and bets), per_anchor_loss
    l (N,R) detached (the NAKHW_loss values, :218), weights w_hat (N,R) detached,
    loss_before_weighting (detached scalar, :589-594), lower_bound (scalar, :17-31,583-587),
    num_foreground."""
    assert output in ("L_BAHW", "L_BAHW_extendtobatch")
    N, R, K = logits.shape
    x = logits.reshape(-1, K)
    g = gt_classes.flatten()
    valid = g >= 0
    t, fg = one_hot_targets(g, K, x)
    nf = fg.sum()
    cl = cls_loss_elementwise(x, t, mode, alpha, focal_gamma)
    vl = torch.zeros_like(cl)
    vl[valid, :] = cl[valid, :]  # :554-555
    w = bets * mask + temperature  # :568-569, :304
    if normalize:
        if output == "L_BAHW_extendtobatch":
            w = w / w.sum()  # :308-309
        else:
            w = w / w.sum(dim=1, keepdim=True)  # :311
    ell = vl.reshape(N, R, K).sum(dim=2)  # :218 (sum over classes)
    G = (-(w ** gambler_gamma) * ell).sum()  # :250-251
    ell_d = ell.detach()
    if mode == "focal":
        lbw = ell_d.sum() / max(1, nf)
    else:
        lbw = ell_d.sum() / (N * R)
    w_max = (1 + temperature) / (R * temperature + 1)  # :30
    lower = -(kappa * w_max * N * ell_d.max(dim=1).values.sum())
    return {
        "gambler_loss": G,
        "per_anchor_loss": ell_d,
        "weights": w.detach(),
        "loss_before_weighting": lbw,
        "lower_bound": lower,
        "num_foreground": nf,
    }


def train_step(anchors, gt_boxes, gt_classes_list, logits, pred_deltas, bets, num_classes,
               c_cls=1.0, c_reg=1.0, c_gam=-1.0, temperature=0.1, normalize=True, mode="focal",
               alpha=0.25, focal_gamma=2.0, gambler_gamma=1.0, beta=0.1, output="L_BAHW",
               detach_pred=False, need_grad=True):
    """The whole K1+K2 step the benchmark times: GT assignment with both matchers, the two
    RetinaNet losses, the gambler loss, and the backward of
    ``c_cls*loss_cls + c_reg*loss_box_reg + c_gam*gambler_loss``
    (train_net.py:1089-1098 with c = (1, lambda_reg, -lambda_out*kappa))."""
    gt = ground_truth(anchors, gt_boxes, gt_classes_list, num_classes)
    x = logits.detach().clone().requires_grad_(need_grad)
    d = pred_deltas.detach().clone().requires_grad_(need_grad)
    b = bets.detach().clone().requires_grad_(need_grad)
    lc, lr, nf = retinanet_losses(gt["gt_classes"], gt["gt_deltas"], x, d, num_classes, alpha, focal_gamma, beta)
    gl = gambler_loss(x.detach() if detach_pred else x, b, gt["gt_classes"], gt["mask"], num_classes,
                      temperature, normalize, mode, alpha, focal_gamma, gambler_gamma, output)
    total = c_cls * lc + c_reg * lr + c_gam * gl["gambler_loss"]
    out = dict(gt)
    out.update(loss_cls=lc.detach(), loss_box_reg=lr.detach(), gambler_loss=gl["gambler_loss"].detach(),
               total=total.detach(), per_anchor_loss=gl["per_anchor_loss"], weights=gl["weights"],
               loss_before_weighting=gl["loss_before_weighting"], lower_bound=gl["lower_bound"],
               num_foreground=nf)
    if need_grad:
        total.backward()
        zero = lambda v: torch.zeros_like(v)
        out.update(grad_logits=x.grad if x.grad is not None else zero(x),
                   grad_deltas=d.grad if d.grad is not None else zero(d),
                   grad_bets=b.grad if b.grad is not None else zero(b))
    return out


def calc_log_metrics(masked_bet_levels, weights, loss_cls, loss_box_reg, gambler_loss_value, loss_before_weighting,
                     lambda_reg=1.0, kappa=1.0, lambda_out=1.0, mode="cls+reg-gambler"):
    """GANTrainer.calc_log_metrics (ImbalanceDetection/train_net.py:1089-1124).  ``masked_bet_levels``: the betting
    maps as gambler_loss left them (multiplied by the picky mask in place, gambler_heads.py:568-569);
    ``weights``: the normalised weights it returned.  Returns the reference's loss_dict entries."""
    d = {"loss_cls": loss_cls, "loss_box_reg": loss_box_reg * lambda_reg, "loss_gambler": gambler_loss_value * kappa,
         "loss_before_weighting": loss_before_weighting}
    if mode == "cls+reg-gambler":
        d["loss_detector"] = d["loss_box_reg"] + d["loss_cls"] - lambda_out * d["loss_gambler"]
    else:
        d["loss_detector"] = d["loss_box_reg"] - lambda_out * d["loss_gambler"]
    s, mx, n = 0, 0, 0
    for b in masked_bet_levels:          # :1104-1111
        s = s + torch.sum(b)
        n = n + torch.numel(b)
        if torch.max(b) > mx:
            mx = torch.max(b)
    d["gambler_bets/sum"], d["gambler_bets/max"], d["gambler_bets/mean"] = s, mx, s / n
    d["visualized weights/sum"] = torch.sum(weights)
    d["visualized weights/max"] = torch.max(weights)
    d["visualized weights/mean"] = torch.mean(weights)
    d["visualized weights/median"] = torch.median(weights)   # lower median
    return d


# ----------------------------------------------------------------------------------------
# layout helpers (retinanet.py:24-54, gambler_heads.py:34-101)
# ----------------------------------------------------------------------------------------
def nchw_to_n_hwa_k(t, K):
    """(N, A*K, H, W) -> (N, H*W*A, K); anchor index (h*W+w)*A+a, channel a*K+k."""
    N, _, H, W = t.shape
    return t.view(N, -1, K, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, K)


def levels_to_flat(levels, K):
    return torch.cat([nchw_to_n_hwa_k(x, K) for x in levels], dim=1)


def flat_to_nahw(flat, hw_list, A):
    """(N,R) per-anchor values -> list[(N,A,H,W)] (the NAKHW_loss layout, gambler_heads.py:91-101)."""
    N = flat.shape[0]
    out, off = [], 0
    for H, W in hw_list:
        n = H * W * A
        out.append(flat[:, off:off + n].reshape(N, H, W, A).permute(0, 3, 1, 2))
        off += n
    return out


# ----------------------------------------------------------------------------------------
# anchors (anchor_generator.py:121-168)
# ----------------------------------------------------------------------------------------
def cell_anchors(sizes, aspect_ratios):
    out = []
    for s in sizes:
        area = s ** 2.0
        for ar in aspect_ratios:
            w = math.sqrt(area / ar)
            h = ar * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out)


def grid_anchors(grid_sizes, strides, sizes, aspect_ratios):
    """list of (H*W*A, 4) per level; order (h, w, a)."""
    res = []
    for (H, W), stride, sz, ar in zip(grid_sizes, strides, sizes, aspect_ratios):
        base = cell_anchors(sz, ar)
        sx = torch.arange(0, W * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(0, H * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        res.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return res


# ----------------------------------------------------------------------------------------
# detector_postprocess (modeling/postprocessing.py:8-52 with Boxes.scale/clip/nonempty,
# structures/boxes.py:122-151,205-210) for an Instances carrying pred_boxes/scores/pred_classes
# ----------------------------------------------------------------------------------------
def detector_postprocess(boxes, scores, classes, image_size, output_height, output_width):
    """-> (boxes, scores, classes) at the output resolution; empty boxes dropped (stable)."""
    scale_x, scale_y = (output_width / image_size[1], output_height / image_size[0])
    b = boxes.clone().to(torch.float32)
    b[:, 0::2] *= scale_x
    b[:, 1::2] *= scale_y
    h, w = output_height, output_width
    b[:, 0].clamp_(min=0, max=w)
    b[:, 1].clamp_(min=0, max=h)
    b[:, 2].clamp_(min=0, max=w)
    b[:, 3].clamp_(min=0, max=h)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return b[keep], scores[keep], classes[keep]


# ----------------------------------------------------------------------------------------
# inference: decode + top-k + NMS
# ----------------------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold):
    """Greedy NMS as torchvision.ops.nms computes it on CPU (third party; call sites
    detectron2/layers/nms.py:6,22): stable score-descending order; suppress j when
    inter/(area_i+area_j-inter) > thr (fp32 value promoted and compared against the double
    threshold); zero-area pairs give NaN and never suppress.  Returns int64 keep indices
    in score-descending order."""
    b = boxes.detach().to(torch.float32).numpy()
    s = scores.detach().to(torch.float32).numpy()
    n = b.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    with np.errstate(invalid="ignore", divide="ignore"):
        for pos in range(n):
            i = order[pos]
            if dead[i]:
                continue
            keep.append(i)
            rest = order[pos + 1:]
            w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            dead[rest[ovr.astype(np.float64) > thr]] = True
    return torch.as_tensor(np.asarray(keep, dtype=np.int64))


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2/layers/nms.py:9-26 in its per-class form (:20-26; torchvision's
    ``_batched_nms_vanilla`` is the same algorithm): NMS inside each class, result sorted by
    score descending (stable: ties keep the lower index first)."""
    n = boxes.shape[0]
    keep_mask = torch.zeros(n, dtype=torch.bool)
    for c in torch.unique(idxs).tolist():
        sel = (idxs == c).nonzero().view(-1)
        k = nms(boxes[sel], scores[sel], iou_threshold)
        keep_mask[sel[k]] = True
    keep = keep_mask.nonzero().view(-1)
    order = torch.sort(scores[keep], descending=True, stable=True).indices
    return keep[order]


def inference_single_image(box_cls, box_delta, anchors, num_classes, score_threshold=0.05,
                           topk_candidates=1000, nms_threshold=0.5, max_detections=100,
                           weights=(1.0, 1.0, 1.0, 1.0)):
    """retinanet.py:460-520.  Per level: box_cls_i (HWA,K) logits, box_delta_i (HWA,4),
    anchors_i (HWA,4).  Returns (boxes (D,4), scores (D), classes (D) int64) plus the
    pre-NMS candidates (boxes, scores, classes) in the reference's concatenation order."""
    B, S, C = [], [], []
    for cls_i, reg_i, anc_i in zip(box_cls, box_delta, anchors):
        p = cls_i.flatten().sigmoid()
        k = min(topk_candidates, reg_i.size(0))
        prob, idx = torch.sort(p, descending=True, stable=True)
        prob, idx = prob[:k], idx[:k]
        keep = prob > score_threshold
        prob, idx = prob[keep], idx[keep]
        a_idx = idx // num_classes
        c_idx = idx % num_classes
        B.append(apply_deltas(reg_i[a_idx], anc_i[a_idx], weights))
        S.append(prob)
        C.append(c_idx)
    B, S, C = torch.cat(B), torch.cat(S), torch.cat(C)
    keep = batched_nms(B, S, C, nms_threshold)[:max_detections]
    return (B[keep], S[keep], C[keep]), (B, S, C), keep


# ----------------------------------------------------------------------------------------
# two-stage callers of the same ops
# ----------------------------------------------------------------------------------------
def find_top_rpn_proposals(proposals, logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_side_len):
    """proposal_generator/rpn_outputs.py:52-151.  proposals list[(N,S_l,4)], logits list[(N,S_l)].
    -> list over images of (boxes, logits, level ids).  Sort ties: lower index first (stable)."""
    N = logits[0].shape[0]
    tb, ts, lv = [], [], []
    for level_id, (p_i, l_i) in enumerate(zip(proposals, logits)):
        k = min(pre_nms_topk, l_i.shape[1])
        sl, idx = torch.sort(l_i, descending=True, dim=1, stable=True)
        idx = idx[:, :k]
        tb.append(torch.gather(p_i, 1, idx[:, :, None].expand(-1, -1, 4)))
        ts.append(sl[:, :k])
        lv.append(torch.full((k,), level_id, dtype=torch.int64))
    tb, ts, lv = torch.cat(tb, dim=1), torch.cat(ts, dim=1), torch.cat(lv)
    out = []
    for n, (h, w) in enumerate(image_sizes):
        b = tb[n].clone()
        b[:, 0].clamp_(min=0, max=w); b[:, 1].clamp_(min=0, max=h)
        b[:, 2].clamp_(min=0, max=w); b[:, 3].clamp_(min=0, max=h)
        keep = ((b[:, 2] - b[:, 0]) > min_box_side_len) & ((b[:, 3] - b[:, 1]) > min_box_side_len)
        b, sc, l = b[keep], ts[n][keep], lv[keep]
        k = batched_nms(b, sc, l, nms_thresh)[:post_nms_topk]
        out.append((b[k], sc[k], l[k]))
    return out


def rpn_ground_truth(anchors, gt_boxes, thresholds=(0.3, 0.7), labels=(0, -1, 1), weights=(1.0, 1.0, 1.0, 1.0)):
    """RPNOutputs._get_ground_truth (rpn_outputs.py:250-295, boundary_threshold < 0)."""
    L, D = [], []
    for g in gt_boxes:
        m, lab = matcher(pairwise_iou(g, anchors), list(thresholds), list(labels), True)
        L.append(lab)
        D.append(torch.zeros_like(anchors) if g.shape[0] == 0 else get_deltas(anchors, g[m], weights))
    return L, D


def label_proposals(proposals, gt_boxes, gt_classes, num_classes, thresholds=(0.5,), labels=(0, 1)):
    """roi_heads/roi_heads.py:233-246 + _sample_proposals relabelling (:178-186)."""
    m, lab = matcher(pairwise_iou(gt_boxes, proposals), list(thresholds), list(labels), False)
    if gt_classes.numel() > 0:
        c = gt_classes[m].clone()
        c[lab == 0] = num_classes
        c[lab == -1] = -1
    else:
        c = torch.zeros_like(m) + num_classes
    return m, lab, c


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """roi_heads/fast_rcnn.py:76-118 -> (boxes, scores, classes, kept row indices)."""
    scores = scores[:, :-1]
    C = boxes.shape[1] // 4
    b = boxes.reshape(-1, 4).clone()
    h, w = image_shape
    b[:, 0].clamp_(min=0, max=w); b[:, 1].clamp_(min=0, max=h)
    b[:, 2].clamp_(min=0, max=w); b[:, 3].clamp_(min=0, max=h)
    b = b.view(-1, C, 4)
    mask = scores > score_thresh
    inds = mask.nonzero()
    b = b[inds[:, 0], 0] if C == 1 else b[mask]
    sc = scores[mask]
    keep = batched_nms(b, sc, inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    return b[keep], sc[keep], inds[keep, 1], inds[keep, 0]
