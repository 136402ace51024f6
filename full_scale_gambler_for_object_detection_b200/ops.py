"""Functional wrappers, one per C-ABI entry point of ``libfsg_dense.so``.

Everything here takes and returns CUDA tensors; torch is used only to allocate outputs/workspaces and to
supply the current stream.  The reference-shaped classes (``Matcher``, ``Box2BoxTransform``, ...) and the
fused training / inference steps are thin layers over these functions.
"""
import math

import torch

from . import _lib
from ._lib import LossParams, check, count_launches, host_f32, host_i8, host_i64, lib, ptr, stream

SCALE_CLAMP = math.log(1000.0 / 16)  # box_regression.py:8


def _f32c(t):
    return t.to(torch.float32).contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
def pairwise_iou(boxes1, boxes2):
    """(n1,4),(n2,4) fp32 XYXY -> (n1,n2) IoU (boxes.py:243-275)."""
    b1, b2 = _f32c(boxes1).reshape(-1, 4), _f32c(boxes2).reshape(-1, 4)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    if out.numel():
        check(lib().fsg_pairwise_iou(ptr(b1), b1.shape[0], ptr(b2), b2.shape[0], ptr(out), stream()))
        count_launches(1)
    return out


def matcher(mqm, thresholds, labels, allow_low_quality_matches):
    """matcher.py:55-132 on a materialised (M,N) matrix -> (matches int64 (N), match_labels int8 (N))."""
    assert mqm.dim() == 2
    q = _f32c(mqm)
    M, N = q.shape
    matches = torch.empty(N, dtype=torch.int64, device=q.device)
    out_labels = torch.empty(N, dtype=torch.int8, device=q.device)
    if N == 0:
        return matches, out_labels
    rowmax = torch.empty(max(M, 1), dtype=torch.float32, device=q.device)
    check(lib().fsg_matcher(ptr(q) if M > 0 else None, M, N, host_f32(thresholds), host_i8(labels), len(thresholds),
                            int(bool(allow_low_quality_matches)), ptr(matches), ptr(out_labels), ptr(rowmax),
                            stream()))
    count_launches(2 if (M > 0 and allow_low_quality_matches) else 1)
    return matches, out_labels


class PackedGT:
    """Ground truth of a batch in the packed layout the fused kernels read: boxes (sum_M,4) fp32, classes
    (sum_M) int64, offsets (N+1) int32 (device) + the same offsets on the host."""

    def __init__(self, boxes, classes, offsets_dev, offsets_host):
        self.boxes, self.classes, self.offsets, self.offsets_host = boxes, classes, offsets_dev, offsets_host

    @property
    def num_images(self):
        return len(self.offsets_host) - 1

    @property
    def total(self):
        return int(self.offsets_host[-1])

    @staticmethod
    def from_lists(gt_boxes, gt_classes, device):
        """gt_boxes: list of (M_i,4); gt_classes: list of (M_i) int64.  One H2D copy per array."""
        counts = [int(b.shape[0]) for b in gt_boxes]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        if offs[-1] > 0:
            boxes = torch.cat([b.reshape(-1, 4).to(torch.float32) for b in gt_boxes]).to(device).contiguous()
            classes = torch.cat([c.reshape(-1).to(torch.int64) for c in gt_classes]).to(device).contiguous()
        else:
            boxes = torch.zeros((1, 4), dtype=torch.float32, device=device)
            classes = torch.zeros((1,), dtype=torch.int64, device=device)
        offsets = torch.tensor(offs, dtype=torch.int32).to(device)
        return PackedGT(boxes, classes, offsets, offs)


def match_anchors(anchors, gt, num_classes, thresholds=(0.4, 0.5), labels=(0, -1, 1),
                  picky_thresholds=(0.4, 0.9), picky_labels=None, box_weights=(1.0, 1.0, 1.0, 1.0),
                  want=("gt_classes", "mask", "matched_idx32"), bets=None, temperature=0.0,
                  allow_low_quality_matches=True, bet_levels=None, phases=3, workspace=None, out=None, peer=None):
    """Fused IoU + Matcher(s) + GT assignment (retinanet.py:339-363, 400-425) for a batch.

    anchors: (R,4) shared by all images or (N,R,4) per image.  gt: PackedGT.
    want: subset of {matches, match_labels, picky_labels, gt_classes, mask, gt_deltas, matched_idx32}.
    bets (N,R): when given, the loss pre-pass is fused in and ``stats`` is returned too.
    bet_levels list[(N, A, H, W)]: the same with the betting maps read in their own layout (no flattened copy).
    peer: ``sharded.PeerExchange`` -- the all-reduce of [num_foreground, S_batch] over the ranks then happens inside
    the second kernel over NVLink peer memory (every rank must enqueue the same call).
    phases / workspace / out: the two-phase form for anchors sharded by range over ranks (see
    ``sharded.match_anchor_range``): phases=1 runs pass A only, phases=2 pass B on the same ``workspace`` after the
    per-GT maxima in it (``out["gt_max_bits"]``, an int32 view) were all-reduced with MAX.
    Returns a dict of the requested (N,R[,4]) tensors (+ "stats").
    """
    a = _f32c(anchors)
    dev = a.device
    N = gt.num_images
    if a.dim() == 3:
        assert a.shape[0] == N
        R, stride = a.shape[1], a.shape[1] * 4
    else:
        R, stride = a.shape[0], 0
    if picky_thresholds is not None and picky_labels is None:
        picky_labels = labels
    kinds = {
        "matches": (torch.int64, ()), "match_labels": (torch.int8, ()), "picky_labels": (torch.int8, ()),
        "gt_classes": (torch.int64, ()), "mask": (torch.int64, ()), "gt_deltas": (torch.float32, (4,)),
        "matched_idx32": (torch.int32, ()),
    }
    if out is None:
        out = {}
        for k in want:
            dt, tail = kinds[k]
            out[k] = torch.empty((N, R) + tail, dtype=dt, device=dev)
    stats = out.get("stats")
    if bets is not None:
        bets = _f32c(bets)
        assert bets.shape == (N, R)
    lv = None
    if bet_levels is not None:
        assert bets is None
        lv = bet_levels_struct(bet_levels)
        assert sum(b.shape[1] * b.shape[2] * b.shape[3] for b in bet_levels) == R and bet_levels[0].shape[0] == N
    if stats is None and (bets is not None or lv is not None or "stats" in want):
        stats = torch.empty(_lib.STATS_HEADER + N, dtype=torch.float64, device=dev)
    L = lib()
    ws = workspace if workspace is not None else _ws(L.fsg_match_workspace_bytes(N, R, gt.total), dev)
    if R > 0:
        check(L.fsg_match_anchors_ex(
            ptr(a), R, stride, ptr(gt.boxes), ptr(gt.classes), ptr(gt.offsets), N, gt.total, int(num_classes),
            host_f32(thresholds), host_i8(labels), len(thresholds), int(bool(allow_low_quality_matches)),
            host_f32(picky_thresholds) if picky_thresholds is not None else None,
            host_i8(picky_labels) if picky_thresholds is not None else None,
            len(picky_thresholds) if picky_thresholds is not None else 0,
            host_f32(box_weights), ptr(out.get("matches")), ptr(out.get("match_labels")),
            ptr(out.get("picky_labels")), ptr(out.get("gt_classes")), ptr(out.get("mask")),
            ptr(out.get("gt_deltas")), ptr(out.get("matched_idx32")), ptr(bets), lv, float(temperature), ptr(stats),
            peer.ctx if peer is not None else None, ptr(ws), ws.numel(), int(phases), stream()))
        count_launches(3 if phases == 3 else (1 if phases == 1 else 2))   # pass A, pass B (patch), fold
    if stats is not None:
        out["stats"] = stats
    if phases != 3:
        off = L.fsg_match_gt_max_offset(N, R, gt.total)
        out["workspace"] = ws
        out["gt_max_bits"] = ws[off:off + 4 * max(gt.total, 1)].view(torch.int32)   # fp32 bit patterns, all >= 0
    return out


def bet_levels_struct(bet_levels):
    """list[(N, A, H, W)] contiguous fp32 CUDA tensors -> ``fsg_bet_levels`` (keeps no reference: the caller must)."""
    lv = _lib.BetLevels()
    lv.num_levels, lv.A = len(bet_levels), bet_levels[0].shape[1]
    for i, b in enumerate(bet_levels):
        assert b.dtype == torch.float32 and b.shape[1] == lv.A
        lv.bets[i] = ptr(b)
        lv.H[i], lv.W[i] = b.shape[2], b.shape[3]
    return lv


def get_deltas(src_boxes, target_boxes, weights):
    s, t = _f32c(src_boxes), _f32c(target_boxes)
    assert s.shape == t.shape and s.shape[-1] == 4
    out = torch.empty_like(s)
    n = s.shape[0]
    if n:
        check(lib().fsg_box2box_get_deltas(ptr(s), ptr(t), n, host_f32(weights), ptr(out), stream()))
        count_launches(1)
    return out


def apply_deltas(deltas, boxes, weights, scale_clamp=SCALE_CLAMP):
    d, b = _f32c(deltas), _f32c(boxes)
    n = b.shape[0]
    assert d.shape[0] == n and d.shape[1] % 4 == 0
    out = torch.empty_like(d)
    if d.numel():
        check(lib().fsg_box2box_apply_deltas(ptr(d), ptr(b), n, d.shape[1] // 4, host_f32(weights),
                                             float(scale_clamp), ptr(out), stream()))
        count_launches(1)
    return out


# ------------------------------------------------------------------------------------------------
# layout adapter
# ------------------------------------------------------------------------------------------------
def levels_to_flat(levels, K, out=None):
    """list[(N, A*K, H, W)] -> (N, sum HWA, K) (retinanet.py:24-54).  One transpose launch per level."""
    N = levels[0].shape[0]
    dev = levels[0].device
    rows = [x.shape[1] // K * x.shape[2] * x.shape[3] for x in levels]
    R = sum(rows)
    if out is None:
        out = torch.empty((N, R, K), dtype=torch.float32, device=dev)
    off = 0
    for x, r in zip(levels, rows):
        x = _f32c(x)
        C, HW = x.shape[1], x.shape[2] * x.shape[3]
        check(lib().fsg_permute_level(ptr(x), ptr(out), N, C, HW, R * K, off * K, 0, stream()))
        off += r
    count_launches(len(levels))
    return out


def flat_to_levels(flat, shapes):
    """(N, R, K) -> list[(N, A*K, H, W)] for ``shapes`` = [(C, H, W), ...] (the gradient direction)."""
    flat = _f32c(flat)
    N, R, K = flat.shape
    outs, off = [], 0
    for C, H, W in shapes:
        o = torch.empty((N, C, H, W), dtype=torch.float32, device=flat.device)
        check(lib().fsg_permute_level(ptr(o), ptr(flat), N, C, H * W, R * K, off * K, 1, stream()))
        off += (C // K) * H * W
        outs.append(o)
    count_launches(len(shapes))
    return outs


def anchor_maps_to_flat(level_lists, out=None):
    """Up to three lists of per-level (N, A, H, W) maps -> list of flat (N, R) tensors, one launch
    (the betting maps' permute_all_weights_to_N_HWA_K_and_concat_, gambler_heads.py:291-318)."""
    return _anchor_maps(level_lists, out, 0)


def anchor_maps_to_levels(flats, shapes, out=None):
    """Up to three flat (N, R) tensors -> lists of per-level (N, A, H, W) maps for ``shapes`` = [(A, H, W)],
    one launch (NAKHW_loss / d-bets in the gambler's layout, gambler_heads.py:91-101)."""
    N = flats[0].shape[0]
    dev = flats[0].device
    if out is None:
        out = [[torch.empty((N, A, H, W), dtype=torch.float32, device=dev) for (A, H, W) in shapes] for _ in flats]
    _anchor_maps(out, list(flats), 1)
    return out


def _anchor_maps(level_lists, flats, to_levels):
    nt, nl = len(level_lists), len(level_lists[0])
    first = level_lists[0]
    N, A = first[0].shape[0], first[0].shape[1]
    R = sum(A * x.shape[2] * x.shape[3] for x in first)
    dev = first[0].device
    if flats is None:
        flats = [torch.empty((N, R), dtype=torch.float32, device=dev) for _ in range(nt)]
    lp = (_lib.c_ptr * (nt * nl))()
    fp = (_lib.c_ptr * nt)()
    for t in range(nt):
        assert flats[t].shape == (N, R) and flats[t].dtype == torch.float32
        fp[t] = ptr(flats[t])
        for l in range(nl):
            x = level_lists[t][l]
            assert x.dtype == torch.float32 and x.shape == first[l].shape
            lp[t * nl + l] = ptr(x)
    hw = (_lib.c_i32 * nl)(*[x.shape[2] * x.shape[3] for x in first])
    check(lib().fsg_anchor_maps(lp, fp, nt, hw, nl, A, N, int(to_levels), stream()))
    count_launches(1)
    return flats


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def make_loss_params(num_classes, alpha=0.25, gamma=2.0, beta=0.1, temperature=0.1, gambler_gamma=1.0,
                     mode="focal", norm_mode=_lib.NORM_IMAGE, c_cls=1.0, c_reg=1.0, c_gam=0.0,
                     box_weights=(1.0, 1.0, 1.0, 1.0)):
    p = LossParams()
    p.num_classes = int(num_classes)
    p.gambler_mode = _lib.CLS_MODES[mode]
    p.norm_mode = int(norm_mode)
    p.focal_alpha, p.focal_gamma, p.smooth_l1_beta = float(alpha), float(gamma), float(beta)
    p.temperature, p.gambler_gamma = float(temperature), float(gambler_gamma)
    p.c_cls, p.c_reg, p.c_gam = float(c_cls), float(c_reg), float(c_gam)
    for i in range(4):
        p.box_weights[i] = float(box_weights[i])
    return p


def loss_prepass(gt_classes, mask, bets, num_classes, temperature):
    """stats = [num_foreground, S_batch, S[0..N-1]] (double), see include/fsg_dense.h."""
    N, R = gt_classes.shape
    dev = gt_classes.device
    stats = torch.empty(_lib.STATS_HEADER + N, dtype=torch.float64, device=dev)
    L = lib()
    ws = _ws(L.fsg_loss_prepass_workspace_bytes(N, R), dev)
    check(L.fsg_loss_prepass(ptr(gt_classes), ptr(mask), ptr(bets), N, R, int(num_classes), float(temperature),
                             ptr(stats), ptr(ws), ws.numel(), stream()))
    count_launches(1)
    return stats


def loss_main(logits, gt_classes, params, stats, pred_deltas=None, gt_deltas=None, anchors=None, gt=None,
              matched_idx32=None, mask=None, bets=None, want_grad_logits=True, want_grad_deltas=True,
              want_weights=False, grad_logits_out=None):
    """The fused main pass.  Returns dict(grad_logits, grad_deltas, per_anchor_loss, weights, scalars)."""
    x = logits
    assert x.dtype == torch.float32 and x.is_contiguous()
    N, R, K = x.shape
    dev = x.device
    assert K == params.num_classes
    out = {}
    if want_grad_logits:
        out["grad_logits"] = grad_logits_out if grad_logits_out is not None else torch.empty_like(x)
    if pred_deltas is not None and want_grad_deltas:
        out["grad_deltas"] = torch.empty_like(pred_deltas)
    out["per_anchor_loss"] = torch.empty((N, R), dtype=torch.float32, device=dev)
    if want_weights:
        out["weights"] = torch.empty((N, R), dtype=torch.float32, device=dev)
    scalars = torch.empty(_lib.SCALARS_HEADER + N, dtype=torch.float64, device=dev)
    out["scalars"] = scalars
    a_ptr, a_stride = None, 0
    if anchors is not None:
        a_ptr = ptr(anchors)
        a_stride = anchors.shape[1] * 4 if anchors.dim() == 3 else 0
    L = lib()
    ws = _ws(L.fsg_loss_main_workspace_bytes(N, R, K), dev)
    check(L.fsg_loss_main(
        ptr(x), ptr(pred_deltas), ptr(gt_deltas), a_ptr, a_stride, ptr(gt.boxes) if gt is not None else None,
        ptr(gt.offsets) if gt is not None else None, ptr(matched_idx32), ptr(gt_classes), ptr(mask), ptr(bets),
        N, R, params, ptr(stats), ptr(out.get("grad_logits")), ptr(out.get("grad_deltas")),
        ptr(out["per_anchor_loss"]), ptr(out.get("weights")), ptr(scalars), ptr(ws), ws.numel(), stream()))
    count_launches(1)
    return out


def loss_main_levels(logit_levels, gt_classes, params, stats, delta_levels=None, gt_deltas=None, anchors=None,
                     gt=None, matched_idx32=None, mask=None, bets=None, want_grad_logits=True,
                     want_grad_deltas=True, want_weights=False, grad_logits_out=None, grad_deltas_out=None,
                     bet_levels=None, ell_levels_out=None, scalars_out=None, workspace=None):
    """The fused main pass on the head's native layout: ``logit_levels`` list[(N, A*K, H, W)],
    ``delta_levels`` list[(N, A*4, H, W)]; gradients come back as lists of the same shapes.  The
    (N,R)-sized arguments are as in :func:`loss_main`.  No permute/cat copy of the logits is made.
    bet_levels list[(N, A, H, W)] (instead of flat ``bets``) and ell_levels_out (same shapes; the NAKHW_loss is then
    written there and ``per_anchor_loss`` in the result is that list) keep the gambler side in its own layout too."""
    K = params.num_classes
    N = logit_levels[0].shape[0]
    A = logit_levels[0].shape[1] // K
    dev = logit_levels[0].device
    nl = len(logit_levels)
    xs = [x if (x.dtype == torch.float32 and x.is_contiguous()) else _f32c(x) for x in logit_levels]
    ds = None
    if delta_levels is not None:
        ds = [d if (d.dtype == torch.float32 and d.is_contiguous()) else _f32c(d) for d in delta_levels]
    out = {}
    if want_grad_logits:
        out["grad_logits"] = grad_logits_out if grad_logits_out is not None else [torch.empty_like(x) for x in xs]
    if ds is not None and want_grad_deltas:
        out["grad_deltas"] = grad_deltas_out if grad_deltas_out is not None else [torch.empty_like(d) for d in ds]
    R = sum(x.shape[2] * x.shape[3] * A for x in xs)
    assert gt_classes.shape == (N, R), "gt_classes %s vs (N=%d, R=%d)" % (tuple(gt_classes.shape), N, R)
    levels = (_lib.HeadLevel * nl)()
    for i, x in enumerate(xs):
        assert x.shape[0] == N and x.shape[1] == A * K
        levels[i].logits = ptr(x)
        levels[i].grad_logits = ptr(out["grad_logits"][i]) if want_grad_logits else None
        levels[i].pred_deltas = ptr(ds[i]) if ds is not None else None
        levels[i].grad_deltas = ptr(out["grad_deltas"][i]) if "grad_deltas" in out else None
        levels[i].H, levels[i].W = x.shape[2], x.shape[3]
        if ds is not None:
            assert ds[i].shape == (N, A * 4, x.shape[2], x.shape[3])
        if bet_levels is not None:
            assert bets is None and bet_levels[i].shape == (N, A, x.shape[2], x.shape[3])
            levels[i].bets = ptr(bet_levels[i])
        if ell_levels_out is not None:
            assert ell_levels_out[i].shape == (N, A, x.shape[2], x.shape[3])
            levels[i].per_anchor_loss = ptr(ell_levels_out[i])
    out["per_anchor_loss"] = (ell_levels_out if ell_levels_out is not None
                              else torch.empty((N, R), dtype=torch.float32, device=dev))
    if want_weights:
        out["weights"] = torch.empty((N, R), dtype=torch.float32, device=dev)
    scalars = scalars_out if scalars_out is not None else torch.empty(_lib.SCALARS_HEADER + N, dtype=torch.float64,
                                                                      device=dev)
    out["scalars"] = scalars
    a_ptr, a_stride = None, 0
    if anchors is not None:
        a_ptr = ptr(anchors)
        a_stride = anchors.shape[1] * 4 if anchors.dim() == 3 else 0
    L = lib()
    ws = workspace if workspace is not None else _ws(L.fsg_loss_main_levels_workspace_bytes(N, levels, nl, A), dev)
    check(L.fsg_loss_main_levels(
        levels, nl, A, ptr(gt_deltas), a_ptr, a_stride, ptr(gt.boxes) if gt is not None else None,
        ptr(gt.offsets) if gt is not None else None, ptr(matched_idx32), ptr(gt_classes), ptr(mask), ptr(bets),
        N, R, params, ptr(stats), ptr(out["per_anchor_loss"]) if ell_levels_out is None else None,
        ptr(out.get("weights")), ptr(scalars), ptr(ws), ws.numel(), stream()))
    count_launches(1)
    return out


def loss_post_levels(bet_levels, mask, ell_levels, params, stats, scalars, out=None):
    """d(c_gam * gambler_loss)/d bets with everything in the gambler's per-level (N, A, H, W) layout; mask is the
    flat (N, R) K1 output.  -> list of gradients (same shapes)."""
    nl = len(bet_levels)
    N, A = bet_levels[0].shape[0], bet_levels[0].shape[1]
    if out is None:
        out = [torch.empty_like(b) for b in bet_levels]
    lv = (_lib.PostLevel * nl)()
    R = 0
    for i, (b, e, g) in enumerate(zip(bet_levels, ell_levels, out)):
        assert b.shape == e.shape == g.shape and b.dtype == torch.float32
        lv[i].bets, lv[i].per_anchor_loss, lv[i].grad_bets = ptr(b), ptr(e), ptr(g)
        lv[i].H, lv[i].W = b.shape[2], b.shape[3]
        R += A * b.shape[2] * b.shape[3]
    check(lib().fsg_loss_post_levels(lv, nl, A, ptr(mask), N, R, params, ptr(stats), ptr(scalars), stream()))
    count_launches(1)
    return out


def loss_post(bets, mask, per_anchor_loss, params, stats, scalars):
    N, R = bets.shape
    g = torch.empty_like(bets)
    check(lib().fsg_loss_post(ptr(bets), ptr(mask), ptr(per_anchor_loss), N, R, params, ptr(stats), ptr(scalars),
                              ptr(g), stream()))
    count_launches(1)
    return g


BET_STAT_NAMES = ("gambler_bets/sum", "gambler_bets/max", "gambler_bets/mean", "visualized weights/sum",
                  "visualized weights/max", "visualized weights/mean", "visualized weights/median")


def bet_stats(bets, mask, params, stats, bet_levels=None):
    """The bet / weight statistics of GANTrainer.calc_log_metrics (train_net.py:1104-1121) as a (8,) double device
    tensor (names: ``BET_STAT_NAMES``) -- no host sync, no sort.  bets (N,R) flat, or bet_levels list[(N, A, H, W)]
    (the UNMASKED maps; the picky ``mask`` (N,R) is applied here exactly like gambler_loss applies it)."""
    L = lib()
    lv = None
    if bet_levels is not None:
        assert bets is None
        lv = bet_levels_struct(bet_levels)
        N = bet_levels[0].shape[0]
        R = sum(b.shape[1] * b.shape[2] * b.shape[3] for b in bet_levels)
        dev = bet_levels[0].device
    else:
        N, R = bets.shape
        dev = bets.device
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    ws = _ws(L.fsg_bet_stats_workspace_bytes(), dev)
    check(L.fsg_bet_stats(ptr(bets), lv, ptr(mask), N, R, params, ptr(stats), ptr(out), ptr(ws), ws.numel(), stream()))
    count_launches(3)
    return out


def single_use(ctx):
    """The fused kernels write the gradients for a UNIT upstream gradient at forward time; backward rescales those
    buffers in place (through raw pointers, so autograd's version counters never see it; no extra pass over 344 MB
    when the upstream gradient is 1) and hands them to autograd.  A second backward through the same node
    (retain_graph=True, two losses sharing the node) would scale them twice, and with a plan they alias storage the
    next ``plan.run`` overwrites -- so every autograd node of this package is single-use and says so."""
    if getattr(ctx, "used", False):
        raise RuntimeError("fsg_dense: backward() ran twice through the same node; its gradient buffers are scaled in "
                           "place and handed out once (call the op again instead of retain_graph=True)")
    ctx.used = True


def scale_(x, scale, mul=1.0):
    """In-place x *= scale * mul; scale is a python float or a 0-d CUDA tensor (no host sync), mul a python float."""
    if isinstance(scale, torch.Tensor):
        s = scale.detach().to(torch.float32).reshape(1).contiguous()
        check(lib().fsg_scale_inplace(ptr(x), x.numel(), ptr(s), float(mul), stream()))
    else:
        if float(scale) * float(mul) == 1.0:
            return x
        check(lib().fsg_scale_inplace(ptr(x), x.numel(), None, float(scale) * float(mul), stream()))
    count_launches(1)
    return x


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
DETECT_LAUNCHES = 5        # sample bar, scan, finalize, streaming top-k (exact fallback, usually idle), NMS
NMS_SMEM_BOXES = 8192      # up to here one CTA (or 8, split by class) does the whole call in shared memory
NMS_MAX_BOXES = 262144     # beyond: rank / bit-matrix / sweep kernels (csrc/nms_large.cu), 8.6 GB mask at the cap


def nms_raw(boxes, scores, class_ids, iou_threshold):
    """Device-side result: (keep int64 (n) padded, num_keep int32 (1)).  n <= NMS_MAX_BOXES."""
    b, s = _f32c(boxes).reshape(-1, 4), _f32c(scores).reshape(-1)
    n = b.shape[0]
    if n > NMS_MAX_BOXES:
        raise RuntimeError("fsg_nms takes at most %d boxes per call; got %d" % (NMS_MAX_BOXES, n))
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=b.device)
    num = torch.zeros(1, dtype=torch.int32, device=b.device)
    c = class_ids.to(torch.int64).contiguous() if class_ids is not None else None
    L = lib()
    ws = _ws(L.fsg_nms_workspace_bytes(n), b.device)
    check(L.fsg_nms(ptr(b) if n else None, ptr(s) if n else None, ptr(c), n, float(iou_threshold), ptr(keep),
                    ptr(num), ptr(ws), ws.numel(), stream()))
    count_launches(1 if n <= NMS_SMEM_BOXES else 3)
    return keep, num


def detect(logits, deltas, anchors, level_offsets, score_threshold=0.05, topk=1000, nms_threshold=0.5,
           max_det=100, box_weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=SCALE_CLAMP, want_candidates=False,
           postprocess=None):
    """Batched RetinaNet.inference (retinanet.py:431-520) on flattened predictions.

    logits (N,R,K), deltas (N,R,4), anchors (R,4) or (N,R,4); level_offsets: list of L+1 anchor offsets.
    Returns dict(boxes (N,max_det,4), scores (N,max_det), classes (N,max_det) int64, count (N) int32
    [, cand_boxes, cand_scores, cand_classes, cand_count, keep_idx]).

    postprocess: optional (N,4) CUDA fp32 rows [scale_x, scale_y, clip_w, clip_h] -- detector_postprocess
    (modeling/postprocessing.py:8-52) fused into the NMS epilogue (see :func:`postprocess_rows`)."""
    x, d, a = logits, _f32c(deltas), _f32c(anchors)
    assert x.dtype == torch.float32 and x.is_contiguous()
    N, R, K = x.shape
    dev = x.device
    nl = len(level_offsets) - 1
    out = {
        "boxes": torch.empty((N, max_det, 4), dtype=torch.float32, device=dev),
        "scores": torch.empty((N, max_det), dtype=torch.float32, device=dev),
        "classes": torch.empty((N, max_det), dtype=torch.int64, device=dev),
        "count": torch.empty((N,), dtype=torch.int32, device=dev),
    }
    if want_candidates:
        cap = nl * topk
        out["cand_boxes"] = torch.zeros((N, cap, 4), dtype=torch.float32, device=dev)
        out["cand_scores"] = torch.zeros((N, cap), dtype=torch.float32, device=dev)
        out["cand_classes"] = torch.zeros((N, cap), dtype=torch.int64, device=dev)
        out["cand_count"] = torch.empty((N,), dtype=torch.int32, device=dev)
        out["keep_idx"] = torch.empty((N, max_det), dtype=torch.int64, device=dev)
    L = lib()
    ws = _ws(L.fsg_detect_workspace_bytes(N, R, K, nl, topk), dev)
    check(L.fsg_detect(
        ptr(x), ptr(d), ptr(a), a.shape[1] * 4 if a.dim() == 3 else 0, N, R, K, host_i64(level_offsets), nl,
        float(score_threshold), int(topk), float(nms_threshold), int(max_det), host_f32(box_weights),
        float(scale_clamp), ptr(out["boxes"]), ptr(out["scores"]), ptr(out["classes"]), ptr(out["count"]),
        ptr(out.get("cand_boxes")), ptr(out.get("cand_scores")), ptr(out.get("cand_classes")),
        ptr(out.get("cand_count")), ptr(out.get("keep_idx")), ptr(postprocess), ptr(ws), ws.numel(), stream()))
    count_launches(DETECT_LAUNCHES)
    if want_candidates:   # diagnostics: how each (image, level) slab was selected (see fsg_detect_status_offset)
        off = L.fsg_detect_status_offset(N, nl, int(topk))
        out["slab_status"] = ws[off:off + 4 * N * nl].view(torch.int32).view(N, nl)
    return out


def detect_levels(logit_levels, delta_levels, anchors, num_classes, score_threshold=0.05, topk=1000, nms_threshold=0.5,
                  max_det=100, box_weights=(1.0, 1.0, 1.0, 1.0), scale_clamp=SCALE_CLAMP, want_candidates=False,
                  postprocess=None):
    """:func:`detect` straight from the head outputs: ``logit_levels`` list[(N, A*K, H, W)], ``delta_levels``
    list[(N, A*4, H, W)] are read in place (no permute/cat copy, retinanet.py:444-447); anchors (R,4) or (N,R,4) in
    the reference's flattened order.  Same result dict as :func:`detect`, bit-identical to running it on the permuted
    copy."""
    K = int(num_classes)
    xs = [x if (x.dtype == torch.float32 and x.is_contiguous()) else _f32c(x) for x in logit_levels]
    ds = [d if (d.dtype == torch.float32 and d.is_contiguous()) else _f32c(d) for d in delta_levels]
    a = _f32c(anchors)
    N = xs[0].shape[0]
    A = xs[0].shape[1] // K
    dev = xs[0].device
    nl = len(xs)
    levels = (_lib.DetectLevel * nl)()
    R = 0
    for i, (x, d) in enumerate(zip(xs, ds)):
        assert x.shape[0] == N and x.shape[1] == A * K and d.shape == (N, A * 4, x.shape[2], x.shape[3])
        levels[i].logits, levels[i].deltas = ptr(x), ptr(d)
        levels[i].H, levels[i].W = x.shape[2], x.shape[3]
        R += A * x.shape[2] * x.shape[3]
    assert a.shape[-2] == R, "anchors (%d) do not match the levels (%d)" % (a.shape[-2], R)
    out = {
        "boxes": torch.empty((N, max_det, 4), dtype=torch.float32, device=dev),
        "scores": torch.empty((N, max_det), dtype=torch.float32, device=dev),
        "classes": torch.empty((N, max_det), dtype=torch.int64, device=dev),
        "count": torch.empty((N,), dtype=torch.int32, device=dev),
    }
    if want_candidates:
        cap = nl * topk
        out["cand_boxes"] = torch.zeros((N, cap, 4), dtype=torch.float32, device=dev)
        out["cand_scores"] = torch.zeros((N, cap), dtype=torch.float32, device=dev)
        out["cand_classes"] = torch.zeros((N, cap), dtype=torch.int64, device=dev)
        out["cand_count"] = torch.empty((N,), dtype=torch.int32, device=dev)
        out["keep_idx"] = torch.empty((N, max_det), dtype=torch.int64, device=dev)
    L = lib()
    ws = _ws(L.fsg_detect_workspace_bytes(N, R, K, nl, topk), dev)
    check(L.fsg_detect_levels(
        levels, nl, A, K, ptr(a), a.shape[1] * 4 if a.dim() == 3 else 0, N, R, float(score_threshold), int(topk),
        float(nms_threshold), int(max_det), host_f32(box_weights), float(scale_clamp), ptr(out["boxes"]),
        ptr(out["scores"]), ptr(out["classes"]), ptr(out["count"]), ptr(out.get("cand_boxes")),
        ptr(out.get("cand_scores")), ptr(out.get("cand_classes")), ptr(out.get("cand_count")),
        ptr(out.get("keep_idx")), ptr(postprocess), ptr(ws), ws.numel(), stream()))
    count_launches(DETECT_LAUNCHES)
    if want_candidates:   # diagnostics: how each (image, level) slab was selected (see fsg_detect_status_offset)
        off = L.fsg_detect_status_offset(N, nl, int(topk))
        out["slab_status"] = ws[off:off + 4 * N * nl].view(torch.int32).view(N, nl)
    return out


def postprocess_rows(image_sizes, output_sizes, device):
    """[(h, w)] network-input sizes and [(out_h, out_w)] requested sizes -> the (N,4) fp32 device table
    [scale_x, scale_y, clip_w, clip_h] that :func:`detect` takes (postprocessing.py:27: scales are Python
    doubles, applied to fp32 boxes as fp32 factors)."""
    rows = []
    for (h, w), (oh, ow) in zip(image_sizes, output_sizes):
        rows.append([ow / w, oh / h, float(ow), float(oh)])
    return torch.tensor(rows, dtype=torch.float32).to(device)


def postprocess_boxes(boxes, image_size, output_height, output_width):
    """Boxes.scale + Boxes.clip + Boxes.nonempty of detector_postprocess for one image.
    -> (boxes (n,4) at the output resolution, keep (n) bool)."""
    b = _f32c(boxes).reshape(-1, 4)
    n = b.shape[0]
    out = torch.empty_like(b)
    keep = torch.empty(n, dtype=torch.uint8, device=b.device)
    if n:
        sx, sy = output_width / image_size[1], output_height / image_size[0]
        check(lib().fsg_postprocess_boxes(ptr(b), n, sx, sy, float(output_width), float(output_height), ptr(out),
                                          ptr(keep), stream()))
        count_launches(1)
    return out, keep.to(torch.bool)


def grid_anchors(grid_sizes, strides, cell_anchors, device):
    """DefaultAnchorGenerator.grid_anchors (anchor_generator.py:121-129) in one launch.
    grid_sizes [(H, W)], strides [int], cell_anchors [tensor (A_l, 4) fp32 on the host] per level.
    -> (anchors (R,4) on ``device``, level_offsets [L+1])."""
    nl = len(grid_sizes)
    levels = (_lib.AnchorLevel * nl)()
    offs = [0]
    for i, ((H, W), st, cell) in enumerate(zip(grid_sizes, strides, cell_anchors)):
        cell = torch.as_tensor(cell, dtype=torch.float32).cpu()
        A = cell.shape[0]
        if A > 16:
            raise RuntimeError("fsg_grid_anchors holds at most 16 cell anchors per level; got %d" % A)
        levels[i].H, levels[i].W, levels[i].stride, levels[i].A = int(H), int(W), int(st), A
        for a in range(A):
            for j in range(4):
                levels[i].cell[a][j] = float(cell[a, j])
        offs.append(offs[-1] + int(H) * int(W) * A)
    out = torch.empty((offs[-1], 4), dtype=torch.float32, device=device)
    check(lib().fsg_grid_anchors(levels, nl, ptr(out) if offs[-1] else None, offs[-1], stream()))
    count_launches(1)
    return out, offs


# ------------------------------------------------------------------------------------------------
# RPN / ROI-head callers of the same kernels
# ------------------------------------------------------------------------------------------------
def rpn_proposals(level_proposals, level_logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                  min_box_side_len=0.0):
    """Batched find_top_rpn_proposals (rpn_outputs.py:52-151): level_proposals list[(N, S_l, 4)],
    level_logits list[(N, S_l)], image_sizes [(h, w)] -> dict(boxes (N,post,4), logits (N,post),
    levels (N,post) int64, count (N) int32).  Two launches for the whole batch, no host sync."""
    nl = len(level_proposals)
    ps = [_f32c(p) for p in level_proposals]
    ls = [_f32c(x) for x in level_logits]
    N = ls[0].shape[0]
    dev = ls[0].device
    sizes = host_i64([x.shape[1] for x in ls])
    for p, x in zip(ps, ls):
        assert p.shape == (N, x.shape[1], 4)
    pp = (_lib.c_ptr * nl)(*[ptr(p) if p.numel() else None for p in ps])
    lp = (_lib.c_ptr * nl)(*[ptr(x) if x.numel() else None for x in ls])
    isz = torch.tensor([[float(h), float(w)] for (h, w) in image_sizes], dtype=torch.float32).to(dev)
    assert isz.shape == (N, 2)
    out = {
        "boxes": torch.empty((N, post_nms_topk, 4), dtype=torch.float32, device=dev),
        "logits": torch.empty((N, post_nms_topk), dtype=torch.float32, device=dev),
        "levels": torch.empty((N, post_nms_topk), dtype=torch.int64, device=dev),
        "count": torch.empty((N,), dtype=torch.int32, device=dev),
    }
    L = lib()
    need = L.fsg_rpn_proposals_workspace_bytes(N, sizes, nl, int(pre_nms_topk), int(post_nms_topk))
    if need == 0:
        raise RuntimeError("fsg_rpn_proposals: shape outside what the kernels cover (pre_nms_topk <= 16384 per "
                           "level, num_levels * pre_nms_topk <= 262144)")
    ws = _ws(need, dev)
    check(L.fsg_rpn_proposals(pp, lp, sizes, nl, N, ptr(isz), int(pre_nms_topk), int(post_nms_topk),
                              float(nms_thresh), float(min_box_side_len), ptr(out["boxes"]), ptr(out["logits"]),
                              ptr(out["levels"]), ptr(out["count"]), ptr(ws), ws.numel(), stream()))
    count_launches(2)   # (+ 3 per image and a gather on the general-n path)
    return out


def score_filter(boxes, scores, image_shape, score_thresh):
    """Candidate stage of fast_rcnn_inference_single_image (fast_rcnn.py:76-105).  boxes (R, C*4), scores
    (R, K+1) -> (cand_boxes (n,4), cand_scores (n), cand_classes (n) int64, cand_rows (n) int64); the count is read
    back once (the reference's ``nonzero`` synchronises at the same point)."""
    sc = _f32c(scores)
    R, K = sc.shape[0], sc.shape[1] - 1
    b = _f32c(boxes).reshape(R, -1)
    C = b.shape[1] // 4
    dev = sc.device
    cap = max(R * K, 1)
    ob = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    os_ = torch.empty((cap,), dtype=torch.float32, device=dev)
    oc = torch.empty((cap,), dtype=torch.int64, device=dev)
    orow = torch.empty((cap,), dtype=torch.int64, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    check(lib().fsg_score_filter(ptr(b) if R else None, C, ptr(sc) if R else None, R, K, float(image_shape[0]),
                                 float(image_shape[1]), float(score_thresh), ptr(ob), ptr(os_), ptr(oc), ptr(orow),
                                 ptr(cnt), stream()))
    count_launches(1)
    n = int(cnt.item())
    return ob[:n], os_[:n], oc[:n], orow[:n]
