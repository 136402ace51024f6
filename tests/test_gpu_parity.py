"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs."""
import math

import pytest
import torch

from oracle import dense_oracle as orc
from tests.util import assert_close_scalar, assert_close_tensor, assert_equal_int

pytestmark = pytest.mark.gpu


def _fsg():
    import full_scale_gambler_for_object_detection_b200 as fsg
    return fsg


# ------------------------------------------------------------------------------------------------ K1
def test_pairwise_iou_known_answer(cuda):
    """The reference's own KAT, tests/test_boxes.py:36-59."""
    fsg = _fsg()
    b1 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    b2 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 1.0, 0.5],
                       [0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.5, 1.5]])
    expected = torch.tensor([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2)
    got = fsg.pairwise_iou(fsg.Boxes(b1.to(cuda)), fsg.Boxes(b2.to(cuda)))
    assert torch.allclose(got.cpu(), expected)


@pytest.mark.parametrize("n1,n2", [(1, 1), (7, 300), (33, 4097), (0, 5), (5, 0)])
def test_pairwise_iou_bit_exact(cuda, n1, n2):
    fsg = _fsg()
    g = torch.Generator().manual_seed(n1 * 1000 + n2)
    def boxes(n):
        xy = torch.rand((n, 2), generator=g) * 100
        wh = torch.rand((n, 2), generator=g) * 60
        return torch.cat((xy, xy + wh), 1)
    b1, b2 = boxes(n1), boxes(n2)
    if n1 > 2:
        b1[1] = b1[0]          # duplicates -> exact ties
        b1[2, 2:] = b1[2, :2]  # zero-area box
    want = orc.pairwise_iou(b1, b2)
    got = fsg.pairwise_iou(fsg.Boxes(b1.to(cuda)), fsg.Boxes(b2.to(cuda)))
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want), "IoU must be bit-exact"


@pytest.mark.parametrize("M,N,thr,lq", [(8, 1000, [0.4, 0.5], True), (3, 257, [0.3, 0.7], False),
                                        (200, 5000, [0.4, 0.9], True), (0, 77, [0.4, 0.5], True),
                                        (5, 64, [0.5], True)])
def test_matcher_matrix(cuda, M, N, thr, lq):
    fsg = _fsg()
    g = torch.Generator().manual_seed(M * 7 + N)
    q = torch.rand((M, N), generator=g)
    q[q < 0.5] = 0.0                      # many zeros and zero columns
    if M > 1:
        q[1] = q[0]                       # tied rows -> argmax tie rule
    labels = [0, -1, 1] if len(thr) == 2 else [0, 1]
    want_m, want_l = orc.matcher(q, thr, labels, lq)
    m = fsg.Matcher(list(thr), labels, allow_low_quality_matches=lq)
    got_m, got_l = m(q.to(cuda))
    assert_equal_int(got_m, want_m, "matches")
    assert_equal_int(got_l, want_l, "match_labels")


def test_matcher_zero_overlap_gt_quirk(cuda):
    """A GT that overlaps nothing makes every zero-IoU anchor positive."""
    fsg = _fsg()
    q = torch.tensor([[0.9, 0.3, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    want = orc.matcher(q, [0.4, 0.5], [0, -1, 1], True)
    got = fsg.Matcher([0.4, 0.5], [0, -1, 1], True)(q.to(cuda))
    assert_equal_int(got[0], want[0], "matches")
    assert_equal_int(got[1], want[1], "labels")
    assert got[1].cpu().tolist() == [1, 1, 1, 1]


def _train_inputs(cfg_id, N, H, W, K, M=8):
    from full_scale_gambler_for_object_detection_b200 import synthetic
    return synthetic.train_inputs(cfg_id, N, H, W, K, M=M)


@pytest.mark.parametrize("N,H,W,M", [(2, 512, 512, 8), (3, 320, 480, 40), (1, 256, 256, 1)])
def test_fused_match_vs_oracle(cuda, N, H, W, M):
    fsg = _fsg()
    inp = _train_inputs(1, N, H, W, 80, M)
    want = orc.ground_truth(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], 80)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    got = fsg.ops.match_anchors(inp["anchors"].to(cuda), gt, 80,
                                want=("matches", "match_labels", "picky_labels", "gt_classes", "mask", "gt_deltas"))
    for k in ("matches", "match_labels", "picky_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)
    assert_close_tensor(got["gt_deltas"], want["gt_deltas"], "gt_deltas")
    if N > 1:  # N == 1: the synthetic batch's only image is the GT-free one (all-background path)
        assert int((want["match_labels"] == 1).sum()) > 0


@pytest.mark.parametrize("case", ["nested", "untouched", "duplicates"])
def test_fused_match_few_gt_low_quality_corner_cases(cuda, case):
    """Few-GT path (lane-held GT, ballot cull, argmax shortcut of pass B) on the cases that shortcut must get right:
    nested ground truth (an anchor whose argmax is the outer box is the best anchor of the inner one: promoted through a
    GT that is NOT its argmax), a GT that touches no anchor (maximum 0: every anchor of that image is promoted),
    duplicated GT and duplicated anchors (ties: lowest GT index wins, every tied anchor is promoted), IoU exactly 1.
    Labels / classes / mask bit-exact and the loss pre-pass sums against the oracle."""
    fsg = _fsg()
    inp = _train_inputs(4, 3, 384, 512, 80, 6)
    anchors = inp["anchors"].clone()
    gt_boxes = [b.clone() for b in inp["gt_boxes"]]
    gt_classes = [c.clone() for c in inp["gt_classes"]]
    if case == "nested":
        for b in gt_boxes:
            if b.shape[0] >= 4:
                b[0] = torch.tensor([40.0, 40.0, 360.0, 300.0])
                b[1] = torch.tensor([150.0, 120.0, 200.0, 170.0])     # inside GT 0
                b[2] = torch.tensor([140.0, 110.0, 230.0, 200.0])     # around GT 1, inside GT 0
                b[3] = torch.tensor([41.0, 41.0, 359.0, 299.0])       # almost GT 0
    elif case == "untouched":
        for n, b in enumerate(gt_boxes):
            if b.shape[0] >= 2 and n == 0:
                b[1] = torch.tensor([9000.0, 9000.0, 9100.0, 9050.0])
    else:
        for b in gt_boxes:
            if b.shape[0] >= 4:
                b[3] = b[1]                                           # duplicated GT
        anchors[101] = anchors[100]                                   # duplicated anchors
        anchors[102] = anchors[100]
        anchors[200] = gt_boxes[0][2] if gt_boxes[0].shape[0] > 2 else anchors[200]   # IoU exactly 1
    want = orc.ground_truth(anchors, gt_boxes, gt_classes, 80)
    gt = fsg.ops.PackedGT.from_lists(gt_boxes, gt_classes, cuda)
    R = anchors.shape[0]
    bets = torch.sigmoid(torch.randn((len(gt_boxes), R), generator=torch.Generator().manual_seed(3)) - 2.0)
    got = fsg.ops.match_anchors(anchors.to(cuda), gt, 80, bets=bets.to(cuda), temperature=0.1,
                                want=("matches", "match_labels", "picky_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "picky_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)
    if case == "untouched":
        assert bool((want["match_labels"][0] == 1).all())
    stats = got["stats"].cpu()
    fgm = (want["gt_classes"] >= 0) & (want["gt_classes"] != 80)
    assert int(stats[0]) == int(fgm.sum())
    S = (bets.double() * want["mask"].double() + 0.1).sum(dim=1)
    assert_close_tensor(stats[2:2 + len(gt_boxes)].float(), S.float(), "S[n]")
    assert abs(float(stats[1]) - float(S.sum())) <= 1e-6 * float(S.sum())


def test_fused_match_stress_slice(cuda):
    """Config 5 shape at a size the oracle finishes in seconds: 200 GT x 50k free-form anchors, 2 images
    with per-image anchors, > kGtChunk not needed; exercises ties through duplicated anchors."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    inp = synthetic.matcher_stress_inputs(5, 2, 50000, 200)
    inp["anchors"][0, 100] = inp["anchors"][0, 99]
    inp["anchors"][1, :50] = torch.cat((inp["gt_boxes"][1][:50, :2], inp["gt_boxes"][1][:50, 2:]), 1)  # IoU == 1
    anchors_list = [inp["anchors"][i] for i in range(2)]
    want = orc.ground_truth(anchors_list, inp["gt_boxes"], inp["gt_classes"], 80)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    got = fsg.ops.match_anchors(inp["anchors"].to(cuda), gt, 80,
                                want=("matches", "match_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)


@pytest.mark.parametrize("Ms", [(65, 256), (130, 200), (256, 65), (257, 64), (100, 0, 180), (200, 8, 20),
                                (90, 3, 2, 1, 0)])
def test_fused_match_crowded_images_spatial_prefilter(cuda, Ms):
    """Crowded images (64..257 GT per image, mixed with a GT-free one) on the shared-memory GT path; (200, 8, 20):
    few-GT images inside a crowded batch (the crowded pass A kernel with pass B's few-GT path); (90, 3, 2, 1, 0): a
    crowded image inside a few-GT batch (the few-GT pass A kernel with pass B's crowded path).  Includes: anchors far outside the GT extent,
    anchors covering the whole extent, GT that touch nothing (zero maximum: every anchor becomes a low-quality
    match), duplicated anchors and GT (ties -> lowest GT index), degenerate and identical GT boxes."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, R = len(Ms), 30000
    g = torch.Generator().manual_seed(500 + sum(Ms))
    base = synthetic.matcher_stress_inputs(7, N, R, 300)
    anchors = base["anchors"]
    gt_boxes, gt_classes = [], []
    for n, M in enumerate(Ms):
        b = base["gt_boxes"][n][:M].clone()
        if M > 10:
            b[3] = b[2]                                           # duplicated GT: the lower index must win
            b[5] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0]) if n == 0 else b[5]   # touches no anchor
            b[7, 2:] = b[7, :2]                                   # zero-area GT
            anchors[n, 11] = b[4]                                 # IoU exactly 1 with GT 4
            anchors[n, 12] = anchors[n, 11]
            anchors[n, 13] = torch.tensor([-5000.0, -5000.0, -4000.0, -4900.0])          # far outside the extent
            anchors[n, 14] = torch.tensor([-100.0, -100.0, 9000.0, 9000.0])              # covers everything
            anchors[n, 15] = torch.tensor([b[:, 0].min(), b[:, 1].min(), b[:, 0].min() + 1, b[:, 1].min() + 1])
        gt_boxes.append(b)
        gt_classes.append(torch.randint(0, 80, (M,), generator=g))
    want = orc.ground_truth([anchors[n] for n in range(N)], gt_boxes, gt_classes, 80)
    gt = fsg.ops.PackedGT.from_lists(gt_boxes, gt_classes, cuda)
    got = fsg.ops.match_anchors(anchors.to(cuda), gt, 80,
                                want=("matches", "match_labels", "picky_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "picky_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)
    if Ms[0] > 10:
        assert bool((want["match_labels"][0] == 1).all())        # the untouched GT of image 0 flips every anchor


def test_fused_match_crowded_identical_gt(cuda):
    """100 identical GT boxes (every anchor ties on all of them: index 0 must win) and a row of GT sharing one y
    interval."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    base = synthetic.matcher_stress_inputs(8, 2, 20000, 100)
    box = torch.tensor([[300.0, 300.0, 420.0, 380.0]])
    gt_boxes = [box.repeat(100, 1), base["gt_boxes"][1].clone()]
    gt_boxes[1][:, 1] = 200.0
    gt_boxes[1][:, 3] = 260.0
    gt_classes = base["gt_classes"]
    want = orc.ground_truth([base["anchors"][n] for n in range(2)], gt_boxes, gt_classes, 80)
    gt = fsg.ops.PackedGT.from_lists(gt_boxes, gt_classes, cuda)
    got = fsg.ops.match_anchors(base["anchors"].to(cuda), gt, 80, want=("matches", "match_labels", "mask"))
    for k in ("matches", "match_labels", "mask"):
        assert_equal_int(got[k], want[k], k)


@pytest.mark.parametrize("Ms", [(60, 8, 8, 0), (8, 33, 2, 80), (1030,) + (5,) * 40])
def test_fused_match_crowded_image_in_a_few_gt_batch_on_the_anchor_grid(cuda, Ms):
    """Detection training with one crowded image (COCO has images with 60..90 objects): the batch averages at most 32
    GT per image, so the few-GT pass A runs, and the crowded image takes its staged-GT branch with the warp-level
    cull on the regular anchor grid (1030 GT: two shared-memory chunks)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N = len(Ms)
    assert sum(Ms) <= 32 * N
    base = synthetic.train_inputs(71, N, 320, 448, 80, M=max(Ms), empty_image=False, logits=False)
    gt_boxes = [base["gt_boxes"][n][:m].clone() for n, m in enumerate(Ms)]
    gt_classes = [base["gt_classes"][n][:m].clone() for n, m in enumerate(Ms)]
    if Ms[0] >= 60:
        gt_boxes[0][7] = gt_boxes[0][3]                                          # duplicate: the lower index wins
        gt_boxes[0][9] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0])          # reaches no anchor: promotes all
    want = orc.ground_truth(base["anchors"], gt_boxes, gt_classes, 80)
    gt = fsg.ops.PackedGT.from_lists(gt_boxes, gt_classes, cuda)
    got = fsg.ops.match_anchors(base["anchors"].to(cuda), gt, 80,
                                want=("matches", "match_labels", "picky_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "picky_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)


def test_fused_match_crowded_coordinates_beyond_half_range(cuda):
    """The crowded pass A screens pairs in half precision with outward rounding; coordinates beyond the half range
    (65504) saturate towards "may overlap" and the exact drain decides.  Image 1 is image 0 scaled by 100 (up to
    1.5e5) and shifted, image 2 has tiny boxes (subnormal halves)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    base = synthetic.matcher_stress_inputs(9, 3, 20000, 100)
    anchors = base["anchors"].clone()
    gt_boxes = [b.clone() for b in base["gt_boxes"]]
    anchors[1] = anchors[0] * 100.0 - 40000.0
    gt_boxes[1] = gt_boxes[0] * 100.0 - 40000.0
    anchors[2] = anchors[0] * 1e-7
    gt_boxes[2] = gt_boxes[0] * 1e-7
    want = orc.ground_truth([anchors[n] for n in range(3)], gt_boxes, base["gt_classes"], 80)
    gt = fsg.ops.PackedGT.from_lists(gt_boxes, base["gt_classes"], cuda)
    got = fsg.ops.match_anchors(anchors.to(cuda), gt, 80, want=("matches", "match_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)
    assert int((want["match_labels"][1] == 1).sum()) > 0 and int((want["match_labels"][2] == 1).sum()) > 0


def test_fused_match_many_gt_chunks(cuda):
    """More GT than one shared-memory chunk (1024)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    inp = synthetic.matcher_stress_inputs(6, 1, 3000, 2500)
    want = orc.ground_truth([inp["anchors"][0]], inp["gt_boxes"], inp["gt_classes"], 80)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    got = fsg.ops.match_anchors(inp["anchors"].to(cuda), gt, 80, want=("matches", "match_labels", "mask"))
    for k in ("matches", "match_labels", "mask"):
        assert_equal_int(got[k], want[k], k)


def test_box2box_roundtrip_and_parity(cuda):
    """tests/test_box2box_transform.py:16-30 (weights (5,5,10,10)) + parity with the oracle."""
    fsg = _fsg()
    torch.manual_seed(3)
    w = (5, 5, 10, 10)
    src = torch.rand(10, 4) * 1 + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    dst = torch.rand(10, 4) * 1 + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    t = fsg.Box2BoxTransform(weights=w)
    d = t.get_deltas(src.to(cuda), dst.to(cuda))
    rec = t.apply_deltas(d, src.to(cuda))
    assert torch.allclose(dst, rec.cpu())
    assert_close_tensor(d, orc.get_deltas(src, dst, w), "get_deltas")
    big = torch.randn(64, 12) * 3
    boxes = torch.rand(64, 4) * 50
    boxes[:, 2:] += boxes[:, :2] + 1
    assert_close_tensor(t.apply_deltas(big.to(cuda), boxes.to(cuda)), orc.apply_deltas(big, boxes, w), "apply_deltas")


# ------------------------------------------------------------------------------------------------ K2
def _check_step(cuda, inp, K, coeffs, cfg_kwargs=None, detach_pred=False, bets_floor=1e-6):
    fsg = _fsg()
    cfg_kwargs = cfg_kwargs or {}
    cfg = fsg.DenseLossConfig(num_classes=K, **cfg_kwargs)
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                          inp["bets"], K, *coeffs, temperature=cfg.gambler_temperature, normalize=cfg.normalize,
                          mode=cfg.gambler_loss_mode, alpha=cfg.focal_alpha, focal_gamma=cfg.focal_gamma,
                          gambler_gamma=cfg.gambler_gamma, beta=cfg.smooth_l1_beta, output=cfg.gambler_output,
                          detach_pred=detach_pred)
    x = inp["logits"].to(cuda).requires_grad_(True)
    d = inp["deltas"].to(cuda).requires_grad_(True)
    b = inp["bets"].to(cuda).requires_grad_(True)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    res = fsg.dense_train_step(x, d, b, inp["anchors"].to(cuda), gt, cfg, coeffs, detach_pred=detach_pred,
                               want_weights=True)
    res.total.backward()
    assert_equal_int(res.gt_classes, want["gt_classes"], "gt_classes")
    assert_equal_int(res.mask, want["mask"], "mask")
    assert int(res.num_foreground.item()) == int(want["num_foreground"])
    assert_close_scalar(res.loss_cls.item(), want["loss_cls"], "loss_cls")
    assert_close_scalar(res.loss_box_reg.item(), want["loss_box_reg"], "loss_box_reg")
    assert_close_scalar(res.gambler_loss.item(), want["gambler_loss"], "gambler_loss")
    assert_close_scalar(res.total.item(), want["total"], "total", rtol=2e-5)
    assert_close_scalar(res.loss_before_weighting(cfg.gambler_loss_mode).item(), want["loss_before_weighting"],
                        "loss_before_weighting")
    assert_close_scalar(res.lower_bound(cfg.gambler_temperature, 1.0).item(), want["lower_bound"], "lower_bound")
    assert_close_tensor(res.per_anchor_loss, want["per_anchor_loss"], "per_anchor_loss")
    assert_close_tensor(res.weights, want["weights"], "weights")
    if not detach_pred:
        assert_close_tensor(x.grad, want["grad_logits"], "grad_logits")
    else:
        assert x.grad is None
    if coeffs[1] != 0:
        assert_close_tensor(d.grad, want["grad_deltas"], "grad_deltas")
    # d/d bets = -(m/S)(l - A) cancels when l ~ A: absolute floor = fp32 rounding of l and A themselves
    assert_close_tensor(b.grad, want["grad_bets"], "grad_bets", atol_scale=bets_floor)
    return res


def test_step_config1(cuda):
    """BASELINE config 1: 2 x 512x512, K=80, one image without GT."""
    inp = _train_inputs(1, 2, 512, 512, 80)
    assert inp["R"] == 16368
    _check_step(cuda, inp, 80, (1.0, 1.0, -1.0))


def test_step_gambler_phase(cuda):
    inp = _train_inputs(2, 2, 256, 320, 80)
    _check_step(cuda, inp, 80, (0.0, 0.0, 1.0), detach_pred=True)


@pytest.mark.parametrize("kw", [dict(gambler_output="L_BAHW_extendtobatch"), dict(normalize=False),
                                dict(gambler_loss_mode="sigmoid"), dict(focal_gamma=1.5, focal_alpha=-1.0),
                                dict(gambler_gamma=2.0), dict(smooth_l1_beta=0.0, gambler_temperature=0.03)])
def test_step_variants(cuda, kw):
    inp = _train_inputs(3, 3, 256, 256, 80, M=5)
    # "sigmoid" mode: torch computes BCE-with-logits as (1-t)*x - log_sigmoid(x), which cancels for the
    # (typical) very negative logits, so the reference's own fp32 per-anchor loss is only good to ~5e-6 and its
    # d/d bets to 4e-6 of the tensor's scale (measured against an fp64 evaluation of the same graph).
    floor = 1e-5 if kw.get("gambler_loss_mode") == "sigmoid" else 1e-6
    _check_step(cuda, inp, 80, (1.0, 0.5, -2.0), kw, bets_floor=floor)


@pytest.mark.parametrize("K", [1230, 3, 20, 1])
def test_step_other_class_counts(cuda, K):
    """LVIS K=1230 (rows not 16-byte aligned -> float2 path), odd K (scalar path), small K."""
    inp = _train_inputs(4, 2, 192, 192, K, M=6)
    _check_step(cuda, inp, K, (1.0, 1.0, -1.0))


def test_step_run_to_run_deterministic(cuda):
    fsg = _fsg()
    inp = _train_inputs(1, 2, 256, 256, 80)
    cfg = fsg.DenseLossConfig()
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    outs = []
    for _ in range(3):
        x = inp["logits"].to(cuda).requires_grad_(True)
        r = fsg.dense_train_step(x, inp["deltas"].to(cuda), inp["bets"].to(cuda), inp["anchors"].to(cuda), gt, cfg)
        r.total.backward()
        outs.append((r.scalars.clone(), x.grad.clone()))
    for s, g in outs[1:]:
        assert torch.equal(s, outs[0][0]) and torch.equal(g, outs[0][1])


def test_one_call_step_equals_staged_step(cuda):
    """fsg_dense_step (persistent K1 + K2 under programmatic dependent launch, what DenseStepPlan runs) against the
    staged entry points on the same inputs: integer outputs bit-identical, floats equal up to the order in which the
    bet normaliser S[n] is summed (per-CTA partials are grouped differently); also through a CUDA graph, twice."""
    fsg = _fsg()
    inp = _train_inputs(61, 5, 320, 448, 80, M=7)
    N, R, K = inp["N"], inp["R"], 80
    x, d, b = (inp[k].to(cuda) for k in ("logits", "deltas", "bets"))
    anchors = inp["anchors"].to(cuda)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    for kw in (dict(), dict(gambler_output="L_BAHW_extendtobatch"), dict(normalize=False)):
        cfg = fsg.DenseLossConfig(num_classes=K, **kw)
        a = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), want_weights=True)
        assert a.one_call
        s = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), want_weights=True)
        s.one_call = False
        ra, rs = a.run(x, d, b, anchors, gt), s.run(x, d, b, anchors, gt)
        assert torch.equal(ra.gt_classes, rs.gt_classes) and torch.equal(ra.mask, rs.mask)
        assert torch.equal(a.matched, s.matched) and float(ra.stats[0]) == float(rs.stats[0])
        assert_close_tensor(ra.stats, rs.stats, "stats", rtol=1e-6)
        assert_close_tensor(ra.scalars, rs.scalars, "scalars", rtol=1e-6)
        assert_close_tensor(a.grad_logits, s.grad_logits, "grad_logits", rtol=2e-6)
        assert torch.equal(a.grad_deltas, s.grad_deltas)
        assert_close_tensor(a.grad_bets, s.grad_bets, "grad_bets", rtol=1e-5, atol_scale=1e-6)
        assert_close_tensor(a.weights, s.weights, "weights", rtol=2e-6)
        keep = [t.clone() for t in (a.grad_logits, a.grad_deltas, a.grad_bets, a.scalars, a.gt_classes)]
        a.capture(x, d, b, anchors, gt)
        for _ in range(2):
            a.replay()
            torch.cuda.synchronize()
            for t, u in zip(keep, (a.grad_logits, a.grad_deltas, a.grad_bets, a.scalars, a.gt_classes)):
                assert torch.equal(t, u)      # the graph replays the very same launches: bit-identical
        a.release_graphs()


def test_step_workspace_stays_clean_across_changing_ground_truth(cuda):
    """The plan's one-call step enqueues no memset (fsg_match_config.workspace_is_clean): K1's fold kernel zeroes what
    the call dirtied.  Steps with a growing and shrinking number of GT (the per-GT maxima move inside the workspace), a
    GT that touches no anchor (every anchor promoted) and a GT-free batch must each equal the same step on a plan
    whose call clears its workspace itself -- bit for bit."""
    fsg = _fsg()
    N, K = 3, 80
    base = _train_inputs(63, N, 256, 320, K, M=12)
    R = base["R"]
    x, d, b = (base[k].to(cuda) for k in ("logits", "deltas", "bets"))
    anchors = base["anchors"].to(cuda)
    cfg = fsg.DenseLossConfig(num_classes=K)
    clean = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), max_total_gt=256)
    assert clean.one_call and clean._mc.workspace_is_clean == 1
    boxes, classes = base["gt_boxes"], base["gt_classes"]
    far = torch.tensor([[9000.0, 9000.0, 9100.0, 9050.0]])
    rounds = [
        ([bx[:2] for bx in boxes], [c[:2] for c in classes]),
        (boxes, classes),
        ([torch.cat((bx[:1], far)) if bx.shape[0] else bx for bx in boxes],
         [torch.cat((c[:1], c[:1])) if c.shape[0] else c for c in classes]),
        ([bx[:0] for bx in boxes], [c[:0] for c in classes]),
        ([bx[:5] for bx in boxes], [c[:5] for c in classes]),
        (boxes, classes),
    ]
    for i, (gb, gc) in enumerate(rounds):
        gt = fsg.ops.PackedGT.from_lists(gb, gc, cuda)
        fresh = fsg.DenseStepPlan(N, R, K, cfg, cuda, (1.0, 0.5, -2.0), max_total_gt=256)
        fresh._mc.workspace_is_clean = 0
        fresh.ws_step.fill_(0xAB)          # the self-clearing call must not depend on the buffer's content
        rc, rf = clean.run(x, d, b, anchors, gt), fresh.run(x, d, b, anchors, gt)
        for name in ("gt_classes", "mask", "stats", "scalars"):
            assert torch.equal(getattr(rc, name), getattr(rf, name)), "round %d: %s" % (i, name)
        assert torch.equal(clean.matched, fresh.matched)
        assert torch.equal(clean.grad_logits, fresh.grad_logits) and torch.equal(clean.grad_bets, fresh.grad_bets)
        assert torch.equal(clean.grad_deltas, fresh.grad_deltas)


def test_backward_is_single_use(cuda):
    fsg = _fsg()
    inp = _train_inputs(62, 2, 128, 128, 80, M=3)
    x = inp["logits"].to(cuda).requires_grad_(True)
    d = inp["deltas"].to(cuda).requires_grad_(True)
    b = inp["bets"].to(cuda).requires_grad_(True)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    res = fsg.dense_train_step(x, d, b, inp["anchors"].to(cuda), gt, fsg.DenseLossConfig(num_classes=80))
    res.total.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="twice"):
        res.total.backward()


def test_upstream_gradient_scaling(cuda):
    fsg = _fsg()
    inp = _train_inputs(1, 1, 128, 128, 80, M=3)
    cfg = fsg.DenseLossConfig()
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    grads = []
    for scale in (1.0, 3.0):
        x = inp["logits"].to(cuda).requires_grad_(True)
        r = fsg.dense_train_step(x, inp["deltas"].to(cuda), inp["bets"].to(cuda), inp["anchors"].to(cuda), gt,
                                 cfg, (1.0, 1.0, -1.0))
        (r.total * scale).backward()
        grads.append(x.grad.clone())
    assert torch.allclose(grads[1], grads[0] * 3.0, rtol=1e-6, atol=0)


# ---------------------------------------------------------------------------------- drop-in methods
def _levels(inp, K, N, gen, scale=1.0, shift=0.0):
    A = inp["A"]
    return [torch.randn((N, A * K, h, w), generator=gen) * scale + shift for (h, w) in inp["grids"]]


@pytest.mark.parametrize("kw,K,coeffs", [
    (dict(), 80, (1.0, 1.0, -1.0)),
    (dict(gambler_output="L_BAHW_extendtobatch"), 80, (1.0, 0.5, -2.0)),
    (dict(gambler_loss_mode="sigmoid"), 80, (1.0, 0.5, -2.0)),
    (dict(focal_gamma=1.5, focal_alpha=-1.0, gambler_gamma=2.0), 20, (1.0, 1.0, -1.0)),
    (dict(normalize=False, smooth_l1_beta=0.0), 7, (1.0, 1.0, -1.0)),
    (dict(), 1230, (1.0, 1.0, -1.0)),
])
def test_loss_main_on_native_head_layout(cuda, kw, K, coeffs):
    """fsg_loss_main_levels: the fused main pass reading (N, A*K, H, W) / (N, A*4, H, W) in place and writing the
    gradients in that layout, against the oracle's permute+cat data flow."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import _lib, synthetic
    N = 3
    inp = synthetic.train_inputs(21, N, 200, 264, K, M=5, logits=False)
    A, grids = inp["A"], inp["grids"]
    gen = torch.Generator().manual_seed(22)
    cls_l = _levels(inp, K, N, gen, 1.0, synthetic.PRIOR_LOGIT)
    reg_l = _levels(inp, 4, N, gen, 0.1)
    bets = torch.sigmoid(torch.randn((N, inp["R"]), generator=gen) - 4.0)
    cfg = fsg.DenseLossConfig(num_classes=K, **kw)
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], orc.levels_to_flat(cls_l, K),
                          orc.levels_to_flat(reg_l, 4), bets, K, *coeffs, temperature=cfg.gambler_temperature,
                          normalize=cfg.normalize, mode=cfg.gambler_loss_mode, alpha=cfg.focal_alpha,
                          focal_gamma=cfg.focal_gamma, gambler_gamma=cfg.gambler_gamma, beta=cfg.smooth_l1_beta,
                          output=cfg.gambler_output)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    anchors = inp["anchors"].to(cuda)
    b = bets.to(cuda)
    m = fsg.ops.match_anchors(anchors, gt, K, bets=b, temperature=cfg.gambler_temperature)
    params = cfg.loss_params(*coeffs)
    out = fsg.ops.loss_main_levels([t.to(cuda) for t in cls_l], m["gt_classes"], params, m["stats"],
                                   delta_levels=[t.to(cuda) for t in reg_l], anchors=anchors, gt=gt,
                                   matched_idx32=m["matched_idx32"], mask=m["mask"], bets=b, want_weights=True)
    gb = fsg.ops.loss_post(b, m["mask"], out["per_anchor_loss"], params, m["stats"], out["scalars"])
    sc = out["scalars"].cpu()
    assert_close_scalar(sc[5], want["loss_cls"], "loss_cls")
    assert_close_scalar(sc[6], want["loss_box_reg"], "loss_box_reg")
    assert_close_scalar(sc[7], want["gambler_loss"], "gambler_loss")
    assert_close_tensor(out["per_anchor_loss"], want["per_anchor_loss"], "per_anchor_loss")
    assert_close_tensor(out["weights"], want["weights"], "weights")
    assert_close_tensor(orc.levels_to_flat([g.cpu() for g in out["grad_logits"]], K), want["grad_logits"], "grad_logits")
    assert_close_tensor(orc.levels_to_flat([g.cpu() for g in out["grad_deltas"]], 4), want["grad_deltas"], "grad_deltas")
    floor = 1e-5 if kw.get("gambler_loss_mode") == "sigmoid" else 1e-6
    assert_close_tensor(gb, want["grad_bets"], "grad_bets", atol_scale=floor)
    # and against the (N, R, K) kernel on the permuted copy: same element arithmetic, so the per-element
    # gradients agree to the last bit; sums differ only by the order of the K-reduction
    flat = fsg.ops.loss_main(fsg.ops.levels_to_flat([t.to(cuda) for t in cls_l], K), m["gt_classes"], params,
                             m["stats"], pred_deltas=fsg.ops.levels_to_flat([t.to(cuda) for t in reg_l], 4),
                             anchors=anchors, gt=gt, matched_idx32=m["matched_idx32"], mask=m["mask"], bets=b)
    lv = fsg.ops.levels_to_flat(out["grad_logits"], K)
    assert torch.equal(lv, flat["grad_logits"])
    assert torch.equal(fsg.ops.levels_to_flat(out["grad_deltas"], 4), flat["grad_deltas"])


@pytest.mark.parametrize("native", [True, False])
def test_dropin_retinanet_losses_and_gt(cuda, native):
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, K = 2, 80
    inp = synthetic.train_inputs(7, N, 256, 384, K, logits=False)
    gen = torch.Generator().manual_seed(11)
    cls_l = _levels(inp, K, N, gen, 1.0, synthetic.PRIOR_LOGIT)
    reg_l = _levels(inp, 4, N, gen, 0.1)
    want_gt = orc.ground_truth(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], K)
    xs = [t.clone().requires_grad_(True) for t in cls_l]
    ds = [t.clone().requires_grad_(True) for t in reg_l]
    lc, lr, _ = orc.retinanet_losses(want_gt["gt_classes"], want_gt["gt_deltas"], orc.levels_to_flat(xs, K),
                                     orc.levels_to_flat(ds, 4), K)
    (lc + 2 * lr).backward()

    path = fsg.RetinaNetDensePath(num_classes=K, native_layout=native)
    offs = inp["level_offsets"]
    anc_levels = [fsg.Boxes(inp["anchors"][offs[i]:offs[i + 1]].to(cuda)) for i in range(5)]
    anchors = [anc_levels for _ in range(N)]
    targets = []
    for b, c in zip(inp["gt_boxes"], inp["gt_classes"]):
        t = fsg.Instances((256, 384))
        t.gt_boxes = fsg.Boxes(b.to(cuda))
        t.gt_classes = c.to(cuda)
        targets.append(t)
    gt_classes, gt_deltas = path.get_ground_truth(anchors, targets)
    mask = path.get_picky_ground_truth(anchors, targets)
    assert_equal_int(gt_classes, want_gt["gt_classes"], "gt_classes")
    assert_equal_int(mask, want_gt["mask"], "mask")
    assert_close_tensor(gt_deltas, want_gt["gt_deltas"], "gt_deltas")
    gx = [t.to(cuda).requires_grad_(True) for t in cls_l]
    gd = [t.to(cuda).requires_grad_(True) for t in reg_l]
    losses = path.losses(gt_classes, gt_deltas, gx, gd)
    (losses["loss_cls"] + 2 * losses["loss_box_reg"]).backward()
    assert_close_scalar(losses["loss_cls"].item(), lc, "loss_cls")
    assert_close_scalar(losses["loss_box_reg"].item(), lr, "loss_box_reg")
    for a, b in zip(gx, xs):
        assert_close_tensor(a.grad, b.grad, "grad cls level")
    for a, b in zip(gd, ds):
        assert_close_tensor(a.grad, b.grad, "grad reg level")


@pytest.mark.parametrize("detach,native", [(False, True), (True, True), (False, False), (True, False)])
def test_dropin_gambler_loss(cuda, detach, native):
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, K, A = 2, 80, 3
    inp = synthetic.train_inputs(8, N, 256, 256, K, logits=False)
    gen = torch.Generator().manual_seed(12)
    cls_l = _levels(inp, K, N, gen, 1.0, synthetic.PRIOR_LOGIT)
    bet_l = [torch.sigmoid(torch.randn((N, A, h, w), generator=gen) - 4.0) for (h, w) in inp["grids"]]
    gtd = orc.ground_truth(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], K)
    xs = [t.clone().requires_grad_(True) for t in cls_l]
    bs = [t.clone().requires_grad_(True) for t in bet_l]
    xf = orc.levels_to_flat(xs, K)
    want = orc.gambler_loss(xf.detach() if detach else xf, orc.levels_to_flat(bs, 1).reshape(N, -1),
                            gtd["gt_classes"], gtd["mask"], K)
    want["gambler_loss"].backward()

    head = fsg.GamblerLoss(num_classes=K, native_layout=native)
    gx = [t.to(cuda).requires_grad_(True) for t in cls_l]
    gb = [t.to(cuda).requires_grad_(True) for t in bet_l]
    bets_list = list(gb)
    loss_dict, w = head.gambler_loss(gx, bets_list, gtd["gt_classes"].to(cuda), gtd["mask"].to(cuda), detach)
    loss_dict["gambler_loss"].backward()
    assert_close_scalar(loss_dict["gambler_loss"].item(), want["gambler_loss"], "gambler_loss")
    assert_close_scalar(loss_dict["loss_before_weighting"].item(), want["loss_before_weighting"], "lbw")
    assert_close_scalar(head.last_lower_bound.item(), want["lower_bound"], "lower_bound")
    assert_close_tensor(w.reshape(N, -1), want["weights"], "weights")
    nakhw_want = orc.flat_to_nahw(want["per_anchor_loss"], inp["grids"], A)
    for a, b in zip(loss_dict["NAKHW_loss"], nakhw_want):
        assert_close_tensor(a, b, "NAKHW_loss")
    ub = fsg.get_loss_upper_bound(loss_dict["NAKHW_loss"], N, 0.1, 1.0)
    assert_close_scalar(-ub.item(), want["lower_bound"], "get_loss_upper_bound")
    for a, b in zip(gb, bs):
        assert_close_tensor(a.grad, b.grad, "grad bets level", atol_scale=1e-6)
    if detach:
        assert all(t.grad is None for t in gx)
    else:
        for a, b in zip(gx, xs):
            assert_close_tensor(a.grad, b.grad, "grad logits level")
    # the reference's in-place masking of the caller's list (gambler_heads.py:568-569)
    off = 0
    for i, (h, w_) in enumerate(inp["grids"]):
        n = h * w_ * A
        m = gtd["mask"][:, off:off + n].reshape(N, h, w_, A).permute(0, 3, 1, 2)
        assert torch.equal(bets_list[i].detach().cpu(), bet_l[i] * m)
        off += n


def test_layout_roundtrip(cuda):
    fsg = _fsg()
    gen = torch.Generator().manual_seed(5)
    levels = [torch.randn((2, 3 * 7, h, w), generator=gen) for (h, w) in [(9, 13), (5, 6), (2, 2)]]
    want = orc.levels_to_flat(levels, 7)
    got = fsg.ops.levels_to_flat([t.to(cuda) for t in levels], 7)
    assert torch.equal(got.cpu(), want)
    back = fsg.ops.flat_to_levels(got, [tuple(t.shape[1:]) for t in levels])
    for a, b in zip(back, levels):
        assert torch.equal(a.cpu(), b)


# ------------------------------------------------------------------------------------------------ K3
def _nms_inputs(n, ncls, seed, dup=True):
    g = torch.Generator().manual_seed(seed)
    boxes = torch.rand((n, 4), generator=g) * 100
    boxes[:, 2:] += boxes[:, :2]                       # tests/test_nms_rotated.py:35-43
    scores = torch.rand(n, generator=g)
    if dup and n > 10:
        scores[5] = scores[3]                          # tied scores -> stable order
        boxes[7] = boxes[6]                            # identical boxes -> IoU exactly 1
        boxes[9, 2:] = boxes[9, :2]                    # zero-area box (0/0 -> NaN never suppresses)
    idxs = torch.randint(0, ncls, (n,), generator=g)
    return boxes, scores, idxs


@pytest.mark.parametrize("n,thr", [(1, 0.5), (2, 0.5), (300, 0.2), (2000, 0.5), (5000, 0.8), (8192, 0.5), (0, 0.5)])
def test_nms_bit_exact(cuda, n, thr):
    fsg = _fsg()
    boxes, scores, _ = _nms_inputs(n, 1, 100 + n)
    want = orc.nms(boxes, scores, thr)
    got = fsg.nms(boxes.to(cuda), scores.to(cuda), thr)
    assert_equal_int(got, want, "nms keep")


@pytest.mark.parametrize("n,ncls,thr", [(2000, 50, 0.5), (2000, 50, 0.2), (2000, 50, 0.8), (5000, 80, 0.5),
                                        (4000, 1230, 0.5), (3000, 1, 0.5)])
def test_batched_nms_bit_exact(cuda, n, ncls, thr):
    """tests/test_nms_rotated.py:45-66 sizes (N=2000, 50 classes, IoU 0.2/0.5/0.8) and the RetinaNet sizes."""
    fsg = _fsg()
    boxes, scores, idxs = _nms_inputs(n, ncls, 200 + n + ncls)
    want = orc.batched_nms(boxes, scores, idxs, thr)
    got = fsg.batched_nms(boxes.to(cuda), scores.to(cuda), idxs.to(cuda), thr)
    assert_equal_int(got, want, "batched_nms keep")


def _sparse_nms_inputs(n, ncls, seed, extent):
    """Boxes spread over `extent` pixels so that a large share survives (long kept lists, every bit-matrix
    column block in use); a few exact ties / duplicates as in _nms_inputs."""
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand((n, 2), generator=g) * extent
    wh = torch.rand((n, 2), generator=g) * 90 + 10
    boxes = torch.cat([xy, xy + wh], dim=1)
    scores = torch.rand(n, generator=g)
    scores[n // 2] = scores[n // 3]
    scores[n - 1] = scores[0]
    boxes[n // 4] = boxes[n // 5]
    idxs = torch.randint(0, ncls, (n,), generator=g)
    return boxes, scores, idxs


@pytest.mark.parametrize("n,thr,extent", [(8193, 0.5, 100.0), (8256, 0.5, 3000.0), (20000, 0.7, 2000.0),
                                          (40000, 0.5, 800.0)])
def test_nms_large_n_bit_exact(cuda, n, thr, extent):
    """More boxes than the shared-memory kernel holds: rank / bit-matrix / sweep path (csrc/nms_large.cu)."""
    fsg = _fsg()
    boxes, scores, _ = _sparse_nms_inputs(n, 1, 300 + n, extent)
    want = orc.nms(boxes, scores, thr)
    got = fsg.nms(boxes.to(cuda), scores.to(cuda), thr)
    assert_equal_int(got, want, "nms keep (large n)")


@pytest.mark.parametrize("n,ncls,thr", [(12000, 5, 0.7), (30000, 80, 0.5)])
def test_batched_nms_large_n_bit_exact(cuda, n, ncls, thr):
    """RPN-sized call (5 levels x 2000+ proposals, rpn_outputs.py:137) and a large multi-class call."""
    fsg = _fsg()
    boxes, scores, idxs = _sparse_nms_inputs(n, ncls, 400 + n, 1200.0)
    want = orc.batched_nms(boxes, scores, idxs, thr)
    got = fsg.batched_nms(boxes.to(cuda), scores.to(cuda), idxs.to(cuda), thr)
    assert_equal_int(got, want, "batched_nms keep (large n)")


def test_batched_nms_rejects_class_ids_outside_the_key_range(cuda):
    fsg = _fsg()
    b, s_, c = _nms_inputs(100, 5, 3)
    c = c.clone()
    c[7] = 1 << 18
    with pytest.raises(ValueError, match="class ids"):
        fsg.batched_nms(b.to(cuda), s_.to(cuda), c.to(cuda), 0.5)
    c[7] = -1
    with pytest.raises(ValueError, match="class ids"):
        fsg.batched_nms(b.to(cuda), s_.to(cuda), c.to(cuda), 0.5)
    c[7] = (1 << 18) - 1
    assert fsg.batched_nms(b.to(cuda), s_.to(cuda), c.to(cuda), 0.5).numel() > 0


def test_nms_threshold_compare_is_in_double(cuda):
    """IoU float(1/3) against threshold 1/3 (double): float(1/3) > 1/3 -> suppressed."""
    fsg = _fsg()
    boxes = torch.tensor([[0.0, 0.0, 2.0, 1.0], [1.0, 0.0, 3.0, 1.0]])
    scores = torch.tensor([0.9, 0.8])
    want = orc.nms(boxes, scores, 1.0 / 3.0)
    got = fsg.nms(boxes.to(cuda), scores.to(cuda), 1.0 / 3.0)
    assert_equal_int(got, want, "keep")


def _detect_case(cuda, N, counts, K, seed, topk=1000):
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    inp = synthetic.inference_inputs(seed, N, counts, K)
    res = fsg.ops.detect(inp["logits"].to(cuda), inp["deltas"].to(cuda), inp["anchors"].to(cuda),
                         inp["level_offsets"], topk=topk, want_candidates=True)
    offs = inp["level_offsets"]
    for n in range(N):
        cls = [inp["logits"][n, offs[i]:offs[i + 1]] for i in range(len(counts))]
        reg = [inp["deltas"][n, offs[i]:offs[i + 1]] for i in range(len(counts))]
        anc = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(len(counts))]
        (wb, ws, wc), (cb, cs, cc), keep = orc.inference_single_image(cls, reg, anc, K, topk_candidates=topk)
        cnt = int(res["cand_count"][n].item())
        # candidate selection: same (anchor, class) set in the same order, scores/boxes within fp32 tolerance
        assert cnt == cb.shape[0], "candidate count %d vs %d" % (cnt, cb.shape[0])
        assert_equal_int(res["cand_classes"][n, :cnt], cc, "cand classes")
        assert_close_tensor(res["cand_scores"][n, :cnt], cs, "cand scores")
        assert_close_tensor(res["cand_boxes"][n, :cnt], cb, "cand boxes", atol_scale=1e-6)
        # NMS on the GPU's own candidates must be bit-exact with the oracle NMS on the same candidates
        gb, gs, gc = res["cand_boxes"][n, :cnt].cpu(), res["cand_scores"][n, :cnt].cpu(), res["cand_classes"][n, :cnt].cpu()
        keep2 = orc.batched_nms(gb, gs, gc, 0.5)[:100]
        d = int(res["count"][n].item())
        assert d == keep2.shape[0]
        assert_equal_int(res["keep_idx"][n, :d], keep2, "keep idx")
        assert torch.equal(res["boxes"][n, :d].cpu(), gb[keep2])
        assert torch.equal(res["scores"][n, :d].cpu(), gs[keep2])
        assert torch.equal(res["classes"][n, :d].cpu(), gc[keep2])
        assert float(res["scores"][n, d:].abs().sum()) == 0.0


def test_detect_small(cuda):
    _detect_case(cuda, 2, [3000, 800, 200, 60, 20], 80, 41)


def test_detect_multi_part_levels(cuda):
    """Levels large enough to be split over several CTAs (merge path) and to trigger in-loop pruning."""
    _detect_case(cuda, 1, [24000, 6000, 1500], 80, 42)


def test_detect_saturated_score_ties(cuda):
    """Thousands of logits so large that their fp32 sigmoid is exactly 1.0, and a block of equal mid-range logits
    straddling the top-k boundary: the selection among equal scores must go to the lowest indices (the reference's
    stable sort), including for candidates that wait in a warp queue while later ones are already in the buffer."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    K = 80
    inp = synthetic.inference_inputs(44, 2, [24000, 3000], K)
    g = torch.Generator().manual_seed(45)
    flat0 = inp["logits"][0].view(-1)
    sat = torch.randperm(400000, generator=g)[:3500]               # all inside the first CTA's stream of level 0
    flat0[sat] = 30.0                                             # sigmoid == 1.0f: 3500 exact ties, top-k is 1000
    flat1 = inp["logits"][1].view(-1)
    flat1[torch.randperm(flat1.numel(), generator=g)[:900]] = 6.0   # 900 clear winners ...
    flat1[torch.randperm(flat1.numel(), generator=g)[:5000]] = 2.5  # ... and 5000 tied at the boundary
    res = fsg.ops.detect(inp["logits"].to(cuda), inp["deltas"].to(cuda), inp["anchors"].to(cuda),
                         inp["level_offsets"], want_candidates=True)
    offs = inp["level_offsets"]
    for n in range(2):
        cls = [inp["logits"][n, offs[i]:offs[i + 1]] for i in range(2)]
        reg = [inp["deltas"][n, offs[i]:offs[i + 1]] for i in range(2)]
        anc = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(2)]
        _, (cb, cs, cc), _ = orc.inference_single_image(cls, reg, anc, K)
        cnt = int(res["cand_count"][n].item())
        assert cnt == cb.shape[0]
        assert_equal_int(res["cand_classes"][n, :cnt], cc, "cand classes")
        assert_close_tensor(res["cand_boxes"][n, :cnt], cb, "cand boxes", atol_scale=1e-6)
        assert_close_tensor(res["cand_scores"][n, :cnt], cs, "cand scores")


@pytest.mark.parametrize("seed", [46, 47, 48])
def test_detect_tie_boundary_inside_one_stream(cuda, seed):
    """A single-CTA slab (1500 anchors x 80) with 300 clear winners and 3000 logits tied exactly at the top-k
    boundary: 700 of the tied ones must be kept, the ones with the lowest indices -- also when a lower-index
    candidate is still waiting in a warp's queue while higher-index ones already sit in the key buffer."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    K = 80
    inp = synthetic.inference_inputs(seed, 1, [1500], K)
    g = torch.Generator().manual_seed(seed)
    flat = inp["logits"][0].view(-1)
    perm = torch.randperm(flat.numel(), generator=g)
    flat[perm[:3000]] = 2.5
    flat[perm[3000:3300]] = 6.0
    res = fsg.ops.detect(inp["logits"].to(cuda), inp["deltas"].to(cuda), inp["anchors"].to(cuda),
                         inp["level_offsets"], want_candidates=True)
    _, (cb, cs, cc), _ = orc.inference_single_image([inp["logits"][0]], [inp["deltas"][0]], [inp["anchors"]], K)
    cnt = int(res["cand_count"][0].item())
    assert cnt == cb.shape[0] == 1000
    assert_equal_int(res["cand_classes"][0, :cnt], cc, "cand classes")
    assert_close_tensor(res["cand_boxes"][0, :cnt], cb, "cand boxes", atol_scale=1e-6)


def test_detect_small_topk_and_lvis_classes(cuda):
    _detect_case(cuda, 1, [900, 300], 1230, 43, topk=100)


def test_detect_select_paths(cuda):
    """Which select path served each slab: ordinary data goes through the sampled bar + one scan (status 1);
    plateaus of equal scores at the top-k boundary cannot be proven safe from a sample and are redone by the exact
    streaming kernel (status 2).  Both are checked against the oracle by the tests above/below; here only the routing."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    K = 80
    inp = synthetic.inference_inputs(49, 2, [24000, 3000, 300], K)
    res = fsg.ops.detect(inp["logits"].to(cuda), inp["deltas"].to(cuda), inp["anchors"].to(cuda),
                         inp["level_offsets"], want_candidates=True)
    assert res["slab_status"].tolist() == [[1, 1, 1], [1, 1, 1]]
    x = inp["logits"].clone()
    x[1, :24000].view(-1)[::7] = 2.5                  # a plateau far larger than top-k in image 1, level 0
    res = fsg.ops.detect(x.to(cuda), inp["deltas"].to(cuda), inp["anchors"].to(cuda), inp["level_offsets"],
                         want_candidates=True)
    st = res["slab_status"].tolist()
    assert st[1][0] == 2 and st[0] == [1, 1, 1] and st[1][1:] == [1, 1]
    offs = inp["level_offsets"]
    cls = [x[1, offs[i]:offs[i + 1]] for i in range(3)]
    reg = [inp["deltas"][1, offs[i]:offs[i + 1]] for i in range(3)]
    anc = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(3)]
    _, (cb, cs, cc), _ = orc.inference_single_image(cls, reg, anc, K)
    cnt = int(res["cand_count"][1].item())
    assert cnt == cb.shape[0]
    assert_equal_int(res["cand_classes"][1, :cnt], cc, "cand classes")
    assert_close_tensor(res["cand_boxes"][1, :cnt], cb, "cand boxes", atol_scale=1e-6)


@pytest.mark.parametrize("K,counts,A", [(80, [(20, 28), (10, 14), (5, 7), (3, 4), (2, 2)], 3),
                                        (7, [(33, 21), (9, 5)], 3), (3, [(41, 37), (5, 3)], 9)])
def test_detect_native_layout_equals_flat(cuda, K, counts, A):
    """fsg_detect_levels on the head's (N, A*K, H, W) / (N, A*4, H, W) outputs == fsg_detect on the permuted copy,
    bit for bit (candidates, order, detections); odd K / H*W make slabs that do not start on 16-byte boundaries.
    Includes a block of tied logits so that index order in the reference's (h, w, a, k) flattening matters."""
    fsg = _fsg()
    N = 3
    g = torch.Generator().manual_seed(50 + K)
    xs = [torch.randn((N, A * K, h, w), generator=g) * 1.5 - 2.0 for h, w in counts]
    ds = [torch.randn((N, A * 4, h, w), generator=g) * 0.2 for h, w in counts]
    xs[0].view(-1)[torch.randperm(xs[0].numel(), generator=g)[:4000]] = 1.25     # ties across planes
    R = sum(h * w * A for h, w in counts)
    cx, cy = torch.rand(R, generator=g) * 600, torch.rand(R, generator=g) * 400
    sz = torch.rand(R, generator=g) * 100 + 16
    anchors = torch.stack((cx - sz / 2, cy - sz / 2, cx + sz / 2, cy + sz / 2), dim=1).to(torch.float32)
    xs_c, ds_c, an = [t.to(cuda) for t in xs], [t.to(cuda) for t in ds], anchors.to(cuda)
    got = fsg.ops.detect_levels(xs_c, ds_c, an, K, topk=300, want_candidates=True)
    offs = [0]
    for h, w in counts:
        offs.append(offs[-1] + h * w * A)
    want = fsg.ops.detect(fsg.ops.levels_to_flat(xs_c, K), fsg.ops.levels_to_flat(ds_c, 4), an, offs, topk=300,
                          want_candidates=True)
    for k in ("count", "cand_count"):
        assert torch.equal(got[k], want[k]), k
    for n in range(N):
        c, d = int(want["cand_count"][n]), int(want["count"][n])
        for k in ("cand_boxes", "cand_scores", "cand_classes"):
            assert torch.equal(got[k][n, :c], want[k][n, :c]), k
        for k in ("boxes", "scores", "classes", "keep_idx"):
            assert torch.equal(got[k][n, :d], want[k][n, :d]), k
    if K != 80:
        return   # (tiny K: thousands of near-equal scores, whose order may differ by one ulp of the CPU's / GPU's expf)
    # and the flat result against the oracle for image 0
    flat_x, flat_d = orc.levels_to_flat(xs, K), orc.levels_to_flat(ds, 4)
    cls = [flat_x[0, offs[i]:offs[i + 1]] for i in range(len(counts))]
    reg = [flat_d[0, offs[i]:offs[i + 1]] for i in range(len(counts))]
    anc = [anchors[offs[i]:offs[i + 1]] for i in range(len(counts))]
    _, (cb, cs, cc), _ = orc.inference_single_image(cls, reg, anc, K, topk_candidates=300)
    c = int(got["cand_count"][0])
    assert c == cb.shape[0]
    assert_equal_int(got["cand_classes"][0, :c], cc, "cand classes")
    assert_close_tensor(got["cand_boxes"][0, :c], cb, "cand boxes", atol_scale=1e-6)


def test_inference_with_fused_detector_postprocess(cuda):
    """RetinaNet.forward's inference tail (retinanet.py:150-157): inference -> detector_postprocess per image, with
    the post-processing fused into the NMS epilogue; against oracle inference + oracle detector_postprocess."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, K, A = 3, 80, 3
    inp = synthetic.train_inputs(31, N, 256, 320, K, logits=False)
    gen = torch.Generator().manual_seed(32)
    cls_l = _levels(inp, K, N, gen, 1.5, synthetic.PRIOR_LOGIT + 1.0)
    reg_l = _levels(inp, 4, N, gen, 0.5)
    # push some boxes out of the image so that clipping empties them
    reg_l[0][:, 0::4] += 3.0
    offs = inp["level_offsets"]
    anc_levels = [fsg.Boxes(inp["anchors"][offs[i]:offs[i + 1]].to(cuda)) for i in range(5)]
    path = fsg.RetinaNetDensePath(num_classes=K)
    image_sizes = [(256, 320), (250, 300), (256, 310)]
    output_sizes = [(480, 640), (125, 150), (1000, 333)]
    got = path.inference([t.to(cuda) for t in cls_l], [t.to(cuda) for t in reg_l], [anc_levels] * N, image_sizes,
                         output_sizes=output_sizes)
    plain = path.inference([t.to(cuda) for t in cls_l], [t.to(cuda) for t in reg_l], [anc_levels] * N, image_sizes)
    dropped = 0
    for n in range(N):
        cls = [orc.nchw_to_n_hwa_k(t, K)[n] for t in cls_l]
        reg = [orc.nchw_to_n_hwa_k(t, 4)[n] for t in reg_l]
        anc = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(5)]
        (wb, ws, wc), _, _ = orc.inference_single_image(cls, reg, anc, K)
        # the GPU's own pre-postprocess detections must match the oracle's inference (scores within tolerance)
        assert_equal_int(plain[n].pred_classes, wc, "classes before postprocess")
        pb, ps, pc = orc.detector_postprocess(plain[n].pred_boxes.tensor.cpu(), plain[n].scores.cpu(),
                                              plain[n].pred_classes.cpu(), image_sizes[n], *output_sizes[n])
        assert tuple(got[n].image_size) == output_sizes[n]
        assert torch.equal(got[n].pred_boxes.tensor.cpu(), pb)
        assert torch.equal(got[n].scores.cpu(), ps)
        assert_equal_int(got[n].pred_classes, pc, "classes after postprocess")
        # and the stand-alone drop-in on the un-postprocessed Instances gives the same thing
        alone = fsg.detector_postprocess(plain[n], *output_sizes[n])
        assert torch.equal(alone.pred_boxes.tensor, got[n].pred_boxes.tensor)
        dropped += len(plain[n]) - len(got[n])
    assert dropped > 0, "test inputs should make at least one detection empty after clipping"


@pytest.mark.parametrize("detach,coeffs", [(False, (1.0, 1.0, -1.0)), (True, (0.0, 0.0, 1.0))])
def test_fused_step_from_head_outputs(cuda, detach, coeffs):
    """dense_train_step_levels: GT assignment + all three losses + backward from the per-level conv outputs and
    betting maps, gradients delivered in the same per-level layout (no permute/cat copy anywhere)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, K, A = 2, 80, 3
    inp = synthetic.train_inputs(51, N, 256, 320, K, logits=False)
    gen = torch.Generator().manual_seed(52)
    cls_l = _levels(inp, K, N, gen, 1.0, synthetic.PRIOR_LOGIT)
    reg_l = _levels(inp, 4, N, gen, 0.1)
    bet_l = [torch.sigmoid(torch.randn((N, A, h, w), generator=gen) - 4.0) for (h, w) in inp["grids"]]
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], orc.levels_to_flat(cls_l, K),
                          orc.levels_to_flat(reg_l, 4), orc.levels_to_flat(bet_l, 1).reshape(N, -1), K, *coeffs,
                          detach_pred=detach)
    gx = [t.to(cuda).requires_grad_(True) for t in cls_l]
    gd = [t.to(cuda).requires_grad_(True) for t in reg_l]
    gb = [t.to(cuda).requires_grad_(True) for t in bet_l]
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    res = fsg.dense_train_step_levels(gx, gd, gb, inp["anchors"].to(cuda), gt, fsg.DenseLossConfig(num_classes=K),
                                      coeffs, detach_pred=detach)
    (res.total * 2.0).backward()
    assert_equal_int(res.gt_classes, want["gt_classes"], "gt_classes")
    assert_equal_int(res.mask, want["mask"], "mask")
    assert_close_scalar(res.total.item(), want["total"], "total", rtol=2e-5)
    assert_close_scalar(res.gambler_loss.item(), want["gambler_loss"], "gambler_loss")
    for a, b in zip(res.per_anchor_loss, orc.flat_to_nahw(want["per_anchor_loss"], inp["grids"], A)):
        assert_close_tensor(a, b, "NAKHW_loss")
    if detach:
        assert all(t.grad is None for t in gx)
    else:
        assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gx], K), 2.0 * want["grad_logits"], "grad_logits")
    if coeffs[1] != 0:
        assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gd], 4), 2.0 * want["grad_deltas"], "grad_deltas")
    assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gb], 1).reshape(N, -1), 2.0 * want["grad_bets"],
                        "grad_bets", atol_scale=1e-6)


# ------------------------------------------------------------------------------- two-stage callers (8f row 4)
@pytest.mark.parametrize("N,counts,pre,post,thr,min_side", [
    (2, [6000, 2500, 700, 200, 60], 1000, 1000, 0.7, 0.0),        # FPN test-time setting
    (3, [20000, 5000, 2100, 819, 231], 2000, 1000, 0.7, 4.0),     # FPN train-time setting, min size filter
    (1, [300, 40], 6000, 50, 0.5, 0.0),                           # levels shorter than pre_nms_topk
    (2, [5000], 4096, 2000, 0.9, 0.0),                            # single level (C4-style), power-of-two k
])
def test_find_top_rpn_proposals(cuda, N, counts, pre, post, thr, min_side):
    """Batched RPN proposal selection vs the oracle, with exact logit ties (lower index first)."""
    from full_scale_gambler_for_object_detection_b200 import proposals as P, synthetic
    inp = synthetic.rpn_inputs(70 + N + len(counts), N, counts, ties=True)
    want = orc.find_top_rpn_proposals(inp["proposals"], inp["logits"], inp["image_sizes"], thr, pre, post, min_side)
    res = _fsg().ops.rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                   inp["image_sizes"], thr, pre, post, min_side)
    for n in range(N):
        c = int(res["count"][n].item())
        assert c == want[n][0].shape[0], "image %d: %d vs %d proposals" % (n, c, want[n][0].shape[0])
        assert torch.equal(res["boxes"][n, :c].cpu(), want[n][0])
        assert torch.equal(res["logits"][n, :c].cpu(), want[n][1])
        assert_equal_int(res["levels"][n, :c], want[n][2], "levels")
        assert float(res["boxes"][n, c:].abs().sum()) == 0.0
    got = P.find_top_rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                   inp["image_sizes"], thr, pre, post, min_side, training=True)
    assert all(len(got[n]) == want[n][0].shape[0] for n in range(N))


def test_find_top_rpn_proposals_c4_setting(cuda):
    """RPN.PRE_NMS_TOPK_TRAIN = 12000 / POST 2000 on one feature level (the reference's default for C4 models,
    config/defaults.py:219-224): more candidates than the shared-memory NMS holds -> general-n NMS path."""
    from full_scale_gambler_for_object_detection_b200 import synthetic
    inp = synthetic.rpn_inputs(78, 2, [50 * 84 * 15], ties=True)
    want = orc.find_top_rpn_proposals(inp["proposals"], inp["logits"], inp["image_sizes"], 0.7, 12000, 2000, 0.0)
    res = _fsg().ops.rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                   inp["image_sizes"], 0.7, 12000, 2000, 0.0)
    for n in range(2):
        c = int(res["count"][n].item())
        assert c == want[n][0].shape[0], "image %d: %d vs %d proposals" % (n, c, want[n][0].shape[0])
        assert torch.equal(res["boxes"][n, :c].cpu(), want[n][0])
        assert torch.equal(res["logits"][n, :c].cpu(), want[n][1])
        assert float(res["boxes"][n, c:].abs().sum()) == 0.0


def test_rpn_topk_all_equal_logits(cuda):
    """Degenerate row: every logit equal -> the radix select runs through the index digits; lowest indices win."""
    from full_scale_gambler_for_object_detection_b200 import synthetic
    inp = synthetic.rpn_inputs(79, 1, [3000], ties=False)
    inp["logits"][0][:] = 0.25
    want = orc.find_top_rpn_proposals(inp["proposals"], inp["logits"], inp["image_sizes"], 0.7, 500, 500, 0.0)
    res = _fsg().ops.rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                   inp["image_sizes"], 0.7, 500, 500, 0.0)
    c = int(res["count"][0].item())
    assert c == want[0][0].shape[0] and torch.equal(res["boxes"][0, :c].cpu(), want[0][0])


def test_rpn_and_roi_ground_truth(cuda):
    """RPNOutputs._get_ground_truth and the matching part of label_and_sample_proposals vs the oracle; sampling
    (random) checked through its invariants."""
    from full_scale_gambler_for_object_detection_b200 import proposals as P, synthetic
    inp = synthetic.train_inputs(81, 3, 320, 320, 80, M=7)
    wl, wd = orc.rpn_ground_truth(inp["anchors"], inp["gt_boxes"])
    gl, gd = P.rpn_ground_truth(inp["anchors"].to(cuda), [b.to(cuda) for b in inp["gt_boxes"]])
    for n in range(3):
        assert_equal_int(gl[n], wl[n], "rpn labels")
        assert_close_tensor(gd[n], wd[n], "rpn deltas")
    pos, neg = P.subsample_labels(gl[0], 256, 0.5, 0)
    npos = int((wl[0] == 1).sum())
    assert pos.numel() == min(npos, 128) and neg.numel() == 256 - pos.numel()
    assert bool((gl[0][pos] == 1).all()) and bool((gl[0][neg] == 0).all())
    assert pos.unique().numel() == pos.numel() and neg.unique().numel() == neg.numel()
    # ROI heads: proposals (+ appended GT boxes) against the GT of one image, Matcher([0.5], [0, 1], no low-quality)
    rp = synthetic.rpn_inputs(82, 1, [1500], image_hw=(320, 320), ties=False)["proposals"][0][0]
    props = torch.cat([rp, inp["gt_boxes"][0]])
    for gtb, gtc in ((inp["gt_boxes"][0], inp["gt_classes"][0]), (torch.zeros((0, 4)), torch.zeros(0, dtype=torch.int64))):
        wm, wlab, wc = orc.label_proposals(props, gtb, gtc, 80)
        gm, glab, gc = P.label_proposals(props.to(cuda), gtb.to(cuda), gtc.to(cuda), 80)
        assert_equal_int(gm, wm, "matched idxs")
        assert_equal_int(glab, wlab, "matched labels")
        assert_equal_int(gc, wc, "gt classes")


@pytest.mark.parametrize("R,K,spec,sthr,topk", [(1000, 80, True, 0.05, 100), (700, 20, False, 0.02, 100),
                                                 (300, 1230, True, 0.0001, 300), (50, 5, True, 0.9999, -1)])
def test_fast_rcnn_inference_single_image(cuda, R, K, spec, sthr, topk):
    from full_scale_gambler_for_object_detection_b200 import proposals as P, synthetic
    inp = synthetic.fast_rcnn_inputs(90 + K, R, K, spec)
    wb, ws, wc, wr = orc.fast_rcnn_inference_single_image(inp["boxes"], inp["scores"], inp["image_shape"], sthr, 0.5, topk)
    r, rows = P.fast_rcnn_inference_single_image(inp["boxes"].to(cuda), inp["scores"].to(cuda), inp["image_shape"],
                                                 sthr, 0.5, topk)
    assert torch.equal(r.pred_boxes.tensor.cpu(), wb) and torch.equal(r.scores.cpu(), ws)
    assert_equal_int(r.pred_classes, wc, "classes")
    assert_equal_int(rows, wr, "rows")


def test_native_layout_plan_and_graph(cuda):
    """DenseStepPlanLevels: direct run and CUDA-graph replay (with refreshed inputs) equal dense_train_step_levels."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    N, K, A = 2, 80, 3
    inp = synthetic.train_inputs(61, N, 256, 320, K, logits=False)
    gen = torch.Generator().manual_seed(62)
    mk = lambda: ([t.to(cuda) for t in _levels(inp, K, N, gen, 1.0, synthetic.PRIOR_LOGIT)],
                  [t.to(cuda) for t in _levels(inp, 4, N, gen, 0.1)],
                  [torch.sigmoid(torch.randn((N, A, h, w), generator=gen) - 4.0).to(cuda) for (h, w) in inp["grids"]])
    xs, ds, bs = mk()
    anchors = inp["anchors"].to(cuda)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    cfg = fsg.DenseLossConfig(num_classes=K)
    plan = fsg.DenseStepPlanLevels(N, inp["grids"], A, K, cfg, cuda)

    def reference(xs, ds, bs):
        gx = [t.clone().requires_grad_(True) for t in xs]
        gd = [t.clone().requires_grad_(True) for t in ds]
        gb = [t.clone().requires_grad_(True) for t in bs]
        r = fsg.dense_train_step_levels(gx, gd, gb, anchors, gt, cfg)
        r.total.backward()
        return r, gx, gd, gb

    def check(res, ref):
        r, gx, gd, gb = ref
        assert torch.equal(res.scalars, r.scalars) and torch.equal(res.gt_classes, r.gt_classes)
        for a, b in zip(plan.grad_logits, gx):
            assert torch.equal(a, b.grad)
        for a, b in zip(plan.grad_deltas, gd):
            assert torch.equal(a, b.grad)
        for a, b in zip(plan.grad_bets, gb):
            assert torch.equal(a, b.grad)
        for a, b in zip(plan.nakhw_loss, r.per_anchor_loss):
            assert torch.equal(a, b)

    check(plan.run(xs, ds, bs, anchors, gt), reference(xs, ds, bs))
    plan.capture(xs, ds, bs, anchors, gt)
    xs2, ds2, bs2 = mk()                       # new values, same storage: refresh in place and replay
    for dst, src in zip(xs + ds + bs, xs2 + ds2 + bs2):
        dst.copy_(src)
    res = plan.replay()
    torch.cuda.synchronize()
    check(res, reference(xs, ds, bs))


@pytest.mark.parametrize("M,shards", [(200, 3), (8, 2), (1500, 2)])
def test_match_with_anchors_sharded_by_range(cuda, M, shards):
    """one image whose anchors are split by range over `shards` ranks.  Emulated in one process:
    phase 1 (pass A) per shard with its own workspace, MAX of the per-GT maxima over the shards (what
    sharded.all_reduce_gt_max does over NCCL), phase 2 (pass B) per shard; the concatenation must equal the
    unsharded call and the oracle, bit for bit."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import sharded, synthetic
    R = 40001
    inp = synthetic.matcher_stress_inputs(33, 1, R, M)
    anchors = inp["anchors"][0]
    anchors[7] = inp["gt_boxes"][0][3]                              # IoU 1 with GT 3, lives in shard 0
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    want_keys = ("matches", "match_labels", "gt_classes")
    whole = fsg.ops.match_anchors(anchors.to(cuda), gt, 80, want=want_keys, picky_thresholds=None)
    firsts = []
    for r in range(shards):
        lo, hi = sharded.anchor_range(R, shards, r)
        firsts.append(fsg.ops.match_anchors(anchors[lo:hi].to(cuda), gt, 80, want=want_keys, picky_thresholds=None,
                                            phases=1))
    gmax = torch.stack([f["gt_max_bits"] for f in firsts]).max(dim=0).values
    parts = []
    for r, f in enumerate(firsts):
        lo, hi = sharded.anchor_range(R, shards, r)
        f["gt_max_bits"].copy_(gmax)
        parts.append(fsg.ops.match_anchors(anchors[lo:hi].to(cuda), gt, 80, want=want_keys, picky_thresholds=None,
                                           phases=2, workspace=f["workspace"], out={k: f[k] for k in want_keys}))
    oracle = orc.ground_truth([anchors], inp["gt_boxes"], inp["gt_classes"], 80)
    for k in want_keys:
        got = torch.cat([p[k] for p in parts], dim=1)
        assert torch.equal(got, whole[k]), k
        assert_equal_int(got, oracle[k], k)


def test_native_layout_with_nine_anchors_per_cell(cuda):
    """Upstream RetinaNet geometry (3 scales x 3 aspect ratios, A = 9): K1 + K2 on the native layout against the
    oracle, with the anchors produced by the device-side generator."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import anchor_generator as ag, synthetic
    N, K, A, H, W = 2, 80, 9, 224, 288
    gen_a = fsg.DefaultAnchorGenerator([list(s) for s in ag.RETINANET_SIZES], [[0.5, 1.0, 2.0]], ag.RETINANET_STRIDES, cuda)
    grids = ag.retinanet_grid_sizes(H, W)
    anchors, offs = gen_a.flat_for_grids(grids)
    want_anchors, _, _ = ag.retinanet_anchors(H, W, aspect_ratios=((0.5, 1.0, 2.0),) * 5)
    assert torch.equal(anchors.cpu(), want_anchors)
    R = anchors.shape[0]
    g = torch.Generator().manual_seed(71)
    base = synthetic.train_inputs(71, N, H, W, K, M=6, logits=False)
    cls_l = [torch.randn((N, A * K, h, w), generator=g) + synthetic.PRIOR_LOGIT for h, w in grids]
    reg_l = [torch.randn((N, A * 4, h, w), generator=g) * 0.1 for h, w in grids]
    bet_l = [torch.sigmoid(torch.randn((N, A, h, w), generator=g) - 4.0) for h, w in grids]
    bets_flat = orc.levels_to_flat(bet_l, 1).reshape(N, R)
    want = orc.train_step(want_anchors, base["gt_boxes"], base["gt_classes"], orc.levels_to_flat(cls_l, K),
                          orc.levels_to_flat(reg_l, 4), bets_flat, K, 1.0, 1.0, -1.0)
    gx = [t.to(cuda).requires_grad_(True) for t in cls_l]
    gd = [t.to(cuda).requires_grad_(True) for t in reg_l]
    gb = [t.to(cuda).requires_grad_(True) for t in bet_l]
    gt = fsg.ops.PackedGT.from_lists(base["gt_boxes"], base["gt_classes"], cuda)
    res = fsg.dense_train_step_levels(gx, gd, gb, anchors, gt, fsg.DenseLossConfig(num_classes=K))
    res.total.backward()
    assert_equal_int(res.gt_classes, want["gt_classes"], "gt_classes")
    assert_equal_int(res.mask, want["mask"], "mask")
    assert_close_scalar(res.total.item(), want["total"], "total", rtol=2e-5)
    assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gx], K), want["grad_logits"], "grad_logits")
    assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gd], 4), want["grad_deltas"], "grad_deltas")
    assert_close_tensor(orc.levels_to_flat([t.grad.cpu() for t in gb], 1).reshape(N, R), want["grad_bets"], "grad_bets",
                        atol_scale=1e-6)
