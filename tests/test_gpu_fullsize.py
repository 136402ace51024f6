"""GPU parity at the BASELINE.json sizes (per-image sizes are the full ones; batch counts are trimmed where the
CPU oracle would otherwise need minutes), plus size-independent properties on the full config-2 batch."""
import pytest
import torch

from oracle import dense_oracle as orc
from tests.util import assert_close_scalar, assert_close_tensor, assert_equal_int

pytestmark = pytest.mark.gpu


def _fsg():
    import full_scale_gambler_for_object_detection_b200 as fsg
    return fsg


def test_config2_full_batch_properties(cuda):
    """16 x 800x1333 (R = 67200), K = 80 -- the benchmarked batch itself: determinism, linearity of the fused
    backward in the loss coefficients, and oracle parity of EVERY output (labels, mask, losses, per-anchor loss,
    d/d logits, d/d deltas, d/d bets) on all 16 images."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    N, K = 16, 80
    inp = synthetic.train_inputs(2, N, 800, 1333, K)
    R = inp["R"]
    assert R == 67200
    cfg = fsg.DenseLossConfig(num_classes=K)
    x, d, b = (inp[k].to(cuda) for k in ("logits", "deltas", "bets"))
    anchors = inp["anchors"].to(cuda)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)

    def run(coeffs):
        plan = fsg.DenseStepPlan(N, R, K, cfg, cuda, coeffs)
        r = plan.run(x, d, b, anchors, gt)
        return (r.scalars.clone(), plan.grad_logits.clone(), plan.grad_deltas.clone() if plan.grad_deltas is not None
                else None, plan.grad_bets.clone(), r.gt_classes.clone(), r.mask.clone(), r.per_anchor_loss.clone(),
                r.stats.clone())

    s1, gl1, gd1, gb1, gtc, mask, ell, stats = run((1.0, 1.0, -1.0))
    s2, gl2, gd2, gb2, *_ = run((1.0, 1.0, -1.0))
    assert torch.equal(s1, s2) and torch.equal(gl1, gl2) and torch.equal(gb1, gb2)          # run-to-run identical
    # linearity: grad(1,1,-1) = grad(1,0,0) + grad(0,1,0) - grad(0,0,1)
    sa, gla, _, gba, *_ = run((1.0, 0.0, 0.0))
    sb, glb, gdb, gbb, *_ = run((0.0, 1.0, 0.0))
    sc, glc, _, gbc, *_ = run((0.0, 0.0, 1.0))
    assert_close_tensor(gla + glb - glc, gl1, "grad_logits linearity", rtol=1e-5, atol_scale=1e-6)
    assert_close_tensor(gdb, gd1, "grad_deltas linearity", rtol=1e-6)
    assert_close_tensor(gba + gbb - gbc, gb1, "grad_bets linearity", rtol=1e-5, atol_scale=1e-6)
    assert_close_scalar(s1[8], s1[5] + s1[6] - s1[7], "total = cls + reg - gambler", rtol=1e-12)
    assert float(glb.abs().max()) == 0.0 and float(gba.abs().max()) == 0.0     # reg touches no logits, cls no bets
    # structure
    assert int(stats[0]) == int(((gtc >= 0) & (gtc != K)).sum())
    assert bool((gtc[-1] == K).all()) and bool((mask[-1] == K).all())          # GT-free image (retinanet.py:362,425)
    assert float(gl1[gtc < 0].abs().max()) == 0.0                              # ignored anchors: zero gradient
    # oracle parity on the WHOLE benchmarked batch: every integer output bit-exact, the three losses, the per-anchor
    # loss and all three gradients within the north_star tolerance (the oracle needs a few seconds for 16 images)
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                          inp["bets"], K, 1.0, 1.0, -1.0)
    assert_equal_int(gtc, want["gt_classes"], "gt_classes")
    assert_equal_int(mask, want["mask"], "mask")
    assert int(stats[0]) == int(want["num_foreground"])
    for k, j in (("loss_cls", 5), ("loss_box_reg", 6), ("gambler_loss", 7), ("total", 8)):
        assert_close_scalar(s1[j], want[k], k)
    assert_close_tensor(ell, want["per_anchor_loss"], "per_anchor_loss")
    assert_close_tensor(gl1, want["grad_logits"], "grad_logits")
    assert_close_tensor(gd1, want["grad_deltas"], "grad_deltas")
    assert_close_tensor(gb1, want["grad_bets"], "grad_bets", atol_scale=1e-6)


def test_config3_lvis_slice(cuda):
    """LVIS: K = 1230 at the full 800x1333 geometry, 2 images (the per-GPU shard of config 3 is 8)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    N, K = 2, 1230
    inp = synthetic.train_inputs(3, N, 800, 1333, K)
    cfg = fsg.DenseLossConfig(num_classes=K)
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                          inp["bets"], K, 1.0, 1.0, -1.0)
    x = inp["logits"].to(cuda).requires_grad_(True)
    d = inp["deltas"].to(cuda).requires_grad_(True)
    b = inp["bets"].to(cuda).requires_grad_(True)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    res = fsg.dense_train_step(x, d, b, inp["anchors"].to(cuda), gt, cfg)
    res.total.backward()
    assert_equal_int(res.gt_classes, want["gt_classes"], "gt_classes")
    assert_equal_int(res.mask, want["mask"], "mask")
    for k, v in (("loss_cls", res.loss_cls), ("loss_box_reg", res.loss_box_reg), ("gambler_loss", res.gambler_loss)):
        assert_close_scalar(v.item(), want[k], k)
    assert_close_tensor(res.per_anchor_loss, want["per_anchor_loss"], "per_anchor_loss")
    assert_close_tensor(x.grad, want["grad_logits"], "grad_logits")
    assert_close_tensor(d.grad, want["grad_deltas"], "grad_deltas")
    assert_close_tensor(b.grad, want["grad_bets"], "grad_bets", atol_scale=1e-6)


def test_config4_full_image_size(cuda):
    """Inference at 5 x 24000 anchors x 80 classes per image (config 4), 2 images."""
    from tests.test_gpu_parity import _detect_case

    _detect_case(cuda, 2, [24000] * 5, 80, 4)


def test_config5_full_image_size(cuda):
    """Matcher stress at the full per-image size: 200 GT x 1 000 000 anchors, low-quality matches on."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    inp = synthetic.matcher_stress_inputs(5, 1, 1000000, 200)
    want = orc.ground_truth([inp["anchors"][0]], inp["gt_boxes"], inp["gt_classes"], 80)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    got = fsg.ops.match_anchors(inp["anchors"].to(cuda), gt, 80,
                                want=("matches", "match_labels", "picky_labels", "gt_classes", "mask"))
    for k in ("matches", "match_labels", "picky_labels", "gt_classes", "mask"):
        assert_equal_int(got[k], want[k], k)
    assert int((want["match_labels"] == 1).sum()) >= 200


def test_config2_full_batch_native_layout(cuda):
    """Config 2 at full size on the head's native layout: the step from per-level conv outputs against the (N, R, K)
    step on the permuted copy -- same integer outputs; gradients and losses equal up to the order of the
    K-reduction and of the bet normaliser's sum (the in-place form adds R*T once instead of T per anchor)."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    N, K = 16, 80
    inp = synthetic.train_inputs(2, N, 800, 1333, K, logits=False)
    A, grids = inp["A"], inp["grids"]
    g = torch.Generator(device=cuda).manual_seed(9)
    xs = [(torch.randn((N, A * K, h, w), device=cuda, generator=g) + synthetic.PRIOR_LOGIT).requires_grad_(True)
          for h, w in grids]
    ds = [(torch.randn((N, A * 4, h, w), device=cuda, generator=g) * 0.1).requires_grad_(True) for h, w in grids]
    bs = [torch.sigmoid(torch.randn((N, A, h, w), device=cuda, generator=g) + synthetic.PRIOR_LOGIT).requires_grad_(True)
          for h, w in grids]
    anchors = inp["anchors"].to(cuda)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    cfg = fsg.DenseLossConfig(num_classes=K)
    res = fsg.dense_train_step_levels(xs, ds, bs, anchors, gt, cfg)
    res.total.backward()
    # the same step on the permuted copy
    xf = fsg.ops.levels_to_flat([t.detach() for t in xs], K).requires_grad_(True)
    df = fsg.ops.levels_to_flat([t.detach() for t in ds], 4).requires_grad_(True)
    bf = fsg.ops.anchor_maps_to_flat([[t.detach() for t in bs]])[0].requires_grad_(True)
    ref = fsg.dense_train_step(xf, df, bf, anchors, gt, cfg)
    ref.total.backward()
    assert torch.equal(res.gt_classes, ref.gt_classes) and torch.equal(res.mask, ref.mask)
    assert float(res.stats[0]) == float(ref.stats[0])
    assert_close_tensor(res.stats, ref.stats, "stats", rtol=1e-6)
    assert_close_tensor(fsg.ops.levels_to_flat([t.grad for t in xs], K), xf.grad, "grad_logits", rtol=2e-6)
    assert torch.equal(fsg.ops.levels_to_flat([t.grad for t in ds], 4), df.grad)
    for k in (5, 6, 7, 8):
        assert_close_scalar(res.scalars[k].item(), ref.scalars[k].item(), "scalar %d" % k, rtol=1e-6)
    assert_close_tensor(fsg.ops.anchor_maps_to_flat([[t.grad for t in bs]])[0], bf.grad, "grad_bets", rtol=1e-5,
                        atol_scale=1e-6)
    ell = fsg.ops.anchor_maps_to_flat([res.per_anchor_loss])[0]
    assert_close_tensor(ell, ref.per_anchor_loss, "per_anchor_loss", rtol=2e-6)
    # the generated anchors are the ones the synthetic batch was built with
    gen = fsg.DefaultAnchorGenerator([list(s) for s in fsg.anchor_generator.RETINANET_SIZES], [[1.0]],
                                     fsg.anchor_generator.RETINANET_STRIDES, cuda)
    flat, offs = gen.flat(xs)
    assert torch.equal(flat, anchors) and offs == inp["level_offsets"]


def test_rpn_full_size_properties(cuda):
    """find_top_rpn_proposals at the FPN training size (P2..P6 of 800x1333, A = 3, 2000 -> 1000): size-independent
    properties on 4 images -- sorted logits, boxes inside the image and larger than the minimum side, every kept
    proposal among its level's top-k, no kept pair of one level above the NMS threshold; image 0 against the oracle."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    counts = [200 * 336 * 3, 100 * 168 * 3, 50 * 84 * 3, 25 * 42 * 3, 13 * 21 * 3]
    N, pre, post, thr, min_side = 4, 2000, 1000, 0.7, 2.5
    inp = synthetic.rpn_inputs(7, N, counts, ties=False)
    res = fsg.ops.rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                inp["image_sizes"], thr, pre, post, min_side)
    H, W = inp["image_sizes"][0]
    for n in range(N):
        c = int(res["count"][n])
        assert 0 < c <= post
        b, lg, lv = res["boxes"][n, :c], res["logits"][n, :c], res["levels"][n, :c]
        assert bool((lg[:-1] >= lg[1:]).all())
        assert bool((b[:, 0] >= 0).all() and (b[:, 1] >= 0).all() and (b[:, 2] <= W).all() and (b[:, 3] <= H).all())
        assert bool(((b[:, 2] - b[:, 0]) > min_side).all() and ((b[:, 3] - b[:, 1]) > min_side).all())
        for l in range(len(counts)):
            sel = lv == l
            if int(sel.sum()) == 0:
                continue
            kth = torch.topk(inp["logits"][l][n].to(cuda), min(pre, counts[l])).values[-1]
            assert bool((lg[sel] >= kth).all())
            iou = fsg.ops.pairwise_iou(b[sel], b[sel])
            iou.fill_diagonal_(0.0)
            assert float(iou.max()) <= thr
    want = orc.find_top_rpn_proposals([p[:1] for p in inp["proposals"]], [x[:1] for x in inp["logits"]],
                                      inp["image_sizes"][:1], thr, pre, post, min_side)
    c0 = int(res["count"][0])
    assert c0 == want[0][0].shape[0]
    assert torch.equal(res["boxes"][0, :c0].cpu(), want[0][0]) and torch.equal(res["logits"][0, :c0].cpu(), want[0][1])
