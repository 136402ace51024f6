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
    """16 x 800x1333 (R = 67200), K = 80: determinism, linearity of the fused backward in the loss
    coefficients, and oracle parity of matching + per-anchor loss on the first and the GT-free image."""
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
    # oracle parity on image 0 and the GT-free image
    for n in (0, N - 1):
        want = orc.ground_truth(inp["anchors"], [inp["gt_boxes"][n]], [inp["gt_classes"][n]], K)
        assert_equal_int(gtc[n:n + 1], want["gt_classes"], "gt_classes image %d" % n)
        assert_equal_int(mask[n:n + 1], want["mask"], "mask image %d" % n)
        t, _ = orc.one_hot_targets(want["gt_classes"].flatten(), K, inp["logits"][n])
        f = orc.cls_loss_elementwise(inp["logits"][n], t, "focal", 0.25, 2.0)
        f = (f * (want["gt_classes"].flatten() >= 0)[:, None]).sum(dim=1)
        assert_close_tensor(ell[n], f, "per-anchor loss image %d" % n)


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
