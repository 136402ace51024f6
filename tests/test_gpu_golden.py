"""GPU: the CUDA path against the REFERENCE's own outputs committed under tests/golden/ (made by
oracle/make_golden.py from the reference's source files) -- no oracle in between."""
import pytest
import torch

from tests import golden_util as gu
from tests.util import assert_close_scalar, assert_close_tensor, assert_equal_int

pytestmark = pytest.mark.gpu


def _fsg():
    import full_scale_gambler_for_object_detection_b200 as fsg
    return fsg


@pytest.mark.parametrize("name", gu.TRAIN_CASES)
def test_train_step_vs_reference(cuda, name):
    fsg = _fsg()
    inp, g, coeffs, detach, gcfg, K = gu.train_case(name)
    cfg = fsg.DenseLossConfig(num_classes=K, **{gu.CFG_KW[k]: v for k, v in gcfg.items()})
    x = inp["logits"].to(cuda).requires_grad_(True)
    d = inp["deltas"].to(cuda).requires_grad_(True)
    b = inp["bets"].to(cuda).requires_grad_(True)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    res = fsg.dense_train_step(x, d, b, inp["anchors"].to(cuda), gt, cfg, coeffs, detach_pred=detach, want_weights=True)
    res.total.backward()
    assert_equal_int(res.gt_classes, g["gt_classes"], "gt_classes")
    assert_equal_int(res.mask, g["mask"], "mask")
    assert_close_scalar(res.loss_cls.item(), g["loss_cls"], "loss_cls")
    assert_close_scalar(res.loss_box_reg.item(), g["loss_box_reg"], "loss_box_reg")
    assert_close_scalar(res.gambler_loss.item(), g["gambler_loss"], "gambler_loss")
    assert_close_scalar(res.total.item(), g["total"], "total", rtol=2e-5)
    assert_close_scalar(res.loss_before_weighting(cfg.gambler_loss_mode).item(), g["loss_before_weighting"], "lbw")
    assert_close_scalar(res.lower_bound(cfg.gambler_temperature).item(), g["lower_bound"], "lower_bound")
    assert_close_tensor(res.per_anchor_loss, g["per_anchor_loss"], "per_anchor_loss")
    assert_close_tensor(res.weights, g["weights"], "weights")
    if detach:
        assert x.grad is None
    else:
        assert_close_tensor(x.grad.reshape(-1, K)[g["grad_rows"].to(cuda)], g["grad_logits"], "grad_logits")
    if coeffs[1] != 0:
        assert_close_tensor(d.grad, g["grad_deltas"], "grad_deltas")
    floor = 1e-5 if gcfg.get("GAMBLER_LOSS_MODE") == "sigmoid" else 1e-6   # see test_step_variants
    assert_close_tensor(b.grad, g["grad_bets"], "grad_bets", atol_scale=floor)


@pytest.mark.parametrize("native", [False, True])
def test_log_metrics_vs_reference(cuda, native):
    """StepResult.log_metrics (fsg_bet_stats: sum / max / mean of the masked bets, sum / max / mean / median of the
    normalised weights, no sort and no host sync) against what the reference's own calc_log_metrics source produced
    (train_net.py:1089-1124) -- from the flat step and from the step on the head's per-level layout."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    from oracle import dense_oracle as orc

    g = gu.load("log_metrics")
    cid, N, H, W, K, M = [int(v) for v in g["params"]]
    lam_reg, kappa, lam_out = [float(v) for v in g["lambdas"]]
    inp = synthetic.train_inputs(cid, N, H, W, K, M=M)
    cfg = fsg.DenseLossConfig(num_classes=K)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    anchors = inp["anchors"].to(cuda)
    if native:
        A, grids = inp["A"], inp["grids"]
        def lv(flat, C):     # (N, R, C) -> list[(N, A*C, H, W)], the inverse of retinanet.py:24-33
            out, off = [], 0
            for h, w in grids:
                n_l = h * w * A
                out.append(flat[:, off:off + n_l].reshape(N, h, w, A, C).permute(0, 3, 4, 1, 2)
                           .reshape(N, A * C, h, w).contiguous().to(cuda))
                off += n_l
            return out

        xs, ds = lv(inp["logits"], K), lv(inp["deltas"], 4)
        bs = [t.to(cuda) for t in orc.flat_to_nahw(inp["bets"], grids, A)]
        bs = [t.contiguous() for t in bs]
        res = fsg.dense_train_step_levels(xs, ds, bs, anchors, gt, cfg)
        got = res.log_metrics(bs, cfg, lam_reg, kappa, lam_out)
    else:
        b = inp["bets"].to(cuda)
        res = fsg.dense_train_step(inp["logits"].to(cuda), inp["deltas"].to(cuda), b, anchors, gt, cfg)
        got = res.log_metrics(b, cfg, lam_reg, kappa, lam_out)
    for name, want in zip([str(n) for n in g["names"]], g["values"].tolist()):
        assert_close_scalar(float(got[name]), want, name, rtol=2e-5 if "loss" in name else 1e-5)


def test_log_metrics_median_is_exact_rank(cuda):
    """The 3-pass radix select returns exactly the element torch.median picks (lower median) of the device-computed
    weights, for odd and even counts, with ties and with the GT-free image's mask = K quirk in play."""
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic
    for N, H, W in ((3, 128, 160), (2, 96, 96)):
        inp = synthetic.train_inputs(63, N, H, W, 80, M=4)
        cfg = fsg.DenseLossConfig(num_classes=80)
        b = inp["bets"].clone()
        b[0, ::3] = b[0, 0]                                   # many exact ties
        b = b.to(cuda)
        gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
        res = fsg.dense_train_step(inp["logits"].to(cuda), inp["deltas"].to(cuda), b, inp["anchors"].to(cuda), gt, cfg,
                                   want_weights=True)
        got = res.log_metrics(b, cfg)
        w = res.weights.flatten()
        assert float(got["visualized weights/median"]) == float(torch.median(w))
        assert float(got["visualized weights/max"]) == float(w.max())
        assert_close_scalar(float(got["visualized weights/sum"]), float(w.double().sum()), "sum", rtol=1e-9)
        bm = (b * res.mask).flatten()
        assert float(got["gambler_bets/max"]) == float(bm.max())
        assert_close_scalar(float(got["gambler_bets/sum"]), float(bm.double().sum()), "bets sum", rtol=1e-9)


def test_matcher_vs_reference(cuda):
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("matcher_stress")
    cid, N, R, M = [int(v) for v in g["params"]]
    inp = synthetic.matcher_stress_inputs(cid, N, R, M)
    inp["anchors"][0, 100] = inp["anchors"][0, 99]
    inp["anchors"][1, :50] = inp["gt_boxes"][1][:50]
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], cuda)
    got = fsg.ops.match_anchors(inp["anchors"].to(cuda), gt, 80, want=("matches", "match_labels", "picky_labels"))
    for i in range(N):
        assert_equal_int(got["matches"][i], g["matches_%d" % i], "fused matches")
        assert_equal_int(got["match_labels"][i], g["labels_%d" % i], "fused labels")
        assert_equal_int(got["picky_labels"][i], g["picky_labels_%d" % i], "fused picky labels")
        # drop-in forms: materialised IoU matrix + Matcher on it, with and without low-quality matches
        q = fsg.pairwise_iou(fsg.Boxes(inp["gt_boxes"][i].to(cuda)), fsg.Boxes(inp["anchors"][i].to(cuda)))
        assert torch.equal(q[:, :512].cpu(), g["iou_sample_%d" % i])
        assert torch.equal(q.max(dim=1).values.cpu(), g["iou_rowmax_%d" % i])
        m, l = fsg.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True)(q)
        assert_equal_int(m, g["matches_%d" % i], "matrix matches")
        assert_equal_int(l, g["labels_%d" % i], "matrix labels")
        m, l = fsg.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=False)(q)
        assert_equal_int(m, g["nolq_matches_%d" % i], "no-lq matches")
        assert_equal_int(l, g["nolq_labels_%d" % i], "no-lq labels")
        m, l = fsg.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=False).match_boxes(
            inp["gt_boxes"][i].to(cuda), inp["anchors"][i].to(cuda))
        assert_equal_int(m, g["nolq_matches_%d" % i], "fused no-lq matches")
        assert_equal_int(l, g["nolq_labels_%d" % i], "fused no-lq labels")
    m, l = fsg.Matcher([0.4, 0.5], [0, -1, 1], True)(torch.tensor([[0.9, 0.3, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]).to(cuda))
    assert_equal_int(m, g["quirk_matches"], "quirk matches")
    assert_equal_int(l, g["quirk_labels"], "quirk labels")
    m, l = fsg.Matcher([0.4, 0.5], [0, -1, 1], True)(torch.zeros((0, 7), device=cuda))
    assert_equal_int(m, g["empty_matches"], "empty matches")
    assert_equal_int(l, g["empty_labels"], "empty labels")


def test_box2box_vs_reference(cuda):
    fsg = _fsg()
    g = gu.load("box2box")
    t = fsg.Box2BoxTransform(weights=tuple(float(v) for v in g["weights"]))
    assert_close_tensor(t.get_deltas(g["src"].to(cuda), g["dst"].to(cuda)), g["deltas"], "get_deltas")
    assert_close_tensor(t.apply_deltas(g["big"].to(cuda), g["boxes"].to(cuda)), g["applied"], "apply_deltas",
                        atol_scale=1e-6)


def test_nms_vs_reference(cuda):
    """keep indices of the reference's nms / batched_nms (detectron2/layers/nms.py -> torchvision 0.26), bit-exact."""
    fsg = _fsg()
    g = gu.load("nms")
    for name in ("a", "b", "c"):
        b, s, i = g["boxes_" + name].to(cuda), g["scores_" + name].to(cuda), g["idxs_" + name].to(cuda)
        for thr in (0.2, 0.5, 0.8):
            tag = "%s_%02d" % (name, int(thr * 10))
            assert_equal_int(fsg.nms(b, s, thr), g["nms_" + tag], "nms " + tag)
            assert_equal_int(fsg.batched_nms(b, s, i, thr), g["batched_" + tag], "batched_nms " + tag)


def test_inference_vs_reference(cuda):
    fsg = _fsg()
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("inference")
    p = [int(v) for v in g["params"]]
    inp = synthetic.inference_inputs(p[0], p[1], p[2:7], p[7])
    path = fsg.RetinaNetDensePath(num_classes=p[7])
    offs = inp["level_offsets"]
    for n in range(p[1]):
        cls = [inp["logits"][n, offs[i]:offs[i + 1]].to(cuda) for i in range(5)]
        reg = [inp["deltas"][n, offs[i]:offs[i + 1]].to(cuda) for i in range(5)]
        anc = [fsg.Boxes(inp["anchors"][offs[i]:offs[i + 1]].to(cuda)) for i in range(5)]
        r = path.inference_single_image(cls, reg, anc, (800, 1344))
        # same detections (class, order); scores/boxes within fp32 tolerance of the CPU sigmoid/exp
        assert_equal_int(r.pred_classes, g["classes_%d" % n], "pred_classes")
        assert_close_tensor(r.scores, g["scores_%d" % n], "scores")
        assert_close_tensor(r.pred_boxes.tensor, g["boxes_%d" % n], "pred_boxes", atol_scale=1e-6)


def test_gpu_detector_postprocess_vs_reference(cuda):
    """fsg_postprocess_boxes / fsg.detector_postprocess against the reference's stored outputs (bit-exact)."""
    import full_scale_gambler_for_object_detection_b200 as fsg

    g = gu.load("postprocess")
    for i in range(3):
        ih, iw, oh, ow = [int(v) for v in g["sizes_%d" % i]]
        inst = fsg.Instances((ih, iw))
        inst.pred_boxes = fsg.Boxes(g["in_boxes_%d" % i].to(cuda))
        inst.scores = g["in_scores_%d" % i].to(cuda)
        inst.pred_classes = g["in_classes_%d" % i].to(cuda)
        r = fsg.detector_postprocess(inst, oh, ow)
        assert tuple(r.image_size) == (oh, ow)
        assert torch.equal(r.pred_boxes.tensor.cpu(), g["boxes_%d" % i])
        assert torch.equal(r.scores.cpu(), g["scores_%d" % i])
        assert_equal_int(r.pred_classes, g["classes_%d" % i], "classes")


def test_gpu_grid_anchors_vs_reference(cuda):
    """fsg_grid_anchors (device-side DefaultAnchorGenerator) against the reference's stored anchors."""
    import full_scale_gambler_for_object_detection_b200 as fsg
    from full_scale_gambler_for_object_detection_b200 import anchor_generator as ag

    g = gu.load("anchors")
    for name, ratios in (("a3", [[1.0]]), ("a9", [[0.5, 1.0, 2.0]])):
        grids = [tuple(int(v) for v in r) for r in g["grids_" + name]]
        gen = fsg.DefaultAnchorGenerator([list(s) for s in ag.RETINANET_SIZES], ratios, ag.RETINANET_STRIDES, cuda)
        flat, offs = gen.flat_for_grids(grids)
        assert flat.shape[0] == int(g["count_" + name][0]) == offs[-1]
        assert torch.equal(flat[::97].cpu(), g["sample_" + name])
        assert torch.equal(flat.cpu().double().sum(dim=0), g["checksum_" + name])
        want = torch.cat(ag.grid_anchors(grids, ag.RETINANET_STRIDES, ag.RETINANET_SIZES,
                                         [r for r in ratios] * 5))
        assert torch.equal(flat.cpu(), want)
        feats = [torch.empty((2, 1, h, w), device=cuda) for h, w in grids]
        per_image = gen(feats)
        assert len(per_image) == 2 and len(per_image[0]) == 5 and len(per_image[0][0]) == offs[1]


def test_gpu_two_stage_callers_vs_reference(cuda):
    """find_top_rpn_proposals, RPN ground truth and fast_rcnn_inference_single_image on the GPU against the
    reference's stored outputs: bit-exact (selection, clipping, NMS keep order, labels); deltas within 1e-5."""
    import full_scale_gambler_for_object_detection_b200 as fsg
    from full_scale_gambler_for_object_detection_b200 import proposals as P, synthetic

    g = gu.load("two_stage")
    for ci in range(2):
        p = [int(v) for v in g["rpn%d_params" % ci]]
        cid, N, pre, post, counts = p[0], p[1], p[2], p[3], p[4:]
        thr, min_side = [float(v) for v in g["rpn%d_fparams" % ci]]
        inp = synthetic.rpn_inputs(cid, N, counts, ties=False)
        got = P.find_top_rpn_proposals([t.to(cuda) for t in inp["proposals"]], [t.to(cuda) for t in inp["logits"]],
                                       inp["image_sizes"], thr, pre, post, min_side, False)
        for n in range(N):
            assert torch.equal(got[n].proposal_boxes.tensor.cpu(), g["rpn%d_boxes_%d" % (ci, n)])
            assert torch.equal(got[n].objectness_logits.cpu(), g["rpn%d_logits_%d" % (ci, n)])
    inp = synthetic.train_inputs(63, 3, 256, 320, 80, M=6)
    gl, gd = P.rpn_ground_truth(inp["anchors"].to(cuda), [b.to(cuda) for b in inp["gt_boxes"]])
    for n in range(3):
        assert_equal_int(gl[n], g["rpngt_labels_%d" % n], "rpn labels")
        assert_close_tensor(gd[n][::13], g["rpngt_deltas_%d" % n], "rpn deltas")
    for ci in range(2):
        cid, R, K, spec = [int(v) for v in g["frcnn%d_params" % ci]]
        inp = synthetic.fast_rcnn_inputs(cid, R, K, bool(spec))
        r, rows = P.fast_rcnn_inference_single_image(inp["boxes"].to(cuda), inp["scores"].to(cuda),
                                                     inp["image_shape"], float(g["frcnn%d_thr" % ci][0]), 0.5, 100)
        assert torch.equal(r.pred_boxes.tensor.cpu(), g["frcnn%d_boxes" % ci])
        assert torch.equal(r.scores.cpu(), g["frcnn%d_scores" % ci])
        assert_equal_int(r.pred_classes, g["frcnn%d_classes" % ci], "classes")
        assert_equal_int(rows, g["frcnn%d_rows" % ci], "rows")
