"""CPU: the oracle against the reference's outputs stored in tests/golden (and, when the reference tree is
present -- the build container -- against the reference executed live)."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as orc
from oracle import ref_loader
from tests import golden_util as gu
from tests.util import assert_close_scalar, assert_close_tensor, assert_equal_int


def test_pairwise_iou_reference_kat():
    """tests/test_boxes.py:36-59 of the reference, replayed on the oracle."""
    b1 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    b2 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 1.0, 0.5],
                       [0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.5, 1.5]])
    expected = torch.tensor([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2)
    assert torch.allclose(orc.pairwise_iou(b1, b2), expected)


def test_anchor_generator_reference_kat():
    """tests/test_anchor_generator.py:14-43 (sizes [32,64], ratios [.25,1,4], stride 4, 1x2 grid)."""
    from full_scale_gambler_for_object_detection_b200.anchor_generator import grid_anchors

    expected = torch.tensor([
        [-32.0, -8.0, 32.0, 8.0], [-16.0, -16.0, 16.0, 16.0], [-8.0, -32.0, 8.0, 32.0],
        [-64.0, -16.0, 64.0, 16.0], [-32.0, -32.0, 32.0, 32.0], [-16.0, -64.0, 16.0, 64.0],
        [-28.0, -8.0, 36.0, 8.0], [-12.0, -16.0, 20.0, 16.0], [-4.0, -32.0, 12.0, 32.0],
        [-60.0, -16.0, 68.0, 16.0], [-28.0, -32.0, 36.0, 32.0], [-12.0, -64.0, 20.0, 64.0]])
    for fn in (grid_anchors, orc.grid_anchors):
        got = fn([(1, 2)], [4], [[32, 64]], [[0.25, 1, 4]])[0]
        assert torch.allclose(got, expected)


def test_box2box_roundtrip_reference_test():
    """tests/test_box2box_transform.py:16-30."""
    torch.manual_seed(0)
    w = (5, 5, 10, 10)
    src = torch.rand(10, 4) + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    dst = torch.rand(10, 4) + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    assert torch.allclose(dst, orc.apply_deltas(orc.get_deltas(src, dst, w), src, w))


def test_box2box_golden():
    g = gu.load("box2box")
    w = tuple(float(v) for v in g["weights"])
    assert torch.allclose(orc.get_deltas(g["src"], g["dst"], w), g["deltas"], rtol=1e-6, atol=0)
    assert torch.allclose(orc.apply_deltas(g["big"], g["boxes"], w), g["applied"], rtol=1e-6, atol=1e-6)


def test_matcher_golden():
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("matcher_stress")
    cid, N, R, M = [int(v) for v in g["params"]]
    inp = synthetic.matcher_stress_inputs(cid, N, R, M)
    inp["anchors"][0, 100] = inp["anchors"][0, 99]
    inp["anchors"][1, :50] = inp["gt_boxes"][1][:50]
    for i in range(N):
        q = orc.pairwise_iou(inp["gt_boxes"][i], inp["anchors"][i])
        assert torch.equal(q[:, :512], g["iou_sample_%d" % i])
        assert torch.equal(q.max(dim=1).values, g["iou_rowmax_%d" % i])
        m, l = orc.matcher(q, [0.4, 0.5], [0, -1, 1], True)
        assert_equal_int(m, g["matches_%d" % i], "matches")
        assert_equal_int(l, g["labels_%d" % i], "labels")
        _, pl = orc.matcher(q, [0.4, 0.9], [0, -1, 1], True)
        assert_equal_int(pl, g["picky_labels_%d" % i], "picky labels")
        m2, l2 = orc.matcher(q, [0.3, 0.7], [0, -1, 1], False)
        assert_equal_int(m2, g["nolq_matches_%d" % i], "no-lq matches")
        assert_equal_int(l2, g["nolq_labels_%d" % i], "no-lq labels")
    q = torch.tensor([[0.9, 0.3, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    m, l = orc.matcher(q, [0.4, 0.5], [0, -1, 1], True)
    assert_equal_int(m, g["quirk_matches"], "quirk matches")
    assert_equal_int(l, g["quirk_labels"], "quirk labels")
    m, l = orc.matcher(torch.zeros((0, 7)), [0.4, 0.5], [0, -1, 1], True)
    assert_equal_int(m, g["empty_matches"], "empty matches")
    assert_equal_int(l, g["empty_labels"], "empty labels")


def test_nms_golden():
    g = gu.load("nms")
    for name in ("a", "b", "c"):
        b, s, i = g["boxes_" + name], g["scores_" + name], g["idxs_" + name]
        for thr in (0.2, 0.5, 0.8):
            tag = "%s_%02d" % (name, int(thr * 10))
            assert_equal_int(orc.nms(b, s, thr), g["nms_" + tag], "nms " + tag)
            assert_equal_int(orc.batched_nms(b, s, i, thr), g["batched_" + tag], "batched " + tag)


def test_nms_matches_reference_python_nms():
    """The pure-Python greedy NMS the reference's tests use as their oracle (tests/test_nms_rotated.py:11-33):
    sort descending, keep the head, drop everything with IoU > thr."""
    g = torch.Generator().manual_seed(9)
    boxes = torch.rand((300, 4), generator=g) * 100
    boxes[:, 2:] += boxes[:, :2]
    scores = torch.rand(300, generator=g)
    for thr in (0.2, 0.5, 0.8):
        picked = []
        idx = scores.sort(descending=True, stable=True).indices
        while len(idx) > 0:
            cur = idx[0]
            picked.append(int(cur))
            if len(idx) == 1:
                break
            rest = idx[1:]
            iou = orc.pairwise_iou(boxes[rest], boxes[cur][None]).squeeze(1)
            idx = rest[iou <= thr]
        assert orc.nms(boxes, scores, thr).tolist() == picked


def test_inference_golden():
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("inference")
    p = [int(v) for v in g["params"]]
    inp = synthetic.inference_inputs(p[0], p[1], p[2:7], p[7])
    offs = inp["level_offsets"]
    for n in range(p[1]):
        cls = [inp["logits"][n, offs[i]:offs[i + 1]] for i in range(5)]
        reg = [inp["deltas"][n, offs[i]:offs[i + 1]] for i in range(5)]
        anc = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(5)]
        (b, s, c), _, _ = orc.inference_single_image(cls, reg, anc, p[7])
        assert_equal_int(c, g["classes_%d" % n], "classes")
        assert torch.equal(s, g["scores_%d" % n]) and torch.equal(b, g["boxes_%d" % n])


@pytest.mark.parametrize("name", gu.TRAIN_CASES)
def test_train_step_golden(name):
    inp, g, coeffs, detach, gcfg, K = gu.train_case(name)
    okw = {gu.ORACLE_KW[k]: v for k, v in gcfg.items()}
    got = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                         inp["bets"], K, *coeffs, detach_pred=detach, **okw)
    assert_equal_int(got["gt_classes"], g["gt_classes"], "gt_classes")
    assert_equal_int(got["mask"], g["mask"], "mask")
    for k in ("loss_cls", "loss_box_reg", "gambler_loss", "total", "loss_before_weighting", "lower_bound"):
        assert_close_scalar(got[k], g[k], k, rtol=2e-6)
    for k in ("gt_deltas", "per_anchor_loss", "weights", "grad_deltas"):
        assert_close_tensor(got[k], g[k], k, rtol=2e-6)
    # d/d bets = -(m/S)(l - A) cancels when l ~ A: the absolute floor is fp32 rounding of l and A themselves
    assert_close_tensor(got["grad_bets"], g["grad_bets"], "grad_bets", rtol=2e-6, atol_scale=1e-6)
    assert_close_tensor(got["grad_logits"].reshape(-1, K)[g["grad_rows"]], g["grad_logits"], "grad_logits", rtol=2e-6)


def test_gradient_formulas_match_autograd_fp64():
    """the closed forms the CUDA kernels use, against autograd in fp64."""
    torch.manual_seed(1)
    N, R, K, T = 2, 50, 7, 0.1
    x = torch.randn(N, R, K, dtype=torch.float64, requires_grad=True)
    b = torch.rand(N, R, dtype=torch.float64, requires_grad=True)
    gtc = torch.randint(-1, K + 1, (N, R))
    m = torch.randint(0, 2, (N, R))
    out = orc.gambler_loss(x, b, gtc, m, K, temperature=T)
    out["gambler_loss"].backward()
    t, _ = orc.one_hot_targets(gtc.flatten(), K, x.detach().reshape(-1, K))
    t = t.reshape(N, R, K)
    p = torch.sigmoid(x.detach())
    pt = p * t + (1 - p) * (1 - t)
    at = 0.25 * t + 0.75 * (1 - t)
    fprime = at * (2 * t - 1) * (1 - pt) ** 2 * (2 * pt * torch.log(pt) - (1 - pt))
    w = b.detach() * m + T
    S = w.sum(1, keepdim=True)
    w_hat = w / S
    valid = (gtc >= 0).double()[..., None]
    assert torch.allclose(x.grad, -w_hat[..., None] * fprime * valid, atol=1e-12)
    ell = out["per_anchor_loss"]
    A = (w_hat * ell).sum(1, keepdim=True)
    assert torch.allclose(b.grad, -(m / S) * (ell - A), atol=1e-12)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_oracle_vs_live_reference():
    """Build container only: run the reference's own files on a fresh seed and compare."""
    from oracle import make_golden as mg
    from full_scale_gambler_for_object_detection_b200 import synthetic

    ref = ref_loader.load_reference()
    inp = synthetic.train_inputs(77, 2, 192, 256, 80, M=4)
    want = mg.run_reference_train(ref, inp, 80, (192, 256))
    got = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                         inp["bets"], 80, 1.0, 1.0, -1.0)
    assert_equal_int(got["gt_classes"], want["gt_classes"], "gt_classes")
    assert_equal_int(got["mask"], want["mask"], "mask")
    for k in ("loss_cls", "loss_box_reg", "gambler_loss"):
        assert_close_scalar(got[k], want[k], k, rtol=2e-6)
    assert_close_tensor(got["grad_logits"], want["grad_logits"], "grad_logits", rtol=2e-6)


def test_detector_postprocess_golden():
    """oracle detector_postprocess == the reference's (postprocessing.py:8-52) stored outputs."""
    g = gu.load("postprocess")
    for i in range(3):
        ih, iw, oh, ow = [int(v) for v in g["sizes_%d" % i]]
        b, s, c = orc.detector_postprocess(g["in_boxes_%d" % i], g["in_scores_%d" % i], g["in_classes_%d" % i],
                                           (ih, iw), oh, ow)
        assert torch.equal(b, g["boxes_%d" % i]) and torch.equal(s, g["scores_%d" % i])
        assert_equal_int(c, g["classes_%d" % i], "classes")
        assert 0 < b.shape[0] < g["in_boxes_%d" % i].shape[0]   # some boxes really were dropped


def test_grid_anchors_golden():
    """oracle grid_anchors == the reference's DefaultAnchorGenerator (A = 3 gambler config and A = 9 upstream)."""
    from full_scale_gambler_for_object_detection_b200 import anchor_generator as ag

    g = gu.load("anchors")
    for name, ratios in (("a3", ((1.0,),) * 5), ("a9", ((0.5, 1.0, 2.0),) * 5)):
        grids = [tuple(int(v) for v in r) for r in g["grids_" + name]]
        for fn in (orc.grid_anchors, ag.grid_anchors):
            flat = torch.cat(fn(grids, ag.RETINANET_STRIDES, ag.RETINANET_SIZES, ratios))
            assert flat.shape[0] == int(g["count_" + name][0])
            assert torch.equal(flat[::97], g["sample_" + name])
            assert torch.equal(flat.double().sum(dim=0), g["checksum_" + name])


def test_two_stage_callers_golden():
    """oracle find_top_rpn_proposals / rpn_ground_truth / fast_rcnn_inference_single_image == the reference's
    stored outputs (rpn_outputs.py:52-151,250-295; fast_rcnn.py:76-118)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("two_stage")
    for ci in range(2):
        p = [int(v) for v in g["rpn%d_params" % ci]]
        cid, N, pre, post, counts = p[0], p[1], p[2], p[3], p[4:]
        thr, min_side = [float(v) for v in g["rpn%d_fparams" % ci]]
        inp = synthetic.rpn_inputs(cid, N, counts, ties=False)
        got = orc.find_top_rpn_proposals(inp["proposals"], inp["logits"], inp["image_sizes"], thr, pre, post, min_side)
        for n in range(N):
            assert torch.equal(got[n][0], g["rpn%d_boxes_%d" % (ci, n)])
            assert torch.equal(got[n][1], g["rpn%d_logits_%d" % (ci, n)])
    inp = synthetic.train_inputs(63, 3, 256, 320, 80, M=6)
    ol, od = orc.rpn_ground_truth(inp["anchors"], inp["gt_boxes"])
    for n in range(3):
        assert torch.equal(ol[n], g["rpngt_labels_%d" % n]) and torch.equal(od[n][::13], g["rpngt_deltas_%d" % n])
    for ci in range(2):
        cid, R, K, spec = [int(v) for v in g["frcnn%d_params" % ci]]
        inp = synthetic.fast_rcnn_inputs(cid, R, K, bool(spec))
        ob, os_, oc, orow = orc.fast_rcnn_inference_single_image(inp["boxes"], inp["scores"], inp["image_shape"],
                                                                 float(g["frcnn%d_thr" % ci][0]), 0.5, 100)
        assert torch.equal(ob, g["frcnn%d_boxes" % ci]) and torch.equal(os_, g["frcnn%d_scores" % ci])
        assert_equal_int(oc, g["frcnn%d_classes" % ci], "classes")
        assert_equal_int(orow, g["frcnn%d_rows" % ci], "rows")


def test_log_metrics_golden():
    """GANTrainer.calc_log_metrics (train_net.py:1089-1124): the oracle restatement against the values the
    reference's own function source produced (tests/golden/log_metrics.npz)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    g = gu.load("log_metrics")
    cid, N, H, W, K, M = [int(v) for v in g["params"]]
    lam_reg, kappa, lam_out = [float(v) for v in g["lambdas"]]
    inp = synthetic.train_inputs(cid, N, H, W, K, M=M)
    out = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                         inp["bets"], K, 1.0, 1.0, -1.0, need_grad=False)
    masked = orc.flat_to_nahw(inp["bets"] * out["mask"], inp["grids"], inp["A"])
    got = orc.calc_log_metrics(masked, out["weights"], out["loss_cls"], out["loss_box_reg"], out["gambler_loss"],
                               out["loss_before_weighting"], lam_reg, kappa, lam_out)
    for name, want in zip([str(n) for n in g["names"]], g["values"].tolist()):
        assert_close_scalar(float(got[name]), want, name)
