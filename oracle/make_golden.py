"""Generate tests/golden/*.npz by running the REFERENCE's own source files (container only).

    python -m oracle.make_golden          # writes tests/golden/, prints oracle-vs-reference deltas

Each fixture stores seeded inputs *generation parameters* (regenerated deterministically by
``full_scale_gambler_for_object_detection_b200.synthetic``) or small explicit inputs, together with the outputs the
reference produced for them through ``oracle/ref_loader.py``.  ``tests/test_oracle_golden.py`` checks the
oracle against these files on any machine; the GPU parity tests check the CUDA path against the same files.
TEST INFRASTRUCTURE -- never imported by the product.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dense_oracle as orc  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (list, tuple)) and v and isinstance(v[0], torch.Tensor):
            for i, t in enumerate(v):
                out["%s_%d" % (k, i)] = t.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def ref_targets(ref, inp, hw):
    targets = []
    for b, c in zip(inp["gt_boxes"], inp["gt_classes"]):
        t = ref.Instances(hw)
        t.gt_boxes = ref.Boxes(b)
        t.gt_classes = c
        targets.append(t)
    return targets


def ref_anchor_lists(ref, inp):
    offs = inp["level_offsets"]
    per_level = [ref.Boxes(inp["anchors"][offs[i]:offs[i + 1]].clone()) for i in range(len(offs) - 1)]
    return [per_level for _ in range(inp["N"])]


def flat_to_levels_nchw(flat, grids, A):
    """(N,R,K) -> list[(N, A*K, H, W)], inverse of retinanet.py:24-33."""
    N, R, K = flat.shape
    out, off = [], 0
    for H, W in grids:
        n = H * W * A
        t = flat[:, off:off + n].reshape(N, H, W, A, K).permute(0, 3, 4, 1, 2).reshape(N, A * K, H, W)
        out.append(t.contiguous())
        off += n
    return out


def run_reference_train(ref, inp, K, hw, coeffs=(1.0, 1.0, -1.0), detach_pred=False, **gcfg):
    """The reference's own train-step path on flattened synthetic inputs: get_ground_truth,
    get_picky_ground_truth, losses, gambler_loss, backward of the combination."""
    cfg = rl.make_cfg(NUM_CLASSES=K, IN_LAYERS=[g[0] for g in inp["grids"]], **gcfg)
    cfg.MODEL.RETINANET.NUM_CLASSES = K
    rl.set_global_cfg(ref, cfg)
    me = rl.retinanet_self(ref, num_classes=K)
    anchors = ref_anchor_lists(ref, inp)
    targets = ref_targets(ref, inp, hw)
    gt_classes, gt_deltas = ref.RetinaNet.get_ground_truth(me, anchors, targets)
    mask = ref.RetinaNet.get_picky_ground_truth(me, anchors, targets)
    A = inp["A"]
    cls_levels = [t.requires_grad_(True) for t in flat_to_levels_nchw(inp["logits"], inp["grids"], A)]
    reg_levels = [t.requires_grad_(True) for t in flat_to_levels_nchw(inp["deltas"], inp["grids"], A)]
    bet_levels = [t.reshape(t.shape[0], A, t.shape[2], t.shape[3]).requires_grad_(True)
                  for t in flat_to_levels_nchw(inp["bets"][..., None], inp["grids"], A)]
    losses = ref.RetinaNet.losses(me, gt_classes, gt_deltas, cls_levels, reg_levels)
    g = rl.gambler_self(ref, cfg)
    bets_in = list(bet_levels)
    loss_dict, weights = g.gambler_loss(cls_levels, bets_in, gt_classes, mask, detach_pred=detach_pred)
    total = coeffs[0] * losses["loss_cls"] + coeffs[1] * losses["loss_box_reg"] + coeffs[2] * loss_dict["gambler_loss"]
    total.backward()

    def flat_grad(levels, Kc):
        gs = [t.grad if t.grad is not None else torch.zeros_like(t) for t in levels]
        return orc.levels_to_flat(gs, Kc)

    N = inp["N"]
    ell = torch.cat([l.permute(0, 2, 3, 1).reshape(N, -1) for l in loss_dict["NAKHW_loss"]], dim=1)
    return dict(
        gt_classes=gt_classes, gt_deltas=gt_deltas, mask=mask,
        loss_cls=losses["loss_cls"].detach(), loss_box_reg=losses["loss_box_reg"].detach(),
        gambler_loss=loss_dict["gambler_loss"].detach(), total=total.detach(),
        loss_before_weighting=loss_dict["loss_before_weighting"].detach(),
        lower_bound=torch.as_tensor(ref.storage.scalars["loss_gambler/lower_bound"]).detach(),
        per_anchor_loss=ell.detach(), weights=weights.reshape(N, -1),
        grad_logits=flat_grad(cls_levels, K), grad_deltas=flat_grad(reg_levels, 4),
        grad_bets=flat_grad(bet_levels, 1).reshape(N, -1),
    )


def sample_rows(gt_classes, K, stride=41):
    """Flattened (n*R + r) anchor rows whose grad_logits go into the fixture."""
    g = gt_classes.flatten()
    special = (g != K).nonzero().flatten()           # foreground and ignored anchors
    strided = torch.arange(0, g.numel(), stride)
    return torch.unique(torch.cat((special, strided)))


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    d = (a - b).abs().max().item() if a.numel() else 0.0
    s = b.abs().max().item() if b.numel() else 0.0
    return d / s if s > 0 else d


TRAIN_CASES = {
    # name: (config_id, N, H, W, K, M, coeffs, detach_pred, gambler cfg overrides)
    "train_config1": (1, 2, 512, 512, 80, 8, (1.0, 1.0, -1.0), False, {}),
    "train_gphase": (2, 2, 256, 320, 80, 8, (0.0, 0.0, 1.0), True, {}),
    "train_extendtobatch": (3, 3, 256, 256, 80, 5, (1.0, 0.5, -2.0), False, {"GAMBLER_OUTPUT": "L_BAHW_extendtobatch"}),
    "train_sigmoid": (3, 3, 256, 256, 80, 5, (1.0, 0.5, -2.0), False, {"GAMBLER_LOSS_MODE": "sigmoid"}),
    "train_nonorm": (3, 3, 256, 256, 80, 5, (1.0, 0.5, -2.0), False, {"NORMALIZE": False}),
    "train_lvis_k1230": (4, 2, 192, 192, 1230, 6, (1.0, 1.0, -1.0), False, {}),
}

ORACLE_KW = {"GAMBLER_OUTPUT": "output", "GAMBLER_LOSS_MODE": "mode", "NORMALIZE": "normalize"}


def make_train(ref):
    for name, (cid, N, H, W, K, M, coeffs, detach, gcfg) in TRAIN_CASES.items():
        inp = synthetic.train_inputs(cid, N, H, W, K, M=M)
        want = run_reference_train(ref, inp, K, (H, W), coeffs, detach, **gcfg)
        okw = {ORACLE_KW[k]: v for k, v in gcfg.items()}
        got = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                             inp["bets"], K, *coeffs, detach_pred=detach, **okw)
        print("[%s] oracle vs reference:" % name)
        for k in ("gt_classes", "mask"):
            assert torch.equal(got[k], want[k]), k
        for k in ("gt_deltas", "loss_cls", "loss_box_reg", "gambler_loss", "total", "loss_before_weighting",
                  "lower_bound", "per_anchor_loss", "weights", "grad_logits", "grad_deltas", "grad_bets"):
            print("    %-22s rel-to-max err %.3g" % (k, rel(got[k], want[k])))
        keep = dict(want)
        # keep fixtures small: grad_logits only for every foreground/ignored anchor plus a strided sample
        rows = sample_rows(want["gt_classes"], K)
        keep["grad_rows"] = rows
        keep["grad_logits"] = want["grad_logits"].reshape(-1, K)[rows].contiguous()
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            params=np.asarray([cid, N, H, W, K, M], dtype=np.int64),
                            coeffs=np.asarray(coeffs, dtype=np.float64), detach=np.asarray(int(detach)),
                            gcfg=np.asarray(repr(gcfg)), **_np(keep))


def make_matcher(ref):
    """Matcher + pairwise_iou through the reference on the stress shape and on hand-made edge cases."""
    inp = synthetic.matcher_stress_inputs(5, 2, 20000, 200)
    inp["anchors"][0, 100] = inp["anchors"][0, 99]
    inp["anchors"][1, :50] = inp["gt_boxes"][1][:50]
    res = {}
    for i in range(2):
        q = ref.pairwise_iou(ref.Boxes(inp["gt_boxes"][i]), ref.Boxes(inp["anchors"][i]))
        m = ref.Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True)
        pm = ref.Matcher([0.4, 0.9], [0, -1, 1], allow_low_quality_matches=True)
        nm = ref.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=False)
        res["matches_%d" % i], res["labels_%d" % i] = m(q)
        _, res["picky_labels_%d" % i] = pm(q)
        res["nolq_matches_%d" % i], res["nolq_labels_%d" % i] = nm(q)
        res["iou_rowmax_%d" % i] = q.max(dim=1).values
        res["iou_sample_%d" % i] = q[:, :512].contiguous()
        oq = orc.pairwise_iou(inp["gt_boxes"][i], inp["anchors"][i])
        assert torch.equal(oq, q), "oracle IoU differs from the reference"
        om, ol = orc.matcher(oq, [0.4, 0.5], [0, -1, 1], True)
        assert torch.equal(om, res["matches_%d" % i]) and torch.equal(ol, res["labels_%d" % i])
    # quirks: zero-overlap GT, empty GT, tie rows
    q = torch.tensor([[0.9, 0.3, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]])
    res["quirk_matches"], res["quirk_labels"] = ref.Matcher([0.4, 0.5], [0, -1, 1], True)(q)
    e = torch.zeros((0, 7))
    res["empty_matches"], res["empty_labels"] = ref.Matcher([0.4, 0.5], [0, -1, 1], True)(e)
    np.savez_compressed(os.path.join(OUT, "matcher_stress.npz"), params=np.asarray([5, 2, 20000, 200]), **_np(res))
    print("[matcher_stress] oracle == reference (bit-exact)")


def make_box2box(ref):
    torch.manual_seed(3)
    w = (5, 5, 10, 10)
    src = torch.rand(10, 4) + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    dst = torch.rand(10, 4) + torch.tensor([10, 10, 20, 20], dtype=torch.float)
    t = ref.Box2BoxTransform(weights=w)
    deltas = t.get_deltas(src, dst)
    big = torch.randn(64, 12) * 3
    boxes = torch.rand(64, 4) * 50
    boxes[:, 2:] += boxes[:, :2] + 1
    applied = t.apply_deltas(big, boxes)
    assert torch.allclose(orc.get_deltas(src, dst, w), deltas, rtol=1e-6, atol=0)
    assert torch.allclose(orc.apply_deltas(big, boxes, w), applied, rtol=1e-6, atol=1e-6)
    np.savez_compressed(os.path.join(OUT, "box2box.npz"), src=src.numpy(), dst=dst.numpy(), deltas=deltas.numpy(),
                        big=big.numpy(), boxes=boxes.numpy(), applied=applied.numpy(), weights=np.asarray(w))
    print("[box2box] ok")


def make_nms(ref):
    res = {}
    for name, n, ncls, seed in (("a", 2000, 50, 2250), ("b", 5000, 80, 5280), ("c", 600, 3, 803)):
        g = torch.Generator().manual_seed(seed)
        boxes = torch.rand((n, 4), generator=g) * 100
        boxes[:, 2:] += boxes[:, :2]
        scores = torch.rand(n, generator=g)
        scores[5] = scores[3]
        boxes[7] = boxes[6]
        boxes[9, 2:] = boxes[9, :2]
        idxs = torch.randint(0, ncls, (n,), generator=g)
        res["boxes_" + name], res["scores_" + name], res["idxs_" + name] = boxes, scores, idxs
        for thr in (0.2, 0.5, 0.8):
            tag = "%s_%02d" % (name, int(thr * 10))
            # the reference's own entry points (detectron2/layers/nms.py) -> installed torchvision 0.26
            res["nms_" + tag] = ref.nms(boxes, scores, thr)
            res["batched_" + tag] = ref.batched_nms(boxes, scores, idxs, thr)
            assert torch.equal(orc.nms(boxes, scores, thr), res["nms_" + tag]), "oracle nms != torchvision"
            ob = orc.batched_nms(boxes, scores, idxs, thr)
            same = torch.equal(ob, res["batched_" + tag])
            print("[nms %s thr %.1f] oracle batched_nms == reference: %s (n=%d keep=%d)" % (name, thr, same, n, ob.numel()))
            assert same
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **_np(res))


def make_inference(ref):
    K = 80
    inp = synthetic.inference_inputs(41, 2, [3000, 800, 200, 60, 20], K)
    me = rl.retinanet_self(ref, num_classes=K)
    offs = inp["level_offsets"]
    res = {}
    for n in range(2):
        cls = [inp["logits"][n, offs[i]:offs[i + 1]].clone() for i in range(5)]
        reg = [inp["deltas"][n, offs[i]:offs[i + 1]] for i in range(5)]
        anc = [ref.Boxes(inp["anchors"][offs[i]:offs[i + 1]]) for i in range(5)]
        r = ref.RetinaNet.inference_single_image(me, cls, reg, anc, (800, 1344))
        res["boxes_%d" % n], res["scores_%d" % n], res["classes_%d" % n] = r.pred_boxes.tensor, r.scores, r.pred_classes
        cls2 = [inp["logits"][n, offs[i]:offs[i + 1]] for i in range(5)]
        anc2 = [inp["anchors"][offs[i]:offs[i + 1]] for i in range(5)]
        (ob, os_, oc), _, _ = orc.inference_single_image(cls2, reg, anc2, K)
        assert torch.equal(oc, r.pred_classes) and torch.equal(os_, r.scores) and torch.equal(ob, r.pred_boxes.tensor), \
            "oracle inference != reference"
    np.savez_compressed(os.path.join(OUT, "inference.npz"), params=np.asarray([41, 2, 3000, 800, 200, 60, 20, K]), **_np(res))
    print("[inference] oracle == reference (bit-exact on CPU)")


def make_postprocess(ref):
    """detector_postprocess on the reference's own Instances: down- and up-scaling, boxes leaving the image."""
    g = torch.Generator().manual_seed(77)
    res = {}
    for idx, (img, out) in enumerate([((800, 1344), (480, 640)), ((512, 512), (1024, 768)), ((800, 1216), (427, 640))]):
        n = 300
        xy = torch.rand((n, 2), generator=g) * torch.tensor([img[1] * 1.2, img[0] * 1.2]) - 60.0
        wh = torch.rand((n, 2), generator=g) * 200
        wh[::17] = 0.0                       # zero-area boxes
        boxes = torch.cat([xy, xy + wh], dim=1)
        boxes[5] = torch.tensor([img[1] + 3.0, 10.0, img[1] + 50.0, 60.0])   # entirely right of the image
        boxes[6] = torch.tensor([-80.0, -40.0, -1.0, -2.0])                  # entirely outside (top-left)
        scores = torch.rand(n, generator=g)
        classes = torch.randint(0, 80, (n,), generator=g)
        inst = ref.Instances(img)
        inst.pred_boxes = ref.Boxes(boxes.clone())
        inst.scores = scores.clone()
        inst.pred_classes = classes.clone()
        r = ref.detector_postprocess(inst, out[0], out[1])
        ob, os_, oc = orc.detector_postprocess(boxes, scores, classes, img, out[0], out[1])
        assert torch.equal(ob, r.pred_boxes.tensor) and torch.equal(os_, r.scores) and torch.equal(oc, r.pred_classes), \
            "oracle detector_postprocess != reference"
        assert tuple(r.image_size) == out
        res.update({"in_boxes_%d" % idx: boxes, "in_scores_%d" % idx: scores, "in_classes_%d" % idx: classes,
                    "sizes_%d" % idx: torch.tensor([img[0], img[1], out[0], out[1]]),
                    "boxes_%d" % idx: r.pred_boxes.tensor, "scores_%d" % idx: r.scores, "classes_%d" % idx: r.pred_classes})
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **_np(res))
    print("[postprocess] oracle == reference (bit-exact on CPU)")


def make_anchors(ref):
    """DefaultAnchorGenerator of the reference on the RetinaNet+gambler config (A = 3) and on upstream RetinaNet
    (3 scales x 3 ratios, A = 9): oracle grid_anchors must be bit-identical; a strided sample is stored."""
    import types as _t
    from full_scale_gambler_for_object_detection_b200 import anchor_generator as ag
    res = {}
    cases = {"a3": (ag.RETINANET_SIZES, ((1.0,),) * 5),
             "a9": (ag.RETINANET_SIZES, ((0.5, 1.0, 2.0),) * 5)}
    for name, (sizes, ratios) in cases.items():
        cfg = rl._AttrDict(MODEL=rl._AttrDict(ANCHOR_GENERATOR=rl._AttrDict(
            SIZES=[list(s) for s in sizes], ASPECT_RATIOS=[list(r) for r in ratios])))
        shapes = [_t.SimpleNamespace(stride=s) for s in ag.RETINANET_STRIDES]
        gen = ref.anchor_generator.DefaultAnchorGenerator(cfg, shapes)
        grids = ag.retinanet_grid_sizes(800, 1333)
        want = gen.grid_anchors(grids)
        got = orc.grid_anchors(grids, ag.RETINANET_STRIDES, sizes, ratios)
        for w, g_ in zip(want, got):
            assert torch.equal(w, g_), "oracle grid_anchors != reference"
        flat = torch.cat(want)
        res["grids_" + name] = torch.tensor(grids)
        res["count_" + name] = torch.tensor([flat.shape[0]])
        res["sample_" + name] = flat[::97].clone()
        res["checksum_" + name] = flat.double().sum(dim=0)
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), **_np(res))
    print("[anchors] oracle == reference (bit-exact on CPU)")


def make_rpn(ref):
    """find_top_rpn_proposals / RPN ground truth / fast_rcnn_inference_single_image of the reference (its own
    rpn_outputs.py, matcher.py, box_regression.py, fast_rcnn.py) against the oracle restatements."""
    import types as _t
    res = {}
    cases = [(61, 2, [6000, 2500, 700, 200, 60], 1000, 1000, 0.7, 0.0),
             (62, 1, [9000, 3000, 2000, 819], 2000, 1000, 0.7, 4.0)]
    for ci, (cid, N, counts, pre, post, thr, min_side) in enumerate(cases):
        # no exact logit ties here: the reference's sort is not stable (rpn_outputs.py:104 does not ask for it), so
        # the order of tied proposals is implementation-defined there; ours is "lower index first" (tested
        # against the oracle separately)
        inp = synthetic.rpn_inputs(cid, N, counts, ties=False)
        images = _t.SimpleNamespace(image_sizes=inp["image_sizes"])
        want = ref.rpn_outputs.find_top_rpn_proposals([p.clone() for p in inp["proposals"]],
                                                      [l.clone() for l in inp["logits"]], images, thr, pre, post,
                                                      min_side, False)
        got = orc.find_top_rpn_proposals(inp["proposals"], inp["logits"], inp["image_sizes"], thr, pre, post, min_side)
        for n in range(N):
            assert torch.equal(got[n][0], want[n].proposal_boxes.tensor), "oracle rpn proposals != reference"
            assert torch.equal(got[n][1], want[n].objectness_logits)
            res["rpn%d_boxes_%d" % (ci, n)] = want[n].proposal_boxes.tensor
            res["rpn%d_logits_%d" % (ci, n)] = want[n].objectness_logits
        res["rpn%d_params" % ci] = torch.tensor([cid, N, pre, post] + counts)
        res["rpn%d_fparams" % ci] = torch.tensor([thr, min_side])
    # RPN ground truth: Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True) + get_deltas
    inp = synthetic.train_inputs(63, 3, 256, 320, 80, M=6)
    me = _t.SimpleNamespace(
        anchors=[[ref.Boxes(inp["anchors"])] for _ in range(3)], gt_boxes=[ref.Boxes(b) for b in inp["gt_boxes"]],
        image_sizes=[(256, 320)] * 3, boundary_threshold=-1,
        anchor_matcher=ref.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True),
        box2box_transform=ref.Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)))
    wl, wd = ref.rpn_outputs.RPNOutputs._get_ground_truth(me)
    ol, od = orc.rpn_ground_truth(inp["anchors"], inp["gt_boxes"])
    for n in range(3):
        assert torch.equal(wl[n], ol[n]) and torch.equal(wd[n], od[n]), "oracle rpn ground truth != reference"
        res["rpngt_labels_%d" % n] = wl[n]
        res["rpngt_deltas_%d" % n] = wd[n][::13].clone()
    # fast_rcnn_inference_single_image: class-specific and class-agnostic boxes
    for ci, (cid, R, K, spec, sthr) in enumerate([(64, 1000, 80, True, 0.05), (65, 700, 20, False, 0.02)]):
        inp = synthetic.fast_rcnn_inputs(cid, R, K, spec)
        r, rows = ref.fast_rcnn.fast_rcnn_inference_single_image(inp["boxes"].clone(), inp["scores"].clone(),
                                                                 inp["image_shape"], sthr, 0.5, 100)
        ob, os_, oc, orow = orc.fast_rcnn_inference_single_image(inp["boxes"], inp["scores"], inp["image_shape"],
                                                                 sthr, 0.5, 100)
        assert torch.equal(ob, r.pred_boxes.tensor) and torch.equal(os_, r.scores), "oracle fast_rcnn != reference"
        assert torch.equal(oc, r.pred_classes) and torch.equal(orow, rows)
        res.update({"frcnn%d_boxes" % ci: ob, "frcnn%d_scores" % ci: os_, "frcnn%d_classes" % ci: oc,
                    "frcnn%d_rows" % ci: orow, "frcnn%d_params" % ci: torch.tensor([cid, R, K, int(spec)]),
                    "frcnn%d_thr" % ci: torch.tensor([sthr])})
    np.savez_compressed(os.path.join(OUT, "two_stage.npz"), **_np(res))
    print("[two-stage callers] oracle == reference (bit-exact on CPU)")


def reference_calc_log_metrics():
    """GANTrainer.calc_log_metrics (ImbalanceDetection/train_net.py:1089-1124) as the reference wrote it: the module
    imports half of detectron2 (engine, data, evaluation), so the method's own source is cut out of the file with
    `ast` and compiled stand-alone -- it only needs `torch` and a `self` with five attributes."""
    import ast
    import types

    path = os.path.join(rl.REF_ROOT, "ImbalanceDetection", "train_net.py")
    src = open(path).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "GANTrainer":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "calc_log_metrics":
                    fn = item
    assert fn is not None, "calc_log_metrics not found in %s" % path
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch}
    exec(compile(mod, path, "exec"), ns)
    return ns["calc_log_metrics"], types.SimpleNamespace


def make_log_metrics(ref):
    """calc_log_metrics on the outputs of a reference training step (config-1 shape and a GT-free image)."""
    calc, NS = reference_calc_log_metrics()
    cid, N, H, W, K, M = 6, 3, 256, 320, 80, 5
    inp = synthetic.train_inputs(cid, N, H, W, K, M=M)
    cfg = rl.make_cfg(NUM_CLASSES=K, IN_LAYERS=[g[0] for g in inp["grids"]])
    cfg.MODEL.RETINANET.NUM_CLASSES = K
    rl.set_global_cfg(ref, cfg)
    me = rl.retinanet_self(ref, num_classes=K)
    anchors, targets = ref_anchor_lists(ref, inp), ref_targets(ref, inp, (H, W))
    gt_classes, gt_deltas = ref.RetinaNet.get_ground_truth(me, anchors, targets)
    mask = ref.RetinaNet.get_picky_ground_truth(me, anchors, targets)
    A = inp["A"]
    cls_levels = flat_to_levels_nchw(inp["logits"], inp["grids"], A)
    reg_levels = flat_to_levels_nchw(inp["deltas"], inp["grids"], A)
    bet_levels = [t.reshape(t.shape[0], A, t.shape[2], t.shape[3])
                  for t in flat_to_levels_nchw(inp["bets"][..., None], inp["grids"], A)]
    losses = ref.RetinaNet.losses(me, gt_classes, gt_deltas, cls_levels, reg_levels)
    g = rl.gambler_self(ref, cfg)
    bets_in = list(bet_levels)
    gdict, weights = g.gambler_loss(cls_levels, bets_in, gt_classes, mask, detach_pred=False)
    lam_reg, kappa, lam_out = 0.7, 1.3, 0.9
    me2 = NS(cfg=_AttrNS(MODEL=_AttrNS(GAMBLER_HEAD=_AttrNS(DETECTOR_LOSS_MODE="cls+reg-gambler"))),
             regression_loss_lambda=lam_reg, gambler_loss_kappa=kappa, gambler_outside_lambda=lam_out,
             _detect_anomaly=lambda *a, **k: None)
    out = calc(me2, bets_in, weights, dict(losses), gdict, 0.0)      # bets_in is the MASKED list now (:568-569)
    names = ["loss_cls", "loss_box_reg", "loss_gambler", "loss_before_weighting", "loss_detector",
             "gambler_bets/sum", "gambler_bets/max", "gambler_bets/mean", "visualized weights/sum",
             "visualized weights/max", "visualized weights/mean", "visualized weights/median"]
    want = np.asarray([float(out[k]) for k in names], dtype=np.float64)
    got = orc.calc_log_metrics(bets_in, weights, losses["loss_cls"], losses["loss_box_reg"], gdict["gambler_loss"],
                               gdict["loss_before_weighting"], lam_reg, kappa, lam_out)
    for k, w in zip(names, want):
        assert abs(float(got[k]) - w) <= 1e-6 * max(abs(w), 1e-30), (k, float(got[k]), w)
    np.savez_compressed(os.path.join(OUT, "log_metrics.npz"), params=np.asarray([cid, N, H, W, K, M]),
                        lambdas=np.asarray([lam_reg, kappa, lam_out]), names=np.asarray(names), values=want)
    print("[log_metrics] oracle == reference (calc_log_metrics source executed): %d scalars" % len(names))


def reference_callables(ref):
    """The reference callables whose signatures the drop-ins must keep (tests/test_signatures.py)."""
    import sys as _sys
    smp = _sys.modules["detectron2.modeling.sampling"]
    return {
        "structures.Boxes.__init__": ref.Boxes.__init__,
        "structures.Instances.__init__": ref.Instances.__init__,
        "structures.pairwise_iou": ref.pairwise_iou,
        "modeling.matcher.Matcher.__init__": ref.Matcher.__init__,
        "modeling.matcher.Matcher.__call__": ref.Matcher.__call__,
        "modeling.box_regression.Box2BoxTransform.__init__": ref.Box2BoxTransform.__init__,
        "modeling.box_regression.Box2BoxTransform.get_deltas": ref.Box2BoxTransform.get_deltas,
        "modeling.box_regression.Box2BoxTransform.apply_deltas": ref.Box2BoxTransform.apply_deltas,
        "layers.nms.batched_nms": ref.batched_nms,
        "layers.nms.nms": ref.nms,                      # = torchvision.ops.nms (layers/nms.py:6)
        "meta_arch.retinanet.RetinaNet.losses": ref.RetinaNet.losses,
        "meta_arch.retinanet.RetinaNet.get_ground_truth": ref.RetinaNet.get_ground_truth,
        "meta_arch.retinanet.RetinaNet.get_picky_ground_truth": ref.RetinaNet.get_picky_ground_truth,
        "meta_arch.retinanet.RetinaNet.inference": ref.RetinaNet.inference,
        "meta_arch.retinanet.RetinaNet.inference_single_image": ref.RetinaNet.inference_single_image,
        "gambler_heads.LayeredUnetGambler.gambler_loss": ref.LayeredUnetGambler.gambler_loss,
        "gambler_heads.get_loss_upper_bound": ref.gambler_heads.get_loss_upper_bound,
        "proposal_generator.rpn_outputs.find_top_rpn_proposals": ref.rpn_outputs.find_top_rpn_proposals,
        "modeling.sampling.subsample_labels": smp.subsample_labels,
        "modeling.postprocessing.detector_postprocess": ref.detector_postprocess,
        "roi_heads.fast_rcnn.fast_rcnn_inference": ref.fast_rcnn.fast_rcnn_inference,
        "roi_heads.fast_rcnn.fast_rcnn_inference_single_image": ref.fast_rcnn.fast_rcnn_inference_single_image,
        "modeling.anchor_generator.DefaultAnchorGenerator.grid_anchors":
            ref.anchor_generator.DefaultAnchorGenerator.grid_anchors,
    }


def make_signatures(ref):
    import inspect
    import json

    out = {}
    for name, fn in reference_callables(ref).items():
        out[name] = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default), p.kind.name]
                     for p in inspect.signature(fn).parameters.values()]
    with open(os.path.join(OUT, "signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("[signatures] %d reference callables pinned" % len(out))


class _AttrNS(dict):
    def __getattr__(self, k):
        return self[k]


def main():
    assert rl.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    ref = rl.load_reference()
    make_postprocess(ref)
    make_anchors(ref)
    make_rpn(ref)
    if "--only-new" in sys.argv:
        return
    make_matcher(ref)
    make_box2box(ref)
    make_nms(ref)
    make_inference(ref)
    make_train(ref)
    make_log_metrics(ref)
    make_signatures(ref)


if __name__ == "__main__":
    main()
