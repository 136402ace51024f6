#!/usr/bin/env python
"""Run each hot-path workload a few times (for ncu): train step (config 2), LVIS loss (config 3 shard,
1 image slice x 8 -> N=8), detect (config 4) and matcher stress (config 5).

    python profiles/run_workloads.py [train] [lvis] [detect] [match] [--reps 3]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402

ALL = ["train", "train_native", "lvis", "lvis_native", "detect", "match", "rpn", "nms_large"]
which = [a for a in sys.argv[1:] if a in ALL] or ALL
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
warm = int(sys.argv[sys.argv.index("--warm") + 1]) if "--warm" in sys.argv else 2
dev = torch.device("cuda:0")


def timed(name, fn):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-8s %.3f ms/iter" % (name, a.elapsed_time(b) / reps), flush=True)


def train(K, N, cid):
    inp = synthetic.train_inputs(cid, N, 800, 1333, K)
    cfg = fsg.DenseLossConfig(num_classes=K)
    plan = fsg.DenseStepPlan(N, inp["R"], K, cfg, dev)
    x, d, b = inp["logits"].to(dev), inp["deltas"].to(dev), inp["bets"].to(dev)
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    return lambda: plan.run(x, d, b, anchors, gt)


if "train" in which:
    timed("train", train(80, 16, 2))
if "train_native" in which:
    # config 2 from the per-level head outputs (K1 with in-place betting maps, K2 and post on the native layout)
    K, N = 80, 16
    inp = synthetic.train_inputs(2, N, 800, 1333, K, logits=False)
    A, grids = inp["A"], inp["grids"]
    g = torch.Generator(device=dev).manual_seed(2)
    xs = [torch.randn((N, A * K, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT for h, w in grids]
    ds = [torch.randn((N, A * 4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
    bs = [torch.sigmoid(torch.randn((N, A, h, w), device=dev, generator=g) - 4.6) for h, w in grids]
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    planl = fsg.DenseStepPlanLevels(N, grids, A, K, fsg.DenseLossConfig(num_classes=K), dev)
    timed("train_native", lambda: planl.run(xs, ds, bs, anchors, gt))
    del xs, ds, bs, planl
if "lvis" in which:
    timed("lvis", train(1230, 8, 3))
if "lvis_native" in which:
    # config 3 shard on the head's native layout: K2 alone (fsg_loss_main_levels) and the whole step from head outputs
    K, N = 1230, 8
    inp = synthetic.train_inputs(3, N, 800, 1333, K, logits=False)
    A, grids, R = inp["A"], inp["grids"], inp["R"]
    g = torch.Generator(device=dev).manual_seed(3)
    xs = [torch.randn((N, A * K, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT for h, w in grids]
    ds = [torch.randn((N, A * 4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
    bs = [torch.sigmoid(torch.randn((N, A, h, w), device=dev, generator=g) - 4.6) for h, w in grids]
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    cfg = fsg.DenseLossConfig(num_classes=K)
    params = cfg.loss_params(1.0, 1.0, -1.0)
    bets = fsg.ops.anchor_maps_to_flat([bs])[0]
    m = fsg.ops.match_anchors(anchors, gt, K, bets=bets, temperature=0.1)
    gl = [torch.empty_like(x) for x in xs]
    timed("lvis_native_k2", lambda: fsg.ops.loss_main_levels(
        xs, m["gt_classes"], params, m["stats"], delta_levels=ds, anchors=anchors, gt=gt,
        matched_idx32=m["matched_idx32"], mask=m["mask"], bets=bets))
    del gl
if "detect" in which:
    inp = synthetic.inference_inputs(4, 1, [24000] * 5, 80)
    g = torch.Generator().manual_seed(4)
    N4 = 32
    logits = (torch.randn((N4, inp["R"], 80), generator=g) * 1.5 + synthetic.PRIOR_LOGIT).to(dev)
    deltas = (torch.randn((N4, inp["R"], 4), generator=g) * 0.2).to(dev)
    anchors = inp["anchors"].to(dev)
    thr = float(os.environ.get("FSG_DETECT_THR", "0.05"))   # e.g. 0.9999: nothing passes -> cost of the bare scan
    timed("detect", lambda: fsg.ops.detect(logits, deltas, anchors, inp["level_offsets"], score_threshold=thr))
if "rpn" in which:
    # find_top_rpn_proposals, FPN Faster R-CNN training setting (Base-RCNN-FPN.yaml): 16 images, P2..P6 of an
    # 800x1333 input with A = 3, pre_nms_topk 2000 per level, post_nms_topk 1000, NMS 0.7
    counts = [200 * 336 * 3, 100 * 168 * 3, 50 * 84 * 3, 25 * 42 * 3, 13 * 21 * 3]
    ri = synthetic.rpn_inputs(6, 16, counts, ties=False)
    P = [t.to(dev) for t in ri["proposals"]]
    Lg = [t.to(dev) for t in ri["logits"]]
    timed("rpn", lambda: fsg.ops.rpn_proposals(P, Lg, ri["image_sizes"], 0.7, 2000, 1000, 0.0))
if "nms_large" in which:
    g = torch.Generator().manual_seed(8)
    n = 20000
    xy = torch.rand((n, 2), generator=g) * 2000
    wh = torch.rand((n, 2), generator=g) * 90 + 10
    bx, sc = torch.cat([xy, xy + wh], 1).to(dev), torch.rand(n, generator=g).to(dev)
    timed("nms20k", lambda: fsg.ops.nms_raw(bx, sc, None, 0.5))
if "match" in which:
    inp5 = synthetic.matcher_stress_inputs(5, 8, 1000000, 200)
    a5 = inp5["anchors"].to(dev)
    gt5 = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"], inp5["gt_classes"], dev)
    timed("match", lambda: fsg.ops.match_anchors(a5, gt5, 80, want=("matches", "match_labels"),
                                                 picky_thresholds=None))
