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

ALL = ["train", "lvis", "detect", "match"]
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
if "lvis" in which:
    timed("lvis", train(1230, 8, 3))
if "detect" in which:
    inp = synthetic.inference_inputs(4, 1, [24000] * 5, 80)
    g = torch.Generator().manual_seed(4)
    N4 = 32
    logits = (torch.randn((N4, inp["R"], 80), generator=g) * 1.5 + synthetic.PRIOR_LOGIT).to(dev)
    deltas = (torch.randn((N4, inp["R"], 4), generator=g) * 0.2).to(dev)
    anchors = inp["anchors"].to(dev)
    timed("detect", lambda: fsg.ops.detect(logits, deltas, anchors, inp["level_offsets"]))
if "match" in which:
    inp5 = synthetic.matcher_stress_inputs(5, 8, 1000000, 200)
    a5 = inp5["anchors"].to(dev)
    gt5 = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"], inp5["gt_classes"], dev)
    timed("match", lambda: fsg.ops.match_anchors(a5, gt5, 80, want=("matches", "match_labels"),
                                                 picky_thresholds=None))
