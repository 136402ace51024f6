#!/usr/bin/env python
"""K1 on config 2 with ONE crowded image in the batch (60 GT instead of 8; COCO has such images): us per call.
The batch still averages <= 32 GT per image, so the few-GT pass A runs and the crowded image takes its staged-GT
branch (kernel experiments: warp-level cull in that branch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402

dev = torch.device("cuda")
N, K = 16, 80
for m0 in (8, 60, 90):
    inp = synthetic.train_inputs(2, N, 800, 1333, K)
    big = synthetic.train_inputs(3, 1, 800, 1333, K, M=m0, empty_image=False, logits=False)
    inp["gt_boxes"][0], inp["gt_classes"][0] = big["gt_boxes"][0], big["gt_classes"][0]
    cfg = fsg.DenseLossConfig(num_classes=K)
    plan = fsg.DenseStepPlan(N, inp["R"], K, cfg, dev)
    b = inp["bets"].to(dev)
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            plan.stage_match(b, anchors, gt)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            plan.stage_match(b, anchors, gt)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("K1 with image 0 at %2d GT: %.1f us per call   num_fg %d" % (m0, e0.elapsed_time(e1) / 100 * 1e3,
                                                                    int(plan.stats[0])))
