#!/usr/bin/env python
"""K1 (two launches) alone on config 2, graph of 10 calls: us per call (kernel experiments)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402

dev = torch.device("cuda")
N, K = 16, 80
inp = synthetic.train_inputs(2, N, 800, 1333, K)
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
print("K1: %.1f us per call   num_fg %d" % (e0.elapsed_time(e1) / 100 * 1e3, int(plan.stats[0])))
