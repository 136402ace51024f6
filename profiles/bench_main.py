#!/usr/bin/env python
"""K2 main pass alone on config 2 (graph of 10 launches), for kernel experiments: prints us per launch."""
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
x, d, b = inp["logits"].to(dev), inp["deltas"].to(dev), inp["bets"].to(dev)
anchors = inp["anchors"].to(dev)
gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
plan.run(x, d, b, anchors, gt)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        plan.stage_main(x, d, b, anchors, gt)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        plan.stage_main(x, d, b, anchors, gt)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print("K2 main: %.1f us  %.0f GB/s (%.1f%% of 6461)  loss %.6f" % (
    us, 712 * N * inp["R"] / us / 1e3, 712 * N * inp["R"] / us / 1e3 / 64.612, float(plan.scalars[8])))
