#!/usr/bin/env python
"""The config-2 fused step (DenseStepPlan, CUDA graph) and its stages, us per call -- kernel experiments.
Environment switches read by the library: FSG_MATCH_TWO_KERNELS=1, FSG_STEP_NO_PDL=1; FSG_STEP_STAGED=1 (python)."""
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
x, d, b = (inp[k].to(dev) for k in ("logits", "deltas", "bets"))
anchors = inp["anchors"].to(dev)
gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)


def graph_us(fn, calls=10, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (calls * reps) * 1e3


if "--once" in sys.argv:        # for ncu: a few plain launches
    for _ in range(3):
        plan.run(x, d, b, anchors, gt)
    torch.cuda.synchronize()
    sys.exit(0)
print("step   %.1f us   (one_call=%s)" % (graph_us(lambda: plan.run(x, d, b, anchors, gt)), plan.one_call))
print("match  %.1f us" % graph_us(lambda: plan.stage_match(b, anchors, gt)))
print("main   %.1f us" % graph_us(lambda: plan.stage_main(x, d, b, anchors, gt)))
print("post   %.1f us" % graph_us(lambda: plan.stage_post(b)))
print("num_fg %d  total %.6f" % (int(plan.stats[0]), float(plan.scalars[8])))
