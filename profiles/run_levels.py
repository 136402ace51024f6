#!/usr/bin/env python
"""A few direct launches of fsg_loss_main_levels on the config-2 shapes (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402

dev = torch.device("cuda")
N, K = 16, 80
inp = synthetic.train_inputs(2, N, 800, 1333, K, M=8, logits=False)
A, grids, R = inp["A"], inp["grids"], inp["R"]
g = torch.Generator(device="cuda").manual_seed(1)
cls_l = [torch.randn((N, A * K, h, w), device=dev, generator=g) - 4.595 for h, w in grids]
reg_l = [torch.randn((N, A * 4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
bets = torch.sigmoid(torch.randn((N, R), device=dev, generator=g) - 4.595)
anchors = inp["anchors"].to(dev)
gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
params = fsg.DenseLossConfig(num_classes=K).loss_params(1.0, 1.0, -1.0)
m = fsg.ops.match_anchors(anchors, gt, K, bets=bets, temperature=0.1)
for _ in range(4):
    out = fsg.ops.loss_main_levels(cls_l, m["gt_classes"], params, m["stats"], delta_levels=reg_l, anchors=anchors,
                                   gt=gt, matched_idx32=m["matched_idx32"], mask=m["mask"], bets=bets)
torch.cuda.synchronize()
print("ok", float(out["scalars"][8]))
