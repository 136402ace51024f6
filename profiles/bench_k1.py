#!/usr/bin/env python
"""K1 (config 2) in pieces, us per call inside a CUDA graph: pass A alone (phases=1), pass B + fold on the same
workspace (phases=2), the whole K1 (phases=3), and a 1 M-element elementwise kernel as the latency floor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402


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


dev = torch.device("cuda")
N, K = 16, 80
inp = synthetic.train_inputs(2, N, 800, 1333, K, logits=False)
R = inp["R"]
anchors = inp["anchors"].to(dev)
bets = torch.sigmoid(torch.randn((N, R), device=dev) - 4.6)
gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
L = fsg.ops.lib()
ws = torch.zeros(L.fsg_match_workspace_bytes(N, R, gt.total), dtype=torch.uint8, device=dev)
out = {k: torch.empty((N, R), dtype=dt, device=dev) for k, dt in
       (("gt_classes", torch.int64), ("mask", torch.int64), ("matched_idx32", torch.int32))}


def k1(ph):
    return lambda: fsg.ops.match_anchors(anchors, gt, K, bets=bets, temperature=0.1, phases=ph, workspace=ws, out=out)


y = torch.zeros(N * R, device=dev)
print("floor (y += 1 on N*R floats) %.1f us" % graph_us(lambda: y.add_(1.0)))
print("K1 pass A            %.1f us" % graph_us(k1(1)))
print("K1 pass B + fold     %.1f us" % graph_us(k1(2)))
print("K1 A + B + fold      %.1f us" % graph_us(k1(3)))
nobets = lambda ph, lq: (lambda: fsg.ops.match_anchors(anchors, gt, K, phases=ph, workspace=ws, out=out,
                                                       allow_low_quality_matches=lq))
print("K1 pass B alone      %.1f us   (no pre-pass sums: no fold)" % graph_us(nobets(2, True)))
print("K1 fold alone        %.1f us   (allow_low_quality_matches=False: no pass B)" % graph_us(
    lambda: fsg.ops.match_anchors(anchors, gt, K, bets=bets, temperature=0.1, phases=2, workspace=ws, out=out,
                                  allow_low_quality_matches=False)))
