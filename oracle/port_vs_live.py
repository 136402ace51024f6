"""Calibration of the CPU baseline (TEST INFRASTRUCTURE, build container only): how fast is the oracle port
(``oracle/dense_oracle.py:train_step``) compared with the LIVE reference -- the reference's own source files run
through ``oracle/ref_loader.py`` (get_ground_truth, get_picky_ground_truth, losses, gambler_loss, backward; including
the permute + cat of the head outputs that the reference performs) -- on the same slice of the config-2 batch?

    python -m oracle.port_vs_live        # writes oracle/PORT_VS_LIVE.json

``bench.py`` attaches the committed result to ``cpu_baseline`` (the GPU box has no /root/reference, so the live
reference cannot be timed there).  ratio > 1: the port is FASTER than the live reference, i.e. a tougher baseline.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dense_oracle as orc  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import synthetic  # noqa: E402


def main():
    ref = rl.load_reference()
    images = 2
    inp = synthetic.train_inputs(2, images, 800, 1333, 80, M=8)
    port_fn = lambda: orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"],
                                     inp["deltas"], inp["bets"], 80, 1.0, 1.0, -1.0)
    live_fn = lambda: mg.run_reference_train(ref, inp, 80, (800, 1333))
    port_fn()
    live_fn()
    tp, tl = [], []
    for _ in range(6):            # interleaved: the container shares its cores, single timings wander by 2-3x
        for fn, acc in ((port_fn, tp), (live_fn, tl)):
            t0 = time.perf_counter()
            fn()
            acc.append(time.perf_counter() - t0)
    port, live = min(tp), min(tl)
    out = {"slice": "%d images of the config-2 batch (800x1333, K=80, 8 GT/img), 1 warm-up + min of 6 interleaved passes" % images,
           "threads": torch.get_num_threads(), "port_s": port, "live_reference_s": live,
           "port_speed_over_live": live / port, "all_port_s": tp, "all_live_s": tl,
           "where": "build container (the reference's own files through oracle/ref_loader.py)"}
    with open(os.path.join(ROOT, "oracle", "PORT_VS_LIVE.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
