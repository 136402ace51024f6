#!/usr/bin/env python
"""Multi-GPU check (launch with torchrun, one rank per GPU):

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/run_sharded_gpu.py

The batch is sharded by image; the sharded step (NCCL all-reduce of [num_foreground, S_batch] between K1 and
K2) must equal the same step run on the whole batch by one GPU: losses to 1e-6, gradients bit-for-bit up to
the fp32 rounding of the normaliser.  Checked for the direct-launch path, the graph path and
L_BAHW_extendtobatch.  Prints SHARDED_OK on success."""
import os
import sys
import threading

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import full_scale_gambler_for_object_detection_b200 as fsg  # noqa: E402
from full_scale_gambler_for_object_detection_b200 import sharded, synthetic  # noqa: E402


def close(a, b, tol, what):
    err = float((a.double() - b.double()).abs().max())
    scale = float(b.double().abs().max())
    assert err <= tol * scale + 1e-30, "%s: err %.3g scale %.3g" % (what, err, scale)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    K, per = 80, 2
    N = per * world
    inp = synthetic.train_inputs(31, N, 320, 448, K, M=6)
    R = inp["R"]
    sl = sharded.image_shard(N, world, rank)
    coeffs = (1.0, 0.5, -2.0)
    anchors = inp["anchors"].to(dev)
    peer = sharded.PeerExchange.create(dist.group.WORLD, dev)
    if rank == 0:
        print("peer-memory exchange:", "available" if peer is not None else "unavailable (NCCL only)", flush=True)
    for output, use_peer in (("L_BAHW", False), ("L_BAHW_extendtobatch", False), ("L_BAHW", True),
                             ("L_BAHW_extendtobatch", True)):
        if use_peer and peer is None:
            continue
        cfg = fsg.DenseLossConfig(num_classes=K, gambler_output=output)
        # whole batch on this GPU, no group
        gt_all = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
        full = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs)
        rf = full.run(inp["logits"].to(dev), inp["deltas"].to(dev), inp["bets"].to(dev), anchors, gt_all)
        want = dict(scal=rf.scalars.clone(), nf=rf.stats[0].clone(), gl=full.grad_logits[sl].clone(),
                    gd=full.grad_deltas[sl].clone(), gb=full.grad_bets[sl].clone())
        # my shard, with the exchange
        gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"][sl], inp["gt_classes"][sl], dev)
        x, d, b = (inp[k][sl].contiguous().to(dev) for k in ("logits", "deltas", "bets"))
        plan = fsg.DenseStepPlan(per, R, K, cfg, dev, coeffs, group=dist.group.WORLD,
                                 peer=peer if use_peer else None)
        for mode in ("direct", "graph"):
            if mode == "graph":
                plan.capture(x, d, b, anchors, gt)
                r = plan.replay()
            else:
                r = plan.run(x, d, b, anchors, gt)
            assert float(r.stats[0]) == float(want["nf"]), (float(r.stats[0]), float(want["nf"]))
            g = sharded.global_losses(r.scalars, r.stats, coeffs, dist.group.WORLD,
                                      batch_sum_is_global=(output == "L_BAHW_extendtobatch"))
            for i, j in enumerate((5, 6, 7, 8)):
                assert abs(float(g[i]) - float(want["scal"][j])) <= 1e-6 * abs(float(want["scal"][j])), \
                    (output, mode, use_peer, i)
            close(plan.grad_logits, want["gl"], 1e-6, "grad_logits")
            close(plan.grad_deltas, want["gd"], 1e-6, "grad_deltas")
            close(plan.grad_bets, want["gb"], 2e-6, "grad_bets")
        plan.release_graphs()
        if use_peer:
            for _ in range(5):      # repeated launches: epochs advance, mailboxes alternate
                r = plan.run(x, d, b, anchors, gt)
            assert float(r.stats[0]) == float(want["nf"])
            peer.check()
    # ---- one image, anchors sharded by range (matcher stress with fewer images than GPUs): pass A, NCCL
    #      all-reduce(MAX) of the per-GT maxima, pass B; every rank's slice must equal the unsharded result
    ms = synthetic.matcher_stress_inputs(35, 1, 300001, 200)
    a_all = ms["anchors"][0]
    gts = fsg.ops.PackedGT.from_lists(ms["gt_boxes"], ms["gt_classes"], dev)
    keys = ("matches", "match_labels", "gt_classes")
    whole = fsg.ops.match_anchors(a_all.to(dev), gts, 80, want=keys, picky_thresholds=None)
    lo, hi = sharded.anchor_range(a_all.shape[0], world, rank)
    mine = sharded.match_anchor_range(a_all[lo:hi].to(dev), gts, 80, group=dist.group.WORLD, want=keys,
                                      picky_thresholds=None)
    for k in keys:
        assert torch.equal(mine[k], whole[k][:, lo:hi]), "anchor-range sharded %s" % k
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK world=%d (image shards, peer exchange, anchor-range shards)" % world, flush=True)
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


if __name__ == "__main__":
    main()
