"""CPU, world_size 2 over gloo: the multi-GPU form of the step.

Each rank owns a contiguous block of images.  The per-rank stage arithmetic (what the CUDA kernels do on a
GPU) is supplied by the oracle; the *exchange* -- the product's `sharded` helpers -- runs for real over a
process group.  The sharded result must equal the single-process run on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dense_oracle as orc

K, N, H, W, M = 6, 4, 96, 128, 3
COEFFS = (1.0, 0.5, -2.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_stage(inp, sl, output, nf_global=None, S_batch_global=None):
    """Per-rank sums exactly as the kernels produce them: stats = [nf, S_batch, S[n]...], then with the
    *global* normalisers the five loss sums and the gradients."""
    from full_scale_gambler_for_object_detection_b200 import _lib

    T = 0.1
    gt = orc.ground_truth(inp["anchors"], inp["gt_boxes"][sl], inp["gt_classes"][sl], K)
    g = gt["gt_classes"]
    fg = (g >= 0) & (g != K)
    w = inp["bets"][sl] * gt["mask"] + T
    stats = torch.zeros(_lib.STATS_HEADER + g.shape[0], dtype=torch.float64)
    stats[0], stats[1] = fg.sum(), w.sum()
    stats[2:] = w.sum(dim=1)
    if nf_global is None:
        return stats
    x = inp["logits"][sl].clone().requires_grad_(True)
    d = inp["deltas"][sl].clone().requires_grad_(True)
    b = inp["bets"][sl].clone().requires_grad_(True)
    t, _ = orc.one_hot_targets(g.flatten(), K, x.detach().reshape(-1, K))
    focal = orc.cls_loss_elementwise(x.reshape(-1, K), t, "focal", 0.25, 2.0).reshape(x.shape)
    ell = (focal * (g >= 0)[..., None]).sum(dim=2)
    wb = b * gt["mask"] + T
    w_hat = wb / (S_batch_global if output == "L_BAHW_extendtobatch" else wb.sum(dim=1, keepdim=True))
    reg = orc.smooth_l1(d[fg], gt["gt_deltas"][fg], 0.1).sum()
    sums = torch.stack((ell.sum(), reg, (w_hat * ell).sum(), ell.detach().sum(),
                        ell.detach().max(dim=1).values.sum())).double()
    nf = max(1.0, float(nf_global))
    # main pass: gradients wrt logits and deltas (bets enter through the detached weights, as in the kernel)
    total = COEFFS[0] * ell.sum() / nf + COEFFS[1] * reg / nf + COEFFS[2] * (-(w_hat.detach() * ell).sum())
    total.backward()
    post = dict(ell=ell.detach(), w_hat=w_hat.detach(), mask=gt["mask"].to(torch.float32),
                S=(torch.as_tensor(float(S_batch_global)) if output == "L_BAHW_extendtobatch"
                   else wb.detach().sum(dim=1, keepdim=True)))
    return sums, x.grad, d.grad, post


def _post_stage(post, A, output):
    """d(c_gam*G)/d bets = -c_gam (m/S)(l - A), A per image (L_BAHW) or over the whole batch (extendtobatch)."""
    if output == "L_BAHW_extendtobatch":
        a = torch.as_tensor(float(A))
    else:
        a = (post["w_hat"] * post["ell"]).sum(dim=1, keepdim=True)
    return COEFFS[2] * (-(post["mask"] / post["S"]) * (post["ell"] - a))


def _worker(rank, world, port, output, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from full_scale_gambler_for_object_detection_b200 import sharded, synthetic

        torch.set_num_threads(1)
        inp = synthetic.train_inputs(21, N, H, W, K, M=M)
        sl = sharded.image_shard(N, world, rank)
        stats = _rank_stage(inp, sl, output)
        local_S = stats[2:].clone()
        sharded.all_reduce_stats(stats, dist.group.WORLD)             # the exchange before the main pass
        assert torch.equal(stats[2:], local_S)                        # per-image normalisers stay local
        sums, gx, gd, post = _rank_stage(inp, sl, output, stats[0], stats[1])
        scalars = torch.zeros(10, dtype=torch.float64)
        scalars[:5] = sums
        losses = sharded.global_losses(scalars, stats, COEFFS, dist.group.WORLD)
        if output == "L_BAHW_extendtobatch":
            sharded.all_reduce_batch_weighted_sum(scalars, dist.group.WORLD)   # the exchange before the post pass
        gb = _post_stage(post, scalars[2], output)
        ret[rank] = dict(sl=(sl.start, sl.stop), stats=stats[:2].clone().numpy(), losses=losses.detach().numpy(),
                         gx=gx.detach().numpy(), gd=gd.detach().numpy(), gb=gb.detach().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("output", ["L_BAHW", "L_BAHW_extendtobatch"])
def test_sharded_step_equals_single_process(output):
    from full_scale_gambler_for_object_detection_b200 import synthetic

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), output, ret), nprocs=world, join=True)
    inp = synthetic.train_inputs(21, N, H, W, K, M=M)
    want = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                          inp["bets"], K, *COEFFS, output=output)
    for rank in range(world):
        r = ret[rank]
        a, b = r["sl"]
        assert int(r["stats"][0]) == int(want["num_foreground"])      # global foreground count on every rank
        got = r["losses"]
        for i, k in enumerate(("loss_cls", "loss_box_reg", "gambler_loss", "total")):
            assert abs(float(got[i]) - float(want[k])) <= 2e-5 * abs(float(want[k])), (k, float(got[i]), float(want[k]))
        for name, g in (("grad_logits", r["gx"]), ("grad_deltas", r["gd"]), ("grad_bets", r["gb"])):
            g = torch.from_numpy(g)
            w = want[name][a:b]
            assert float((g - w).abs().max()) <= 1e-5 * float(want[name].abs().max()) + 1e-12, name


# ---------------------------------------------------------------------------------------------------------
# one image, anchors sharded by range (the matcher stress with fewer images than GPUs)
# ---------------------------------------------------------------------------------------------------------
def _range_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from full_scale_gambler_for_object_detection_b200 import sharded, synthetic

        torch.set_num_threads(1)
        inp = synthetic.matcher_stress_inputs(31, 1, 6001, 37)
        anchors, gt = inp["anchors"][0], inp["gt_boxes"][0]
        lo, hi = sharded.anchor_range(anchors.shape[0], world, rank)
        iou = orc.pairwise_iou(gt, anchors[lo:hi])                         # pass A on the local anchors
        local_max = iou.max(dim=1).values
        bits = local_max.contiguous().view(torch.int32).clone()            # fp32 bit patterns, all >= 0
        sharded.all_reduce_gt_max(bits, dist.group.WORLD)                  # the exchange between the passes
        gmax = bits.view(torch.float32)
        best, idx = iou.max(dim=0)                                         # pass B on the local anchors
        labels = torch.full_like(idx, 1, dtype=torch.int8)
        labels[best < 0.5] = -1
        labels[best < 0.4] = 0
        labels[(iou == gmax[:, None]).any(dim=0)] = 1                      # matcher.py:99-132 with the GLOBAL maxima
        ret[rank] = dict(lo=lo, hi=hi, matches=idx.numpy(), labels=labels.numpy(), gmax=gmax.numpy())
    finally:
        dist.destroy_process_group()


def test_anchor_range_sharded_matching_equals_single_process():
    from full_scale_gambler_for_object_detection_b200 import sharded, synthetic

    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_range_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    inp = synthetic.matcher_stress_inputs(31, 1, 6001, 37)
    iou = orc.pairwise_iou(inp["gt_boxes"][0], inp["anchors"][0])
    want_m, want_l = orc.matcher(iou, [0.4, 0.5], [0, -1, 1], True)
    covered = 0
    for rank in range(world):
        r = ret[rank]
        assert (r["lo"], r["hi"]) == sharded.anchor_range(6001, world, rank)
        assert torch.equal(torch.from_numpy(r["gmax"]), iou.max(dim=1).values)      # every rank sees the global maxima
        assert torch.equal(torch.from_numpy(r["matches"]), want_m[r["lo"]:r["hi"]])
        assert torch.equal(torch.from_numpy(r["labels"]), want_l[r["lo"]:r["hi"]])
        covered += r["hi"] - r["lo"]
    assert covered == 6001
