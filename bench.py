#!/usr/bin/env python
"""Benchmark of the per-anchor dense-detection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-secondary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the fused match + gambler-loss forward/backward (K1 + K2: four kernel launches) over
one batch of synthetic COCO-shaped input: BASELINE config 2 -- RetinaNet R50-FPN + gambler, 800x1333 images
(padded to 800x1344), 16 images per GPU, K = 80 classes, A = 3 anchors/cell, R = 67 200 anchors/image,
8 GT boxes/image with one GT-free image.  With N GPUs every rank owns 16 images (weak scaling); the only
exchange is the all-reduce of [num_foreground, S_batch] between the matching and the loss kernels.

Prints ONE JSON line (rank 0).  ``value`` = anchors/s with inputs resident in HBM, timed with CUDA events,
max over ranks.  ``e2e`` = the same step through the public API from pinned HOST buffers (H2D of logits,
deltas, bets, GT each step and a D2H read of the loss).  ``roofline`` = the dominant kernel (K2 main pass)
against the measured HBM copy bandwidth.  ``cpu_baseline`` = the oracle port (the reference's algorithm in
torch-CPU ops) timed on this box's host cores on a bounded sample.  ``--impl reference`` times only that.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "anchors/sec (match+gambler loss fwd/bwd)"
UNIT = "anchors/s"
IMG_H, IMG_W, K_CLASSES, IMGS_PER_GPU, GT_PER_IMG = 800, 1333, 80, 16, 8
WORKLOAD = ("config2: RetinaNet R50-FPN + gambler, synthetic 800x1333 (padded 800x1344), %d img/GPU, K=%d, A=3, "
            "R=67200 anchors/img, %d GT/img (one GT-free image), L_BAHW, focal(0.25,2), T=0.1"
            % (IMGS_PER_GPU, K_CLASSES, GT_PER_IMG))
CPU_SAMPLE_IMAGES = 8
# dram__bytes_read.sum + dram__bytes_write.sum of one loss_main_kernel<4,5,0,4> launch on this workload, from
# the committed `ncu --set full` capture profiles/r1c_ncu_full_raw.csv (366.09 MB read + 313.57 MB written);
# the algorithmic figure is (8K+72)*N*R = 765.5 MB -- the (N,R)-sized side inputs mostly hit L2
NCU_TRAFFIC_BYTES = 679664000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is busy (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(num_images):
    """The reference's CPU implementation of the step (oracle port), on `num_images` images of config 2."""
    from oracle import dense_oracle as orc
    from full_scale_gambler_for_object_detection_b200 import synthetic

    inp = synthetic.train_inputs(2, num_images, IMG_H, IMG_W, K_CLASSES, M=GT_PER_IMG, empty_image=False)

    def step():
        out = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                             inp["bets"], K_CLASSES, 1.0, 1.0, -1.0)
        return float(out["total"])

    return step, num_images * inp["R"]


def time_cpu(step, reps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path for this step, all host threads, bounded sample."""
    if rank != 0:
        return
    step, anchors = cpu_reference_step_fn(CPU_SAMPLE_IMAGES)
    ts = time_cpu(step, args.steps, max(1, min(args.warmup, 3)))
    sec = sum(ts) / len(ts)
    val = anchors / sec
    cores = torch.get_num_threads()
    sample = "%d of %d images of the config-2 batch per step (linear in images)" % (CPU_SAMPLE_IMAGES, IMGS_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def secondary_metrics(dev, fsg, with_cpu=True):
    """NMS images/s (config 4) and the matcher stress (config 5), short runs; reported beside the headline."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    out = {}
    hbm, _ = measured_peaks()
    # config 4: 32 images x 5 levels x 24 000 anchors, K=80, top-1000/level, NMS 0.5, 100 detections
    N4 = 32
    inp = synthetic.inference_inputs(4, 1, [24000] * 5, K_CLASSES)
    g = torch.Generator(device="cpu").manual_seed(4)
    logits = (torch.randn((N4, inp["R"], K_CLASSES), generator=g) * 1.5 + synthetic.PRIOR_LOGIT).to(dev)
    deltas = (torch.randn((N4, inp["R"], 4), generator=g) * 0.2).to(dev)
    anchors = inp["anchors"].to(dev)
    offs = inp["level_offsets"]
    run = lambda: fsg.ops.detect(logits, deltas, anchors, offs)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    scan_bytes = 4.0 * K_CLASSES * inp["R"] * N4
    out["detect_config4"] = {"images_per_s": N4 / (ms * 1e-3), "ms_per_batch": ms, "batch": N4,
                             "scan_hbm_frac": scan_bytes / (ms * 1e-3) / (hbm * 1e9)}
    if with_cpu:
        # the reference's inference_single_image (oracle port) on ONE image of the same shape, host cores
        from oracle import dense_oracle as orc
        lg, dl, an = logits[0].cpu(), deltas[0].cpu(), inp["anchors"]
        cls = [lg[offs[i]:offs[i + 1]] for i in range(5)]
        reg = [dl[offs[i]:offs[i + 1]] for i in range(5)]
        anc = [an[offs[i]:offs[i + 1]] for i in range(5)]
        ts = time_cpu(lambda: orc.inference_single_image(cls, reg, anc, K_CLASSES), 2, 1)
        out["detect_config4"]["cpu_baseline"] = {"images_per_s": 1.0 / min(ts), "cores": torch.get_num_threads(),
                                                 "kind": "port", "sample": "1 image of the batch, min of 2"}
    del logits, deltas
    out["native_layout_step_config2"] = native_layout_step(dev, fsg)
    # config 5: 200 GT x 1M anchors per image, 8 images, allow_low_quality_matches
    inp5 = synthetic.matcher_stress_inputs(5, 8, 1000000, 200)
    a5 = inp5["anchors"].to(dev)
    gt5 = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"], inp5["gt_classes"], dev)
    run5 = lambda: fsg.ops.match_anchors(a5, gt5, 80, want=("matches", "match_labels"), picky_thresholds=None)
    for _ in range(3):
        run5()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        run5()
    e1.record()
    torch.cuda.synchronize()
    ms5 = e0.elapsed_time(e1) / reps
    props = torch.cuda.get_device_properties(dev)
    fp32_peak = props.multi_processor_count * 128 * 1.965e9          # lanes x max SM clock (instr/s)
    pairs = 8e6 * 200 / (ms5 * 1e-3)
    out["match_config5"] = {"anchors_per_s": 8e6 / (ms5 * 1e-3), "ms_per_batch": ms5,
                            "iou_pairs_per_s": pairs,
                            "hbm_frac": 25.0 * 8e6 / (ms5 * 1e-3) / (hbm * 1e9),
                            # SURVEY 8d model: ~16 fp32 instructions + 1 IEEE divide per pair
                            "fp32_issue_frac_model": pairs * 17.0 / fp32_peak,
                            "bound": "fp32 issue (ncu: 90% of issue slots busy in pass A, 33 instr/pair)"}
    if with_cpu:
        from oracle import dense_oracle as orc
        a_cpu, g_cpu, c_cpu = inp5["anchors"][0][:100000], inp5["gt_boxes"][:1], inp5["gt_classes"][:1]
        ts = time_cpu(lambda: orc.ground_truth([a_cpu], g_cpu, c_cpu, 80), 2, 1)
        out["match_config5"]["cpu_baseline"] = {"anchors_per_s": 1e5 / min(ts), "cores": torch.get_num_threads(),
                                                "kind": "port",
                                                "sample": "200 GT x 100k anchors of one image (both matchers), min of 2"}
    return out


def native_layout_step(dev, fsg):
    """Config 2 again, but from what the head really produces: per-level (N, A*K, H, W) logits, (N, A*4, H, W)
    deltas and (N, A, H, W) betting maps, gradients delivered in the same layout (dense_train_step_levels:
    K2 reads the conv outputs in place).  Beside it: the reference's data flow on our kernels (permute + cat to
    (N, R, K), the flat fused step, inverse permutes of the gradients)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    N, K = IMGS_PER_GPU, K_CLASSES
    inp = synthetic.train_inputs(2, N, IMG_H, IMG_W, K, M=GT_PER_IMG, logits=False)
    A, grids, R = inp["A"], inp["grids"], inp["R"]
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn((N, A * K, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT for h, w in grids]
    ds = [torch.randn((N, A * 4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
    bs = [torch.sigmoid(torch.randn((N, A, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT) for h, w in grids]
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    cfg = fsg.DenseLossConfig(num_classes=K)
    params = cfg.loss_params(1.0, 1.0, -1.0)
    ops = fsg.ops
    shapes = [tuple(b.shape[1:]) for b in bs]

    def native():
        m = ops.match_anchors(anchors, gt, K, bet_levels=bs, temperature=cfg.gambler_temperature)
        ell = [torch.empty_like(b) for b in bs]
        o = ops.loss_main_levels(xs, m["gt_classes"], params, m["stats"], delta_levels=ds, anchors=anchors, gt=gt,
                                 matched_idx32=m["matched_idx32"], mask=m["mask"], bet_levels=bs, ell_levels_out=ell)
        return o, ops.loss_post_levels(bs, m["mask"], ell, params, m["stats"], o["scalars"])

    def permuted():
        x, d = ops.levels_to_flat(xs, K), ops.levels_to_flat(ds, 4)
        bets = ops.anchor_maps_to_flat([bs])[0]
        m = ops.match_anchors(anchors, gt, K, bets=bets, temperature=cfg.gambler_temperature)
        o = ops.loss_main(x, m["gt_classes"], params, m["stats"], pred_deltas=d, anchors=anchors, gt=gt,
                          matched_idx32=m["matched_idx32"], mask=m["mask"], bets=bets)
        gb = ops.loss_post(bets, m["mask"], o["per_anchor_loss"], params, m["stats"], o["scalars"])
        gl = ops.flat_to_levels(o["grad_logits"], [tuple(t.shape[1:]) for t in xs])
        gd = ops.flat_to_levels(o["grad_deltas"], [tuple(t.shape[1:]) for t in ds])
        return o, gl, gd, ops.anchor_maps_to_levels([gb, o["per_anchor_loss"]], shapes)

    def graph_ms(fn, reps=20):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = fn()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        del keep
        return e0.elapsed_time(e1) / reps

    hbm, _ = measured_peaks()
    ms_n = graph_ms(native)
    ms_p = graph_ms(permuted)
    step_bytes = (8 * K + 76) * N * R
    return {"anchors_per_s": N * R / (ms_n * 1e-3), "ms_per_step": ms_n,
            "step_hbm_frac": step_bytes / (ms_n * 1e-3) / 1e9 / hbm,
            "launches": "K1 x2, K2 native, K2 post native (CUDA graph)",
            "permute_cat_flow_ms_per_step": ms_p, "speedup_vs_permute_cat_flow": ms_p / ms_n}


def run_ours(args, rank, local_rank, world):
    import full_scale_gambler_for_object_detection_b200 as fsg
    from full_scale_gambler_for_object_detection_b200 import _lib, synthetic

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this path has no CPU implementation")
    fsg.ops.lib()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    # ---- synthetic batch: identical generator on every rank, different seed offset per rank
    inp = synthetic.train_inputs(2 + 1000 * rank, IMGS_PER_GPU, IMG_H, IMG_W, K_CLASSES, M=GT_PER_IMG)
    N, R, K = inp["N"], inp["R"], K_CLASSES
    host = {k: inp[k].pin_memory() for k in ("logits", "deltas", "bets")}
    anchors = inp["anchors"].to(dev)
    cfg = fsg.DenseLossConfig(num_classes=K)
    coeffs = (1.0, 1.0, -1.0)

    logits = host["logits"].to(dev).requires_grad_(True)
    deltas = host["deltas"].to(dev).requires_grad_(True)
    bets = host["bets"].to(dev).requires_grad_(True)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)

    peer = None
    if world > 1 and not args.nccl_exchange:
        from full_scale_gambler_for_object_detection_b200.sharded import PeerExchange
        peer = PeerExchange.create(group, dev)
        # all ranks must agree (a rank without P2P would otherwise wait for NCCL while the others spin)
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None
    plan = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs, group=group, peer=peer)
    x_s, d_s, b_s = logits.detach(), deltas.detach(), bets.detach()
    graphed = False
    try:
        # N = 1: the whole step (4 kernels + 2 memset nodes) is one graph launch; N > 1: two graphs with the
        # eager NCCL all-reduce of [num_foreground, S_batch] between them (collectives stay outside graphs)
        plan.capture(x_s, d_s, b_s, anchors, gt)
        graphed = True
    except Exception as e:  # fall back to direct launches
        if rank == 0:
            print("graph capture unavailable (%s); timing direct launches" % (type(e).__name__,), file=sys.stderr)
        plan.graph = None

    def step():
        return plan.replay() if graphed else plan.run(x_s, d_s, b_s, anchors, gt)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    for _ in range(args.warmup):
        step()
    # ---- timed region: exactly K steps, device-timed, barrier + synchronize on both sides
    barrier()
    launches0 = _lib.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = step()
    ev1.record()
    barrier()
    launches = _lib.LAUNCHES - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * N * R / (ms_per_step * 1e-3)

    from full_scale_gambler_for_object_detection_b200 import sharded as _sh
    glob = _sh.global_losses(res.scalars, res.stats, coeffs, group).tolist()   # collective: every rank calls it
    # (config 2 is L_BAHW: per-image normaliser, so scalars[2] is a per-rank sum here)
    num_fg = float(res.num_foreground.item())

    # ---- per-kernel device time: each stage launched back to back K times between two CUDA events on the
    #      launching stream (no host gaps: the queue stays ahead of the device); inputs exceed L2
    def stage_ms(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    reps = max(10, min(args.steps, 100))
    match_ms = stage_ms(lambda: plan.stage_match(b_s, anchors, gt), reps)
    main_ms = stage_ms(lambda: plan.stage_main(x_s, d_s, b_s, anchors, gt), reps)
    post_ms = stage_ms(lambda: plan.stage_post(b_s), reps)
    barrier()

    # ---- end to end: host buffers in, loss out, through the public API
    h2d = sum(host[k].numel() * 4 for k in host) + sum(b.numel() * 4 + c.numel() * 8 for b, c in
                                                       zip(inp["gt_boxes"], inp["gt_classes"])) + 4 * (N + 1)
    d2h = 4

    def e2e_step():
        x = host["logits"].to(dev, non_blocking=True).requires_grad_(True)
        d = host["deltas"].to(dev, non_blocking=True).requires_grad_(True)
        b = host["bets"].to(dev, non_blocking=True).requires_grad_(True)
        g = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
        r = fsg.dense_train_step(x, d, b, anchors, g, cfg, coeffs, group=group, plan=plan_e2e)
        r.total.backward()          # grads are produced by the same fused kernels; this only hands them to autograd
        return r.total.item()

    plan_e2e = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs, group=group, peer=peer)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(e2e_steps):
        loss_val = e2e_step()
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_value = world * N * R / (e2e_ms * 1e-3)

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = secondary_metrics(dev, fsg, with_cpu=not args.no_cpu)

    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on rank 0 at N=1 only (bounded sample)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        stepf, anchors_cpu = cpu_reference_step_fn(CPU_SAMPLE_IMAGES)
        ts = time_cpu(stepf, 3, 1)
        cpu_baseline = {"value": anchors_cpu / min(ts), "unit": UNIT, "cores": torch.get_num_threads(),
                        "kind": "port", "host_cpus": os.cpu_count(),
                        "sample": "%d of %d images of the config-2 batch, 1 warm-up + min of 3 (linear in images)"
                                  % (CPU_SAMPLE_IMAGES, IMGS_PER_GPU)}

    if rank == 0:
        hbm, which = measured_peaks()
        main_bytes = (8 * K + 72) * N * R                # K2 main pass, DESIGN.md section 4
        step_bytes = (8 * K + 76) * N * R                # whole fused step, SURVEY.md section 8d
        achieved = main_bytes / (main_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (344 MB logits + 344 MB grads per GPU)",
                       "anchors_per_step_per_gpu": N * R, "parallelism": "image-sharded x%d" % world,
                       "launch": ("CUDA graph replay" if (world == 1 or peer is not None)
                                  else "2 CUDA graphs + eager NCCL all-reduce") if graphed else "direct launches",
                       "exchange": ("none (1 GPU)" if world == 1 else
                                    "in-kernel all-reduce of [num_fg, S_batch] over NVLink peer memory"
                                    if peer is not None else "NCCL all-reduce of [num_fg, S_batch]")},
            "roofline": {"bound": "hbm", "kernel": "loss_main_kernel (K2 main pass)", "achieved": achieved,
                         "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": NCU_TRAFFIC_BYTES, "bytes_per_anchor": 8 * K + 72, "kernel_ms": main_ms,
                         "timing": "%d back-to-back launches between two CUDA events" % reps,
                         "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm,
                         "stage_ms": {"match_2_kernels": match_ms, "loss_main": main_ms, "loss_post": post_ms}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "loss": loss_val},
            "cpu_baseline": cpu_baseline,
            "gpu_launches": launches,
            "clocks": clocks,
            "losses": {"loss_cls": float(glob[0]), "loss_box_reg": float(glob[1]), "gambler_loss": float(glob[2]),
                       "num_foreground": num_fg, "scope": "whole batch over all ranks"},
        }
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed(plan, plan_e2e)


def shutdown_distributed(*plans):
    """Leave the process group without ever hanging the box: drop the CUDA graphs, drain the device, meet the
    other ranks, then destroy the group under a watchdog that hard-exits if teardown stalls."""
    import torch.distributed as dist

    for p in plans:
        p.release_graphs()
    torch.cuda.synchronize()
    try:
        dist.barrier()
    except Exception:
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(20.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        dist.destroy_process_group()
    finally:
        timer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-4 / config-5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="N > 1: use the NCCL all-reduce instead of the in-kernel peer-memory exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
