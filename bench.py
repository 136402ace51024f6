#!/usr/bin/env python
"""Benchmark of the per-anchor dense-detection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-secondary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the fused match + gambler-loss forward/backward (K1 + K2: five kernels enqueued by one call
into the library, chained with programmatic dependent launch, replayed as one CUDA graph) over one batch of synthetic
COCO-shaped input: BASELINE config 2 -- RetinaNet R50-FPN + gambler, 800x1333 images (padded to 800x1344),
16 images per GPU, K = 80 classes, A = 3 anchors/cell, R = 67 200 anchors/image, 8 GT boxes/image with one
GT-free image.  With N GPUs every rank owns 16 images (weak scaling); the only exchange is the all-reduce of
[num_foreground, S_batch] between the matching and the loss kernels.

Prints ONE JSON line (rank 0).  ``value`` = anchors/s with inputs resident in HBM, timed with CUDA events,
max over ranks.  ``e2e`` = the same step through the public API from pinned HOST buffers (H2D of logits,
deltas, bets, GT each step and a D2H read of the loss; ``e2e.with_grads_d2h`` also copies the gradients back).
``roofline`` = the dominant kernel (K2 main pass) against the measured HBM copy bandwidth.  ``cpu_baseline`` =
the oracle port (the reference's algorithm in torch-CPU ops) timed on this box's host cores on the same batch.
``secondary`` = the other BASELINE configs: config 3 (LVIS, K = 1230, 8 images per GPU, at every N), config 4
(inference, N = 1), config 5 (matcher stress; at N > 1 sharded by image and by anchor range).  ``parity`` (N > 1) =
sharded-vs-whole-batch check of the multi-GPU path on a small batch, run inside the benchmark on every rank.
``--impl reference`` times only the CPU arm.
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
SEED_CONFIG2 = 2          # synthetic.train_inputs seed of rank 0's batch (rank r: 2 + 1000 r)
LVIS_K, LVIS_IMGS_PER_GPU = 1230, 8


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of one launch of the dominant kernel from the round's committed `ncu --set full` capture
    (profiles/traffic.json, written by profiles/ncu_traffic.py from the .ncu-rep): (bytes or None, provenance)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            t = json.load(f)
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), {k: t[k] for k in ("capture", "kernel") if k in t}
    return None, None


def port_calibration():
    """How the oracle port's speed relates to the live reference (oracle/PORT_VS_LIVE.json, measured in the build
    container by oracle/port_vs_live.py -- /root/reference does not exist on the GPU box)."""
    p = os.path.join(ROOT, "oracle", "PORT_VS_LIVE.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return None


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
# CPU arm
# ----------------------------------------------------------------------------------------------------
def cpu_reference_step_fn():
    """The reference's CPU implementation of the step (oracle port) on rank 0's whole config-2 batch: the same
    seed, the same 16 images, the GT-free image included."""
    from oracle import dense_oracle as orc
    from full_scale_gambler_for_object_detection_b200 import synthetic

    inp = synthetic.train_inputs(SEED_CONFIG2, IMGS_PER_GPU, IMG_H, IMG_W, K_CLASSES, M=GT_PER_IMG)

    def step():
        out = orc.train_step(inp["anchors"], inp["gt_boxes"], inp["gt_classes"], inp["logits"], inp["deltas"],
                             inp["bets"], K_CLASSES, 1.0, 1.0, -1.0)
        return float(out["total"])

    return step, IMGS_PER_GPU * inp["R"]


def time_cpu(step, reps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone (the other ranks exit or
    idle), so it takes every core of the host."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path for this step, all host threads.  Every step is rank 0's
    whole per-GPU batch (16 images, the same seed and GT-free image as the GPU arm); anchors/s of the CPU path does
    not depend on how many such batches a job holds, so at N > 1 the line is the same measurement (rank 0 only)."""
    if rank != 0:
        return
    cores = use_all_host_cores()
    step, anchors = cpu_reference_step_fn()
    ts = time_cpu(step, args.steps, max(1, min(args.warmup, 3)))
    sec = sum(ts) / len(ts)
    val = anchors / sec
    sample = ("rank 0's whole per-GPU batch per step (%d images, seed %d, one GT-free image); the job's %d x %d images "
              "scale linearly" % (IMGS_PER_GPU, SEED_CONFIG2, world, IMGS_PER_GPU))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "port_vs_live": port_calibration()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# helpers shared by the GPU legs
# ----------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, dev, rank, world, group, peer):
        self.dev, self.rank, self.world, self.group, self.peer = dev, rank, world, group, peer

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps, warmup=3):
        """ms per call: `reps` back-to-back calls between two CUDA events on the launching stream, barrier +
        synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b) / reps)


def secondary_metrics(ctx, fsg, with_cpu=True):
    """NMS images/s (config 4) and the matcher stress (config 5) on one GPU; reported beside the headline."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    dev = ctx.dev
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
    reps = 10
    ms = ctx.timed(lambda: fsg.ops.detect(logits, deltas, anchors, offs), reps)
    scan_bytes = 4.0 * K_CLASSES * inp["R"] * N4
    st = fsg.ops.detect(logits, deltas, anchors, offs, want_candidates=True)["slab_status"]
    out["detect_config4"] = {"images_per_s": N4 / (ms * 1e-3), "ms_per_batch": ms, "batch": N4,
                             "scan_hbm_frac": scan_bytes / (ms * 1e-3) / (hbm * 1e9), "layout": "(N, sum HWA, K)",
                             "launches": "sample bar, scan, finalize, exact fallback (idle), NMS",
                             "slabs_redone_by_exact_fallback": int((st == 2).sum().item())}
    # the same batch as the head produces it: per-level (N, A*K, H, W) / (N, A*4, H, W), read in place
    A, HW = 3, (80, 100)
    g2 = torch.Generator(device=dev).manual_seed(5)
    xs = [torch.randn((N4, A * K_CLASSES) + HW, device=dev, generator=g2) * 1.5 + synthetic.PRIOR_LOGIT
          for _ in range(5)]
    ds = [torch.randn((N4, A * 4) + HW, device=dev, generator=g2) * 0.2 for _ in range(5)]
    ms_n = ctx.timed(lambda: fsg.ops.detect_levels(xs, ds, anchors, K_CLASSES), reps)
    ms_p = ctx.timed(lambda: fsg.ops.detect(fsg.ops.levels_to_flat(xs, K_CLASSES), fsg.ops.levels_to_flat(ds, 4),
                                            anchors, offs), reps)
    out["detect_config4_native_layout"] = {
        "images_per_s": N4 / (ms_n * 1e-3), "ms_per_batch": ms_n, "batch": N4, "layout": "per-level (N, A*K, H, W)",
        "scan_hbm_frac": scan_bytes / (ms_n * 1e-3) / (hbm * 1e9), "permute_cat_flow_ms_per_batch": ms_p}
    del xs, ds
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
    out["native_layout_step_config2"] = native_layout_step(ctx, fsg)
    # config 5: 200 GT x 1M anchors per image, 8 images, allow_low_quality_matches
    inp5 = synthetic.matcher_stress_inputs(5, 8, 1000000, 200)
    a5 = inp5["anchors"].to(dev)
    gt5 = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"], inp5["gt_classes"], dev)
    ms5 = ctx.timed(lambda: fsg.ops.match_anchors(a5, gt5, 80, want=("matches", "match_labels"),
                                                  picky_thresholds=None), reps)
    props = torch.cuda.get_device_properties(dev)
    fp32_peak = props.multi_processor_count * 128 * 1.965e9          # lanes x max SM clock (instr/s)
    pairs = 8e6 * 200 / (ms5 * 1e-3)
    out["match_config5"] = {"anchors_per_s": 8e6 / (ms5 * 1e-3), "ms_per_batch": ms5,
                            "iou_pairs_per_s": pairs,
                            "hbm_frac": 25.0 * 8e6 / (ms5 * 1e-3) / (hbm * 1e9),
                            # SURVEY 8d model: ~16 fp32 instructions + 1 IEEE divide per pair
                            "fp32_issue_frac_model": pairs * 17.0 / fp32_peak,
                            "bound": "instruction issue (ncu, profiles/r2_match_crowded_ncu_summary.txt: pass A 17.8 "
                                     "instr/pair = 9.1 screen + deferred exact pairs, 78% of issue slots, ALU pipe 63%)"}
    if with_cpu:
        from oracle import dense_oracle as orc
        a_cpu, g_cpu, c_cpu = inp5["anchors"][0][:100000], inp5["gt_boxes"][:1], inp5["gt_classes"][:1]
        ts = time_cpu(lambda: orc.ground_truth([a_cpu], g_cpu, c_cpu, 80), 2, 1)
        out["match_config5"]["cpu_baseline"] = {"anchors_per_s": 1e5 / min(ts), "cores": torch.get_num_threads(),
                                                "kind": "port",
                                                "sample": "200 GT x 100k anchors of one image (both matchers), min of 2"}
    return out


def native_layout_step(ctx, fsg):
    """Config 2 again, but from what the head really produces: per-level (N, A*K, H, W) logits, (N, A*4, H, W)
    deltas and (N, A, H, W) betting maps, gradients delivered in the same layout (DenseStepPlanLevels: K2 reads the
    conv outputs in place).  Beside it: the reference's data flow on our kernels (permute + cat to (N, R, K), the
    flat fused step, inverse permutes of the gradients)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    dev = ctx.dev
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
        ms = ctx.timed(gr.replay, reps, warmup=1)
        del keep
        return ms

    hbm, _ = measured_peaks()
    plan = fsg.DenseStepPlanLevels(N, grids, A, K, cfg, dev)
    plan.capture(xs, ds, bs, anchors, gt)
    ms_n = ctx.timed(plan.replay, 20, warmup=2)
    plan.release_graphs()
    ms_p = graph_ms(permuted)
    step_bytes = (8 * K + 76) * N * R
    return {"anchors_per_s": N * R / (ms_n * 1e-3), "ms_per_step": ms_n,
            "step_hbm_frac": step_bytes / (ms_n * 1e-3) / 1e9 / hbm,
            "launches": "K1 (pass A, patch pass, fold), K2 native, K2 post native: one call, PDL chain, CUDA graph",
            "permute_cat_flow_ms_per_step": ms_p, "speedup_vs_permute_cat_flow": ms_p / ms_n}


def lvis_config3(ctx, fsg):
    """BASELINE config 3: LVIS RetinaNet + gambler, K = 1230 classes, batch 64 over 8 GPUs = 8 images per GPU
    (quick_schedules/lvis.yaml:4-7,25), at EVERY N (weak scaling: each rank owns 8 images; the exchange is the same
    all-reduce of [num_foreground, S_batch]).  Both layouts: the reference's flattened (N, R, K) and the head's
    per-level (N, A*K, H, W).  Logits are generated on the device (661 M values per rank)."""
    from full_scale_gambler_for_object_detection_b200 import synthetic

    dev, K, N = ctx.dev, LVIS_K, LVIS_IMGS_PER_GPU
    hbm, _ = measured_peaks()
    inp = synthetic.train_inputs(3 + 1000 * ctx.rank, N, IMG_H, IMG_W, K, M=GT_PER_IMG, logits=False)
    A, grids, R = inp["A"], inp["grids"], inp["R"]
    anchors = inp["anchors"].to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
    cfg = fsg.DenseLossConfig(num_classes=K)
    g = torch.Generator(device=dev).manual_seed(30 + ctx.rank)
    step_bytes = (8 * K + 76) * N * R
    res = {"workload": "config3: K=1230, %d img/GPU x %d GPUs, R=67200 anchors/img, %d GT/img (one GT-free image)"
                       % (N, ctx.world, GT_PER_IMG),
           "bytes_per_anchor": 8 * K + 76, "anchors_per_step_per_gpu": N * R}
    reps = 10
    # ---- flat (N, R, K)
    x = torch.randn((N, R, K), device=dev, generator=g) + synthetic.PRIOR_LOGIT
    d = torch.randn((N, R, 4), device=dev, generator=g) * 0.1
    b = torch.sigmoid(torch.randn((N, R), device=dev, generator=g) + synthetic.PRIOR_LOGIT)
    plan = fsg.DenseStepPlan(N, R, K, cfg, dev, group=ctx.group, peer=ctx.peer)
    plan.capture(x, d, b, anchors, gt)
    ms = ctx.timed(plan.replay, reps, warmup=2)
    main_ms = ctx.timed(lambda: plan.stage_main(x, d, b, anchors, gt), reps, warmup=1)
    res["flat"] = {"ms_per_step": ms, "anchors_per_s": ctx.world * N * R / (ms * 1e-3),
                   "step_hbm_frac": step_bytes / (ms * 1e-3) / 1e9 / hbm, "loss_main_ms": main_ms,
                   "loss_main_hbm_frac": (8 * K + 72) * N * R / (main_ms * 1e-3) / 1e9 / hbm}
    plan.release_graphs()
    del plan, x, d, b
    torch.cuda.empty_cache()
    # ---- native per-level layout
    xs = [torch.randn((N, A * K, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT for h, w in grids]
    ds = [torch.randn((N, A * 4, h, w), device=dev, generator=g) * 0.1 for h, w in grids]
    bs = [torch.sigmoid(torch.randn((N, A, h, w), device=dev, generator=g) + synthetic.PRIOR_LOGIT) for h, w in grids]
    planl = fsg.DenseStepPlanLevels(N, grids, A, K, cfg, dev, group=ctx.group, peer=ctx.peer)
    planl.capture(xs, ds, bs, anchors, gt)
    ms = ctx.timed(planl.replay, reps, warmup=2)
    res["native_layout"] = {"ms_per_step": ms, "anchors_per_s": ctx.world * N * R / (ms * 1e-3),
                            "step_hbm_frac": step_bytes / (ms * 1e-3) / 1e9 / hbm}
    planl.release_graphs()
    del planl, xs, ds, bs
    torch.cuda.empty_cache()
    return res


def match_config5_sharded(ctx, fsg):
    """BASELINE config 5 on N > 1 GPUs: 8 images x (200 GT x 1 M anchors), allow_low_quality_matches.
    (a) sharded by image (8/N images per rank, no collective); (b) every image's anchors sharded by range over the
    ranks: pass A locally, all-reduce(MAX) of the 8 x 200 per-GT maxima over NCCL, pass B locally -- timed with the
    collective.  Strong scaling: the job is the same 8 M anchors at every N."""
    from full_scale_gambler_for_object_detection_b200 import sharded, synthetic

    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    inp5 = synthetic.matcher_stress_inputs(5, 8, 1000000, 200)
    res = {"workload": "config5: 8 images x 200 GT x 1M anchors, one matcher, low-quality matches on", "n_gpus": world}
    reps = 10
    want = ("matches", "match_labels")
    if 8 % world == 0:
        sl = sharded.image_shard(8, world, rank)
        a = inp5["anchors"][sl].contiguous().to(dev)
        gt = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"][sl], inp5["gt_classes"][sl], dev)
        ms = ctx.timed(lambda: fsg.ops.match_anchors(a, gt, 80, want=want, picky_thresholds=None), reps)
        res["by_image"] = {"ms_per_batch": ms, "anchors_per_s": 8e6 / (ms * 1e-3), "images_per_rank": 8 // world,
                           "collective": "none"}
        del a, gt
    lo, hi = sharded.anchor_range(1000000, world, rank)
    a = inp5["anchors"][:, lo:hi].contiguous().to(dev)
    gt = fsg.ops.PackedGT.from_lists(inp5["gt_boxes"], inp5["gt_classes"], dev)
    ms = ctx.timed(lambda: sharded.match_anchor_range(a, gt, 80, group=ctx.group, want=want, picky_thresholds=None),
                   reps)
    res["by_anchor_range"] = {"ms_per_batch": ms, "anchors_per_s": 8e6 / (ms * 1e-3),
                              "anchors_per_rank_per_image": hi - lo,
                              "collective": "NCCL all-reduce(MAX) of 8 x 200 per-GT maxima between the two passes"}
    return res


def sharded_parity(ctx, fsg):
    """N > 1: the sharded step must equal the same step run on the WHOLE batch by one GPU (SURVEY 8e parity
    definition).  Every rank runs the whole small batch itself (no group) and compares its own shard of the sharded
    run: integer outputs bit-exact, num_foreground equal, losses and gradients within 1e-6 of the tensor scale; for
    the NCCL exchange and (when available) the in-kernel peer-memory exchange, direct launches and CUDA graphs,
    L_BAHW and L_BAHW_extendtobatch; then the anchor-range sharded matcher against the unsharded one, bit-exact.
    Returns {"status": "ok" | "FAILED: ...", ...} (identical on every rank)."""
    import torch.distributed as dist
    from full_scale_gambler_for_object_detection_b200 import sharded, synthetic

    dev, rank, world, group = ctx.dev, ctx.rank, ctx.world, ctx.group
    K, per = 80, 2
    N = per * world
    inp = synthetic.train_inputs(31, N, 320, 448, K, M=6)
    R = inp["R"]
    sl = sharded.image_shard(N, world, rank)
    coeffs = (1.0, 0.5, -2.0)
    anchors = inp["anchors"].to(dev)
    fails, checks, worst = [], 0, 0.0

    def close(a, b, tol, what):
        nonlocal checks, worst
        err = float((a.double() - b.double()).abs().max())
        scale = float(b.double().abs().max())
        rel = err / scale if scale > 0 else err
        worst = max(worst, rel)
        checks += 1
        if not rel <= tol:
            fails.append("%s: rel err %.3g > %.1g" % (what, rel, tol))

    def same(a, b, what):
        nonlocal checks
        checks += 1
        if not torch.equal(a, b):
            fails.append("%s: not bit-exact" % what)

    exchanges = [("nccl", None)] + ([("peer", ctx.peer)] if ctx.peer is not None else [])
    for output in ("L_BAHW", "L_BAHW_extendtobatch"):
        cfg = fsg.DenseLossConfig(num_classes=K, gambler_output=output)
        gt_all = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
        full = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs)
        rf = full.run(inp["logits"].to(dev), inp["deltas"].to(dev), inp["bets"].to(dev), anchors, gt_all)
        want = dict(scal=rf.scalars.clone(), nf=float(rf.stats[0]), gl=full.grad_logits[sl].clone(),
                    gd=full.grad_deltas[sl].clone(), gb=full.grad_bets[sl].clone(),
                    gtc=rf.gt_classes[sl].clone(), mask=rf.mask[sl].clone())
        gt = fsg.ops.PackedGT.from_lists(inp["gt_boxes"][sl], inp["gt_classes"][sl], dev)
        x, d, b = (inp[k][sl].contiguous().to(dev) for k in ("logits", "deltas", "bets"))
        for xname, peer in exchanges:
            plan = fsg.DenseStepPlan(per, R, K, cfg, dev, coeffs, group=group, peer=peer)
            for mode in ("direct", "graph"):
                tag = "%s/%s/%s" % (output, xname, mode)
                if mode == "graph":
                    plan.capture(x, d, b, anchors, gt)
                    r = plan.replay()
                else:
                    r = plan.run(x, d, b, anchors, gt)
                checks += 1
                if float(r.stats[0]) != want["nf"]:
                    fails.append("%s: num_foreground %r vs %r" % (tag, float(r.stats[0]), want["nf"]))
                g = sharded.global_losses(r.scalars, r.stats, coeffs, group,
                                          batch_sum_is_global=(output == "L_BAHW_extendtobatch"))
                for i, j in enumerate((5, 6, 7, 8)):
                    close(g[i], want["scal"][j], 1e-6, tag + " loss %d" % j)
                same(r.gt_classes, want["gtc"], tag + " gt_classes")
                same(r.mask, want["mask"], tag + " mask")
                close(plan.grad_logits, want["gl"], 1e-6, tag + " grad_logits")
                close(plan.grad_deltas, want["gd"], 1e-6, tag + " grad_deltas")
                close(plan.grad_bets, want["gb"], 2e-6, tag + " grad_bets")
            plan.release_graphs()
            if peer is not None:
                peer.check()
        # DDP semantics: a parameter gradient averaged over ranks (what DistributedDataParallel does) must equal the
        # single-process whole-batch gradient.  theta scales the logits; d total / d theta = sum(grad_logits * x).
        xa = inp["logits"].to(dev)
        want_theta = float((full.grad_logits.double() * xa.double()).sum())
        xr = x.clone().requires_grad_(True)
        res = fsg.dense_train_step(xr, d.clone().requires_grad_(True), b.clone().requires_grad_(True), anchors, gt,
                                   cfg, coeffs, group=group, grad_reduction="mean")
        res.total.backward()
        th = (xr.grad.double() * x.double()).sum().reshape(1)
        dist.all_reduce(th, op=dist.ReduceOp.SUM, group=group)
        th = float(th) / world            # DDP: mean over ranks
        checks += 1
        worst = max(worst, abs(th - want_theta) / abs(want_theta))
        if not abs(th - want_theta) <= 1e-5 * abs(want_theta):
            fails.append("%s: DDP-averaged parameter gradient %.9g vs whole-batch %.9g" % (output, th, want_theta))
    # ---- anchor-range sharded matcher
    ms = synthetic.matcher_stress_inputs(35, 1, 300001, 200)
    a_all = ms["anchors"][0]
    gts = fsg.ops.PackedGT.from_lists(ms["gt_boxes"], ms["gt_classes"], dev)
    keys = ("matches", "match_labels", "gt_classes")
    whole = fsg.ops.match_anchors(a_all.to(dev), gts, 80, want=keys, picky_thresholds=None)
    lo, hi = sharded.anchor_range(a_all.shape[0], world, rank)
    mine = sharded.match_anchor_range(a_all[lo:hi].to(dev), gts, 80, group=group, want=keys, picky_thresholds=None)
    for k in keys:
        same(mine[k], whole[k][:, lo:hi], "anchor-range sharded %s" % k)
    torch.cuda.synchronize()
    bad = torch.tensor([len(fails)], dtype=torch.int64, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.SUM, group=group)
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    status = "ok" if int(bad.item()) == 0 else "FAILED: %d check(s) over all ranks; rank %d: %s" % (
        int(bad.item()), rank, "; ".join(fails[:4]) or "-")
    return {"status": status, "checks_per_rank": checks, "max_rel_err": float(w.item()),
            "exchanges": [n for n, _ in exchanges], "batch": "%d img/rank x %d ranks, 320x448, K=80" % (per, world),
            "reference": "the same step on the whole batch by one GPU (SURVEY 8e); ints bit-exact, floats <= 1e-6 of scale",
            "ddp_parameter_gradient": "mean over ranks of grad (grad_reduction='mean') == whole-batch gradient"}


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
    inp = synthetic.train_inputs(SEED_CONFIG2 + 1000 * rank, IMGS_PER_GPU, IMG_H, IMG_W, K_CLASSES, M=GT_PER_IMG)
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
    ctx = Ctx(dev, rank, world, group, peer)
    plan = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs, group=group, peer=peer)
    x_s, d_s, b_s = logits.detach(), deltas.detach(), bets.detach()
    graphed = False
    try:
        # N = 1 (or peer exchange): the whole step is one graph launch; N > 1 with NCCL: two graphs with the eager
        # all-reduce of [num_foreground, S_batch] between them (collectives stay outside graphs)
        plan.capture(x_s, d_s, b_s, anchors, gt)
        graphed = True
    except Exception as e:  # fall back to direct launches
        if rank == 0:
            print("graph capture unavailable (%s); timing direct launches" % (type(e).__name__,), file=sys.stderr)
        plan.graph = None

    def step():
        return plan.replay() if graphed else plan.run(x_s, d_s, b_s, anchors, gt)

    barrier = ctx.barrier
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
    ms_per_step = ctx.max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    value = world * N * R / (ms_per_step * 1e-3)

    from full_scale_gambler_for_object_detection_b200 import sharded as _sh
    glob = _sh.global_losses(res.scalars, res.stats, coeffs, group).tolist()   # collective: every rank calls it
    # (config 2 is L_BAHW: per-image normaliser, so scalars[2] is a per-rank sum here)
    num_fg = float(res.num_foreground.item())
    if peer is not None:
        peer.check()            # a peer that never arrived poisons stats with NaN and raises here

    # ---- per-kernel device time: each stage launched back to back between two CUDA events on the launching
    #      stream (no host gaps: the queue stays ahead of the device); inputs exceed L2
    reps = max(10, min(args.steps, 100))
    match_ms = ctx.timed(lambda: plan.stage_match(b_s, anchors, gt), reps)
    main_ms = ctx.timed(lambda: plan.stage_main(x_s, d_s, b_s, anchors, gt), reps)
    post_ms = ctx.timed(lambda: plan.stage_post(b_s), reps)

    # ---- end to end: host buffers in, loss out, through the public API
    h2d = sum(host[k].numel() * 4 for k in host) + sum(b.numel() * 4 + c.numel() * 8 for b, c in
                                                       zip(inp["gt_boxes"], inp["gt_classes"])) + 4 * (N + 1)
    grad_bytes = 4 * (N * R * K + N * R * 4 + N * R)
    pinned_out = None

    def e2e_step(copy_grads=False):
        x = host["logits"].to(dev, non_blocking=True).requires_grad_(True)
        d = host["deltas"].to(dev, non_blocking=True).requires_grad_(True)
        b = host["bets"].to(dev, non_blocking=True).requires_grad_(True)
        g = fsg.ops.PackedGT.from_lists(inp["gt_boxes"], inp["gt_classes"], dev)
        r = fsg.dense_train_step(x, d, b, anchors, g, cfg, coeffs, group=group, plan=plan_e2e)
        r.total.backward()          # grads are produced by the same fused kernels; this only hands them to autograd
        if copy_grads:
            pinned_out[0].copy_(x.grad, non_blocking=True)
            pinned_out[1].copy_(d.grad, non_blocking=True)
            pinned_out[2].copy_(b.grad, non_blocking=True)
        return r.total.item()

    plan_e2e = fsg.DenseStepPlan(N, R, K, cfg, dev, coeffs, group=group, peer=peer)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_time(copy_grads):
        for _ in range(3):
            e2e_step(copy_grads)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            lv = e2e_step(copy_grads)
        b.record()
        barrier()
        return ctx.max_over_ranks(a.elapsed_time(b)) / e2e_steps, lv

    e2e_ms, loss_val = e2e_time(False)
    e2e_value = world * N * R / (e2e_ms * 1e-3)
    pinned_out = [torch.empty(s, dtype=torch.float32).pin_memory() for s in ((N, R, K), (N, R, 4), (N, R))]
    e2e_g_ms, _ = e2e_time(True)
    pinned_out = None

    secondary, parity = None, None
    if not args.no_secondary:
        secondary = {}
        if world == 1:
            secondary.update(secondary_metrics(ctx, fsg, with_cpu=not args.no_cpu))
        else:
            parity = sharded_parity(ctx, fsg)
            secondary["match_config5_sharded"] = match_config5_sharded(ctx, fsg)
        del logits, deltas, bets, x_s, d_s, b_s
        plan.release_graphs()
        torch.cuda.empty_cache()
        secondary["lvis_config3"] = lvis_config3(ctx, fsg)

    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline on rank 0 at N=1 only (the same whole batch; 1 warm-up + 2 timed passes)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = use_all_host_cores()
        stepf, anchors_cpu = cpu_reference_step_fn()
        ts = time_cpu(stepf, 2, 1)
        cpu_baseline = {"value": anchors_cpu / min(ts), "unit": UNIT, "cores": cores,
                        "kind": "port", "host_cpus": os.cpu_count(), "port_vs_live": port_calibration(),
                        "sample": "the whole config-2 batch (%d images, same seed and GT-free image as the GPU arm), "
                                  "1 warm-up + min of 2" % IMGS_PER_GPU}

    if rank == 0:
        hbm, which = measured_peaks()
        main_bytes = (8 * K + 72) * N * R                # K2 main pass, DESIGN.md section 4
        step_bytes = (8 * K + 76) * N * R                # whole fused step, SURVEY.md section 8d
        achieved = main_bytes / (main_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
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
                                    if peer is not None else "NCCL all-reduce of [num_fg, S_batch]"),
                       "tolerances": "ints bit-exact; floats |a-b| <= 1e-5|b| + 1e-7 max|b| (1e-6 max|b| for d/d bets, "
                                     "1e-5 max|b| in sigmoid mode: cancellation in the reference's own fp32, DESIGN.md 2)"},
            "roofline": {"bound": "hbm", "kernel": "loss_main_kernel (K2 main pass)", "achieved": achieved,
                         "peak": hbm, "peak_source": which, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "traffic_source": traffic_src, "bytes_per_anchor": 8 * K + 72,
                         "kernel_ms": main_ms,
                         "timing": "%d back-to-back launches between two CUDA events" % reps,
                         "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm,
                         "stage_ms": {"match": match_ms, "loss_main": main_ms, "loss_post": post_ms}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "loss": loss_val,
                    "outputs_left_on_device_bytes": grad_bytes,
                    "note": "the step's product (gradients) is consumed by autograd on the same GPU; "
                            "with_grads_d2h copies it to pinned host memory as well",
                    "with_grads_d2h": {"value": world * N * R / (e2e_g_ms * 1e-3), "ms_per_step": e2e_g_ms,
                                       "d2h_bytes_per_step": 4 + grad_bytes}},
            "cpu_baseline": cpu_baseline,
            "gpu_launches": launches,
            "clocks": clocks,
            "losses": {"loss_cls": float(glob[0]), "loss_box_reg": float(glob[1]), "gambler_loss": float(glob[2]),
                       "num_foreground": num_fg, "scope": "whole batch over all ranks"},
        }
        if secondary:
            line["secondary"] = secondary
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    failed = parity is not None and parity["status"] != "ok"
    if world > 1:
        shutdown_distributed(plan, plan_e2e)
    if failed:
        sys.exit(3)


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-3 / 4 / 5 legs and the N > 1 parity block")
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
