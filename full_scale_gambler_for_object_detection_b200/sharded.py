"""Multi-GPU form of the path: shard the batch by image, one process per GPU.

Matching, encode, per-anchor loss, d/d logits, decode and NMS are independent per image, so ranks own
contiguous blocks of images (the reference's data-parallel split, detectron2/data/build.py:350-362) and the
data path needs no collective.  The only exchange is a handful of scalars:

  before the main pass   all-reduce(SUM) [num_foreground, S_batch]        (the gradient scale depends on it)
  L_BAHW_extendtobatch   all-reduce(SUM) sum_n A[n]                       (before d/d bets)
  for reporting          all-reduce(SUM) of the five loss sums

Parity definition: the sharded result equals the reference run single-process on the whole batch (global
``num_foreground``), which intentionally differs from the reference's own DDP behaviour (local normaliser,
retinanet.py:226,238, gradients averaged by DDP).  These helpers are device-agnostic torch.distributed code
(NCCL on the GPUs; the CPU tests drive them over gloo).
"""
import torch
import torch.distributed as dist

from . import _lib


def image_shard(num_images, world_size, rank):
    """Contiguous block of images owned by ``rank`` (equal blocks; num_images must divide evenly, as
    IMS_PER_BATCH // world_size does in the reference)."""
    if num_images % world_size != 0:
        raise ValueError("batch of %d images does not split evenly over %d ranks" % (num_images, world_size))
    per = num_images // world_size
    return slice(rank * per, (rank + 1) * per)


def all_reduce_stats(stats, group=None):
    """stats = [num_foreground, S_batch, S[n]...] (double).  Sums the two global entries over ranks in
    place; the per-image normalisers S[n] stay local."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats[:_lib.STATS_HEADER], op=dist.ReduceOp.SUM, group=group)
    return stats


def all_reduce_batch_weighted_sum(scalars, group=None):
    """L_BAHW_extendtobatch only: scalars[2] = sum_n A[n] must be global before the post pass."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(scalars[2:3], op=dist.ReduceOp.SUM, group=group)
    return scalars


def global_losses(scalars, stats, coeffs, group=None, batch_sum_is_global=False):
    """Whole-batch loss values from the per-rank sums: returns a (4,) double tensor
    [loss_cls, loss_box_reg, gambler_loss, total].  ``stats[0]`` must already be global.
    ``batch_sum_is_global``: scalars[2] was already all-reduced (``all_reduce_batch_weighted_sum``, the
    L_BAHW_extendtobatch exchange) and must not be summed over ranks a second time."""
    sums = scalars[:5].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        keep = sums[2].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        if batch_sum_is_global:
            sums[2] = keep
    nf = torch.clamp(stats[0], min=1.0)
    loss_cls = sums[0] / nf
    loss_reg = sums[1] / nf
    gam = -sums[2]
    total = coeffs[0] * loss_cls + coeffs[1] * loss_reg + coeffs[2] * gam
    return torch.stack((loss_cls, loss_reg, gam, total))


class PeerExchange:
    """NVLink peer-memory mailboxes for the in-kernel all-reduce of [num_foreground, S_batch]
    (``struct fsg_peer_ctx`` in include/fsg_dense.h; kernel side: the tail of K1's second kernel).

    The 1 KiB mailbox of every rank lives in symmetric memory (``torch.distributed._symmetric_memory``: CUDA
    VMM allocations exchanged between the processes of one NVSwitch box and mapped into each of them), so a
    rank's kernel can store its two partial sums plus a release flag straight into every peer and spin on its
    own mailbox.  Replaces an NCCL launch (~15 us of host + device latency for 16 bytes) by ~3 us inside a
    kernel that is already running, and keeps the whole step inside one CUDA graph.

    ``PeerExchange.create(group, device)`` returns None when symmetric memory cannot be set up (no P2P,
    container without the needed handle passing): callers then keep the NCCL exchange.
    """

    MAILBOX_BYTES = 1024

    def __init__(self, handle, mailbox, epoch, error, rank, world, timeout_s=60.0):
        from . import _lib

        self._handle, self.mailbox, self.epoch, self.error = handle, mailbox, epoch, error
        self.rank, self.world = rank, world
        ctx = _lib.PeerCtx()
        # SM cycles a kernel waits for a peer before it raises `error` and poisons the exchanged sums with NaN
        khz = torch.cuda.get_device_properties(mailbox.device).clock_rate if hasattr(
            torch.cuda.get_device_properties(mailbox.device), "clock_rate") else 1965000
        ctx.timeout_cycles = int(float(timeout_s) * khz * 1e3)
        ptrs = list(handle.buffer_ptrs)
        for p in range(world):
            ctx.mailbox[p] = int(ptrs[p])
        ctx.epoch = epoch.data_ptr()
        ctx.error = error.data_ptr()
        ctx.rank, ctx.world = rank, world
        self.ctx = ctx

    @staticmethod
    def create(group, device, timeout_s=60.0):
        """timeout_s: how long a kernel waits for a peer (a rank stalled in its data loader, a checkpoint on rank 0,
        first-step autotuning ...) before the step is declared dead: the error flag is raised and the exchanged
        sums become NaN, so the step's losses and gradients are NaN and ``check()`` raises."""
        try:
            import torch.distributed._symmetric_memory as symm_mem

            world = dist.get_world_size(group)
            if world < 2 or world > 8:
                return None
            buf = symm_mem.empty(PeerExchange.MAILBOX_BYTES // 8, dtype=torch.float64, device=device)
            buf.zero_()
            handle = symm_mem.rendezvous(buf, group)
            epoch = torch.zeros(1, dtype=torch.int64, device=device)
            error = torch.zeros(1, dtype=torch.int32, device=device)
            torch.cuda.synchronize(device)
            dist.barrier(group)     # every mailbox is zeroed before anyone may write into it
            return PeerExchange(handle, buf, epoch, error, dist.get_rank(group), world, timeout_s)
        except Exception as e:  # noqa: BLE001 -- any failure means "not available here"
            import sys

            print("PeerExchange unavailable: %s: %s" % (type(e).__name__, str(e)[:200]), file=sys.stderr)
            return None

    def check(self):
        """Host-synchronising: raise if a peer failed to arrive within the kernel's time-out."""
        if int(self.error.item()) != 0:
            raise RuntimeError("peer exchange timed out: a rank did not enqueue the matching step")


def all_reduce_gt_max(gt_max_bits, group=None):
    """All-reduce(MAX) of the per-GT maxima between the two matching passes when ONE image's anchors are sharded by
    range over the ranks.  ``gt_max_bits``: int32 view of the fp32 bit patterns; IoU values
    are non-negative, so their bit patterns order like integers and MAX on the view is MAX on the floats."""
    dist.all_reduce(gt_max_bits, op=dist.ReduceOp.MAX, group=group)
    return gt_max_bits


def anchor_range(num_anchors, world_size, rank):
    """Contiguous anchor range [lo, hi) of ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(num_anchors, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def match_anchor_range(local_anchors, gt, num_classes, group=None, want=("matches", "match_labels"), **kw):
    """Matcher (with low-quality matches) for images whose anchors are sharded by range: every rank holds the same
    GT and its own slice ``local_anchors`` ((R_local,4) or (N,R_local,4)).  Pass A locally, all-reduce(MAX) of the
    M per-GT maxima, pass B locally; the concatenation of the ranks' outputs equals the unsharded result."""
    from . import ops
    first = ops.match_anchors(local_anchors, gt, num_classes, want=want, phases=1, **kw)
    all_reduce_gt_max(first["gt_max_bits"], group)
    return ops.match_anchors(local_anchors, gt, num_classes, want=want, phases=2, workspace=first["workspace"],
                             out={k: first[k] for k in want}, **kw)
