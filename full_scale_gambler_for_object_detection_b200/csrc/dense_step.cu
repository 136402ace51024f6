// The fused training step as one entry point (include/fsg_dense.h: fsg_dense_step).
//
// What the reference does across RetinaNet.get_ground_truth / get_picky_ground_truth / losses
// (detectron2/modeling/meta_arch/retinanet.py:201-248, 309-429), LayeredUnetGambler.gambler_loss
// (ImbalanceDetection/imbalancedetection/gambler_heads.py:502-602) and loss.backward() becomes
//
//   memset (every completion counter and the per-GT maxima of the step)
//   K1 A     per-anchor best / argmax with warp-level GT culling, per-GT maxima, and everything that depends only
//            on the anchor's own best IoU: labels, gt_classes, mask, partial pre-pass sums
//   K1 B     low-quality rule as a patch pass, sums folded; the peer sums are only posted       (PDL)
//   K2 main  polls the peer mailboxes itself, in every CTA                                      (PDL)
//   K2 post                                                                                     (PDL)
// PDL = programmatic dependent launch: the kernel is scheduled while its predecessor drains and waits on the device.
//
// Nothing here computes: the kernels live in iou_match.cu and dense_loss.cu.
#include <stdlib.h>

#include "common.cuh"
#include "step_internal.cuh"

using namespace fsg;

namespace {
// FSG_STEP_NO_PDL=1: plain stream-ordered launches (A/B timing of the programmatic dependent launch)
bool step_no_pdl() {
  static const int v = [] {
    const char* e = getenv("FSG_STEP_NO_PDL");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v != 0;
}
}  // namespace

extern "C" size_t fsg_dense_step_workspace_bytes(int N, int64_t R, int K, int64_t sum_M) {
  if (N <= 0 || R <= 0 || K <= 0 || sum_M < 0) return 0;
  return align_up(fsg_match_workspace_bytes(N, R, sum_M), 256) + loss_main_ws_bytes(N, R, K);
}

extern "C" int fsg_dense_step(const fsg_step_io* io, int N, int64_t R, const fsg_match_config* mc,
                              const fsg_loss_params* hp, const fsg_peer_ctx* h_peer, void* workspace,
                              size_t workspace_bytes, fsg_stream_t stream) {
  if (!io || !mc || !hp || N <= 0 || R <= 0) return FSG_ERR_INVALID_ARG;
  if (!io->logits || !io->pred_deltas || !io->bets || !io->anchors || !io->gt_offsets || !io->gt_classes ||
      !io->mask || !io->matched_idx32 || !io->stats || !io->scalars || !io->grad_bets || !io->per_anchor_loss)
    return FSG_ERR_INVALID_ARG;
  if (mc->num_thresholds < 1 || mc->num_thresholds > 4 || mc->num_picky_thresholds < 1 || mc->num_picky_thresholds > 4)
    return FSG_ERR_INVALID_ARG;
  const bool sharded = h_peer && h_peer->world > 1;
  if (sharded && hp->norm_mode == FSG_NORM_BATCH) return FSG_ERR_UNSUPPORTED;
  const int K = hp->num_classes;
  // K1's workspace in front, the loss kernel's at the very END of the buffer: its position then depends on the buffer
  // and not on this call's GT count, which a workspace that is kept clean across calls (workspace_is_clean) needs
  const size_t match_bytes = fsg_match_workspace_bytes(N, R, io->sum_M);
  const size_t loss_bytes = loss_main_ws_bytes(N, R, K);
  if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < align_up(match_bytes, 256) + loss_bytes)
    return FSG_ERR_WORKSPACE;
  const size_t off_loss = (workspace_bytes - loss_bytes) & ~(size_t)255;
  const size_t need = off_loss + loss_bytes;
  char* ws = (char*)workspace;
  // ONE memset node in front of K1 clears the zero-initialised tail of K1's workspace and, with it, the loss
  // kernel's completion counter at the head of the workspace that follows
  const size_t zero_tail = off_loss - match_bytes + 16;
  int st = match_enqueue(io->anchors, R, io->anchor_image_stride, io->gt_boxes, io->gt_class_ids, io->gt_offsets, N,
                         io->sum_M, K, mc->thresholds, mc->labels, mc->num_thresholds,
                         mc->allow_low_quality_matches, mc->picky_thresholds, mc->picky_labels,
                         mc->num_picky_thresholds, hp->box_weights, nullptr, nullptr, nullptr, io->gt_classes, io->mask,
                         nullptr, io->matched_idx32, io->bets, nullptr, hp->temperature, io->stats,
                         sharded ? h_peer : nullptr, ws, off_loss, 3,
                         kMatchPeerPolled | (step_no_pdl() ? 0 : kMatchPdl) | (mc->workspace_is_clean ? kMatchSelfClean : 0), zero_tail,
                         stream);
  if (st != FSG_OK) return st;
  const int pdl = step_no_pdl() ? 0 : kLossPdl;
  st = loss_main_enqueue(io->logits, io->pred_deltas, nullptr, io->anchors, io->anchor_image_stride, io->gt_boxes,
                         io->gt_offsets, io->matched_idx32, io->gt_classes, io->mask, io->bets, N, R, hp, io->stats,
                         io->grad_logits, io->grad_deltas, io->per_anchor_loss, io->weights_out, io->scalars,
                         ws + off_loss, need - off_loss, sharded ? h_peer : nullptr, pdl | kLossCounterZeroed, stream);
  if (st != FSG_OK) return st;
  return loss_post_enqueue(io->bets, io->mask, io->per_anchor_loss, N, R, hp, io->stats, io->scalars, io->grad_bets,
                           pdl, stream);
}

extern "C" size_t fsg_dense_step_levels_workspace_bytes(int N, const fsg_head_level* h_levels, int num_levels, int A,
                                                        int64_t sum_M) {
  if (N <= 0 || !h_levels || num_levels <= 0 || num_levels > FSG_MAX_LEVELS || A <= 0 || sum_M < 0) return 0;
  int64_t R = 0;
  for (int l = 0; l < num_levels; ++l) R += (int64_t)h_levels[l].H * h_levels[l].W * A;
  const size_t lw = loss_main_levels_ws_bytes(N, h_levels, num_levels, A);
  if (R <= 0 || lw == 0) return 0;
  return align_up(fsg_match_workspace_bytes(N, R, sum_M), 256) + lw;
}

extern "C" int fsg_dense_step_levels(const fsg_step_levels_io* io, const fsg_head_level* h_levels,
                                     const fsg_post_level* h_post, int num_levels, int A, int N, int64_t R,
                                     const fsg_match_config* mc, const fsg_loss_params* hp,
                                     const fsg_peer_ctx* h_peer, void* workspace, size_t workspace_bytes,
                                     fsg_stream_t stream) {
  if (!io || !h_levels || !h_post || !mc || !hp || N <= 0 || R <= 0 || num_levels <= 0 || num_levels > FSG_MAX_LEVELS)
    return FSG_ERR_INVALID_ARG;
  if (!io->anchors || !io->gt_offsets || !io->gt_classes || !io->mask || !io->matched_idx32 || !io->stats ||
      !io->scalars)
    return FSG_ERR_INVALID_ARG;
  if (mc->num_thresholds < 1 || mc->num_thresholds > 4 || mc->num_picky_thresholds < 1 || mc->num_picky_thresholds > 4)
    return FSG_ERR_INVALID_ARG;
  const bool sharded = h_peer && h_peer->world > 1;
  if (sharded && hp->norm_mode == FSG_NORM_BATCH) return FSG_ERR_UNSUPPORTED;
  fsg_bet_levels bl = {};
  bl.num_levels = num_levels;
  bl.A = A;
  for (int l = 0; l < num_levels; ++l) {
    if (!h_levels[l].bets) return FSG_ERR_INVALID_ARG;   // this entry point is for the all-native form
    bl.bets[l] = h_levels[l].bets;
    bl.H[l] = h_levels[l].H;
    bl.W[l] = h_levels[l].W;
  }
  const size_t match_bytes = fsg_match_workspace_bytes(N, R, io->sum_M);
  const size_t lw = loss_main_levels_ws_bytes(N, h_levels, num_levels, A);
  if (lw == 0) return FSG_ERR_INVALID_ARG;
  if (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < align_up(match_bytes, 256) + lw)
    return FSG_ERR_WORKSPACE;
  const size_t off_loss = (workspace_bytes - lw) & ~(size_t)255;   // (see fsg_dense_step)
  char* ws = (char*)workspace;
  const size_t zero_tail = off_loss - match_bytes + 16;
  const int pdl = step_no_pdl() ? 0 : kLossPdl;
  int st = match_enqueue(io->anchors, R, io->anchor_image_stride, io->gt_boxes, io->gt_class_ids, io->gt_offsets, N,
                         io->sum_M, hp->num_classes, mc->thresholds, mc->labels, mc->num_thresholds,
                         mc->allow_low_quality_matches, mc->picky_thresholds, mc->picky_labels,
                         mc->num_picky_thresholds, hp->box_weights, nullptr, nullptr, nullptr, io->gt_classes, io->mask,
                         nullptr, io->matched_idx32, nullptr, &bl, hp->temperature, io->stats,
                         sharded ? h_peer : nullptr, ws, off_loss, 3,
                         kMatchPeerPolled | (step_no_pdl() ? 0 : kMatchPdl) | (mc->workspace_is_clean ? kMatchSelfClean : 0), zero_tail,
                         stream);
  if (st != FSG_OK) return st;
  st = loss_main_levels_enqueue(h_levels, num_levels, A, nullptr, io->anchors, io->anchor_image_stride, io->gt_boxes,
                                io->gt_offsets, io->matched_idx32, io->gt_classes, io->mask, nullptr, N, R, hp,
                                io->stats, nullptr, io->weights_out, io->scalars, ws + off_loss, lw,
                                sharded ? h_peer : nullptr, pdl | kLossCounterZeroed, stream);
  if (st != FSG_OK) return st;
  return loss_post_levels_enqueue(h_post, num_levels, A, io->mask, N, R, hp, io->stats, io->scalars, pdl, stream);
}
