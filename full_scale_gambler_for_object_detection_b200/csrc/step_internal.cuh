// Internal (not part of the C ABI): launchers shared between translation units so that the fused training step
// (fsg_dense_step, dense_step.cu) can enqueue K1 and K2 back to back under programmatic dependent launch.
#pragma once
#include "common.cuh"

namespace fsg {

constexpr int kMatchPeerPolled = 1;   // match_enqueue flag: post the peer exchange, the consumer kernel polls it
constexpr int kMatchSelfClean = 4;    // no memset: the zero-initialised part of the workspace is clean and is left clean
constexpr int kMatchPdl = 2;          // launch pass B and the fold under programmatic dependent launch

int match_enqueue(const float* anchors, int64_t R, int64_t anchor_image_stride, const float* gt_boxes,
                  const int64_t* gt_class_ids, const int32_t* gt_offsets, int N, int64_t sum_M, int num_classes,
                  const float* h_thresholds, const int8_t* h_labels, int num_thresholds, int allow_lq,
                  const float* h_picky_thresholds, const int8_t* h_picky_labels, int num_picky_thresholds,
                  const float* h_box_weights, int64_t* matches, int8_t* match_labels, int8_t* picky_labels,
                  int64_t* gt_classes_out, int64_t* mask_out, float* gt_deltas, int32_t* matched_idx32,
                  const float* bets, const fsg_bet_levels* h_bet_levels, float temperature, double* stats,
                  const fsg_peer_ctx* h_peer, void* workspace, size_t workspace_bytes, int phases, int flags,
                  size_t zero_tail_bytes, fsg_stream_t stream);
// zero_tail_bytes: bytes right behind fsg_match_workspace_bytes() that the memset in front of pass A clears as well

constexpr int kLossPdl = 1;            // launch under programmatic dependent launch (the kernel waits on-device)
constexpr int kLossCounterZeroed = 2;  // the caller already zeroed the first 16 bytes of the loss workspace

// stats is written (entries 0..1) only when h_peer is given: the kernel then polls the peer mailboxes itself
int loss_main_enqueue(const float* logits, const float* pred_deltas, const float* gt_deltas, const float* anchors,
                      int64_t anchor_image_stride, const float* gt_boxes, const int32_t* gt_offsets,
                      const int32_t* matched_idx32, const int64_t* gt_classes, const int64_t* mask, const float* bets,
                      int N, int64_t R, const fsg_loss_params* hp, double* stats, float* grad_logits,
                      float* grad_deltas, float* per_anchor_loss, float* weights_out, double* scalars, void* workspace,
                      size_t workspace_bytes, const fsg_peer_ctx* h_peer, int flags, fsg_stream_t stream);
int loss_post_enqueue(const float* bets, const int64_t* mask, const float* per_anchor_loss, int N, int64_t R,
                      const fsg_loss_params* hp, const double* stats, const double* scalars, float* grad_bets,
                      int flags, fsg_stream_t stream);
size_t loss_main_ws_bytes(int N, int64_t R, int K);

// the same on the head's native per-level layout (dense_loss_levels.cu)
int loss_main_levels_enqueue(const fsg_head_level* h_levels, int num_levels, int A, const float* gt_deltas,
                             const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                             const int32_t* gt_offsets, const int32_t* matched_idx32, const int64_t* gt_classes,
                             const int64_t* mask, const float* bets, int N, int64_t R, const fsg_loss_params* hp,
                             double* stats, float* per_anchor_loss, float* weights_out, double* scalars,
                             void* workspace, size_t workspace_bytes, const fsg_peer_ctx* h_peer, int flags,
                             fsg_stream_t stream);
int loss_post_levels_enqueue(const fsg_post_level* h_levels, int num_levels, int A, const int64_t* mask, int N,
                             int64_t R, const fsg_loss_params* hp, const double* stats, const double* scalars,
                             int flags, fsg_stream_t stream);
size_t loss_main_levels_ws_bytes(int N, const fsg_head_level* h_levels, int num_levels, int A);

}  // namespace fsg
