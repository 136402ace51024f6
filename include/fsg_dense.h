/*
 * fsg_dense.h -- C ABI of the B200-native (sm_100a) per-anchor dense-detection hot path.
 *
 * Drop-in boundary for the calls the reference makes on this path (paths relative to the
 * reference tree):
 *   detectron2/structures/boxes.py:243-275        pairwise_iou
 *   detectron2/modeling/matcher.py:55-132         Matcher.__call__ / set_low_quality_matches_
 *   detectron2/modeling/box_regression.py:34-107  Box2BoxTransform.get_deltas / apply_deltas
 *   detectron2/modeling/meta_arch/retinanet.py:201-248,309-429,460-520
 *                                                 RetinaNet.losses / get_ground_truth /
 *                                                 get_picky_ground_truth / inference_single_image
 *   ImbalanceDetection/imbalancedetection/gambler_heads.py:104-128,131-253,291-318,502-602
 *                                                 calc_cls_loss / calc_gambler_loss / gambler_loss
 *   detectron2/layers/nms.py:6,9-26               nms / batched_nms (torchvision semantics)
 *
 * The reference's own native FFI convention (detectron2/layers/csrc/vision.cpp:58-95: pybind11
 * functions taking at::Tensor, CUDA stream from at::cuda::getCurrentCUDAStream(), errors through
 * AT_ASSERTM) is replaced by this torch-free ABI: raw device pointers, explicit sizes, scalar
 * hyper-parameters, a cudaStream_t, and an int status (0 = ok).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_` (host, read at call time);
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never allocates,
 *     never synchronises the device, never changes the current device, keeps no mutable global
 *     state; every call only enqueues work on `stream`;
 *   - all float data is fp32, boxes are XYXY, class ids / match indices are int64 at the boundary
 *     (the reference's dtypes), match labels are int8;
 *   - "anchor index" r runs over the reference's (N, sum_l H_l*W_l*A, K) flattening
 *     (retinanet.py:24-54): r = level_offset + (h*W + w)*A + a.
 */
#ifndef FSG_DENSE_H_
#define FSG_DENSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSG_ABI_VERSION 4

#if defined(__GNUC__)
#define FSG_API __attribute__((visibility("default")))
#else
#define FSG_API
#endif

typedef void* fsg_stream_t; /* cudaStream_t */

enum fsg_status {
  FSG_OK = 0,
  FSG_ERR_INVALID_ARG = 1,  /* bad size / null pointer / unsupported combination      */
  FSG_ERR_WORKSPACE = 2,    /* workspace too small (see the *_workspace_bytes queries) */
  FSG_ERR_UNSUPPORTED = 3,  /* shape outside what the kernels cover                    */
  FSG_ERR_CUDA = 1000       /* FSG_ERR_CUDA + cudaError_t of the failed launch         */
};

FSG_API int fsg_abi_version(void);
/* static string for a status returned by any entry point */
FSG_API const char* fsg_status_string(int status);

/* ------------------------------------------------------------------------------------------
 * K1 -- IoU and matching
 * ---------------------------------------------------------------------------------------- */

/* boxes.py:243-275.  iou[i*n2 + j] = IoU(boxes1[i], boxes2[j]); exactly 0 where inter <= 0.
 * Bit-exact with the reference's fp32 op sequence (no FMA contraction, IEEE division). */
FSG_API int fsg_pairwise_iou(const float* boxes1, int64_t n1, const float* boxes2, int64_t n2, float* iou,
                     fsg_stream_t stream);

/* matcher.py:55-132 on a materialised (M, N) quality matrix (row-major).
 * h_thresholds[num_thresholds] ascending, h_labels[num_thresholds + 1] in {-1,0,1}.
 * M == 0: matches = 0, match_labels = h_labels[0] (matcher.py:70-80).
 * ws_rowmax: workspace of M floats. */
FSG_API int fsg_matcher(const float* mqm, int64_t M, int64_t N, const float* h_thresholds,
                const int8_t* h_labels, int num_thresholds, int allow_low_quality_matches,
                int64_t* matches, int8_t* match_labels, float* ws_rowmax, fsg_stream_t stream);

/* Peer-memory exchange context for a batch sharded by image over the GPUs of one NVLink/NVSwitch box.
 * mailbox[p] is rank p's mailbox (1 KiB of symmetric, zero-initialised device memory; the upper 512 bytes are
 * scratch of the owning rank: the first CTA of the loss kernel that has all peers' sums publishes them there) as mapped in THIS
 * process (its own mailbox included).  When a context is passed to fsg_match_anchors, the last CTA of its
 * second kernel all-reduces stats[0..1] = [num_foreground, S_batch] itself: it stores its two partial sums
 * and a release flag (a per-launch epoch) into every peer's mailbox with plain st.global over NVLink, spins
 * (acquire loads, with a clock time-out that raises *error) until every peer's flag shows the same epoch,
 * and sums the slots in rank order (identical result on every rank).  No NCCL launch, capturable in a CUDA
 * graph; every rank must enqueue the same sequence of calls.
 * fsg_dense_step / fsg_dense_step_levels shorten that to what the loss kernel really has to wait for: K1's fold
 * kernel posts one 8-byte word per peer (epoch << 32 | num_foreground, an atomic store: no payload / fence / flag
 * sequence), the first CTAs of the loss kernel poll those words, and the complete records for stats[0..1] are posted
 * by one CTA of the loss kernel and read by its last one. */
typedef struct fsg_peer_ctx {
  uint64_t mailbox[8]; /* device pointers */
  uint64_t epoch;      /* device pointer to a local uint64 counter, zero-initialised                */
  uint64_t error;      /* device pointer to a local int32 flag, set to 1 if a peer never showed up  */
  int32_t rank, world; /* world <= 8 */
  uint64_t timeout_cycles; /* SM clock cycles a kernel waits for a peer; 0 = default (about 60 s).  On time-out
                              *error is raised AND the exchanged sums become NaN, so the step's losses and
                              gradients are NaN: a step with a missing peer can never be used silently. */
} fsg_peer_ctx;

/* The gambler's betting maps in their own layout (level l: (N, A, H_l, W_l), gambler_heads.py:463-470), for the
 * entry points that can read them in place instead of a flattened (N, R) copy. */
#define FSG_MAX_LEVELS 8
typedef struct fsg_bet_levels {
  const float* bets[FSG_MAX_LEVELS];
  int32_t H[FSG_MAX_LEVELS], W[FSG_MAX_LEVELS];
  int32_t num_levels, A;
} fsg_bet_levels;

/* Fused IoU + Matcher(s) + GT assignment for a batch of images, never materialising the
 * (M, R) matrix.  Replaces, per image, retinanet.py:339-363 and :400-425:
 *   pairwise_iou -> Matcher(thresholds) [-> picky Matcher(picky_thresholds)] ->
 *   gt_classes relabel -> get_deltas -> picky mask.
 *
 * anchors            (R,4) shared by all images when anchor_image_stride == 0, else image n's
 *                    anchors start at anchors + n*anchor_image_stride (in floats).
 * gt_boxes           packed (sum_M, 4); image n owns rows gt_offsets[n] .. gt_offsets[n+1]-1
 * gt_class_ids       packed (sum_M) int64; may be NULL when gt_classes_out == NULL
 * gt_offsets         (N+1) int32, device
 * allow_low_quality_matches applies to both matchers (RetinaNet: 1, retinanet.py:91-100)
 * h_picky_thresholds NULL -> no second matcher (mask_out must be NULL)
 * outputs (each may be NULL, shapes (N,R) unless noted):
 *   matches int64, match_labels int8 (first matcher), picky_labels int8,
 *   gt_classes_out int64 in {-1, 0..K-1, K}  (K = num_classes; image without GT -> all K),
 *   mask_out int64 (1 iff picky label == 1; image without GT -> all K, retinanet.py:425),
 *   gt_deltas (N,R,4) fp32 (zeros for an image without GT),
 *   matched_idx32 (N,R) int32 copy of matches for the loss kernel.
 * Optional fused loss pre-pass (bets != NULL, or h_bet_levels != NULL for per-level betting maps read in place --
 * not both): stats as defined by fsg_loss_prepass.
 * workspace: fsg_match_workspace_bytes(N, R, sum_M) bytes, 16-byte aligned. */
FSG_API size_t fsg_match_workspace_bytes(int N, int64_t R, int64_t sum_M);
FSG_API int fsg_match_anchors(const float* anchors, int64_t R, int64_t anchor_image_stride,
                      const float* gt_boxes, const int64_t* gt_class_ids, const int32_t* gt_offsets,
                      int N, int64_t sum_M, int num_classes, const float* h_thresholds,
                      const int8_t* h_labels, int num_thresholds, int allow_low_quality_matches,
                      const float* h_picky_thresholds, const int8_t* h_picky_labels,
                      int num_picky_thresholds, const float* h_box_weights /* 4 */, int64_t* matches, int8_t* match_labels,
                      int8_t* picky_labels, int64_t* gt_classes_out, int64_t* mask_out,
                      float* gt_deltas, int32_t* matched_idx32, const float* bets,
                      const fsg_bet_levels* h_bet_levels /* NULL: flat bets or none */,
                      float temperature, double* stats, const fsg_peer_ctx* h_peer /* NULL: no exchange */,
                      void* workspace, size_t workspace_bytes, fsg_stream_t stream);

/* The same in two phases, for ONE image (or a few) whose ANCHORS are sharded by range over the GPUs (SURVEY section
 * 8e, the matcher stress with fewer images than GPUs): phases = 1 runs pass A on the local anchors (per-anchor
 * best/argmax and the per-GT maxima over the local anchors), phases = 2 runs pass B with whatever per-GT maxima are
 * in the workspace, phases = 3 is fsg_match_anchors.  Between the two calls the caller all-reduces (MAX) the
 * sum_M uint32 words at workspace + fsg_match_gt_max_offset(N, R, sum_M) over the ranks: they hold the per-GT
 * maxima as fp32 bit patterns, which order like unsigned integers for the non-negative IoU values.  The same
 * workspace must be passed to both calls; stats, if requested, are partial sums over the local anchors. */
FSG_API int fsg_match_anchors_ex(const float* anchors, int64_t R, int64_t anchor_image_stride,
                         const float* gt_boxes, const int64_t* gt_class_ids, const int32_t* gt_offsets,
                         int N, int64_t sum_M, int num_classes, const float* h_thresholds,
                         const int8_t* h_labels, int num_thresholds, int allow_low_quality_matches,
                         const float* h_picky_thresholds, const int8_t* h_picky_labels,
                         int num_picky_thresholds, const float* h_box_weights, int64_t* matches,
                         int8_t* match_labels, int8_t* picky_labels, int64_t* gt_classes_out, int64_t* mask_out,
                         float* gt_deltas, int32_t* matched_idx32, const float* bets,
                         const fsg_bet_levels* h_bet_levels, float temperature, double* stats,
                         const fsg_peer_ctx* h_peer, void* workspace, size_t workspace_bytes, int phases,
                         fsg_stream_t stream);
FSG_API size_t fsg_match_gt_max_offset(int N, int64_t R, int64_t sum_M);

/* box_regression.py:34-67 / :69-107.  h_weights = (wx, wy, ww, wh).
 * apply_deltas: deltas (n, 4k) -> out (n, 4k), dw/dh clamped to scale_clamp (max only). */
FSG_API int fsg_box2box_get_deltas(const float* src_boxes, const float* target_boxes, int64_t n,
                           const float* h_weights, float* deltas, fsg_stream_t stream);
FSG_API int fsg_box2box_apply_deltas(const float* deltas, const float* boxes, int64_t n, int k,
                             const float* h_weights, float scale_clamp, float* out,
                             fsg_stream_t stream);

/* DefaultAnchorGenerator.grid_anchors (anchor_generator.py:41-50,121-129) on the device, one launch for all
 * levels: anchors[r] for r = level_offset + (y*W + x)*A + a is (x*stride, y*stride, x*stride, y*stride) +
 * cell[a] in fp32 (bit-exact with the reference).  cell = generate_cell_anchors(sizes, aspect_ratios)
 * (:131-168, host arithmetic in double, rounded to fp32).  R must equal sum_l H*W*A. */
#define FSG_MAX_CELL_ANCHORS 16
typedef struct fsg_anchor_level {
  int32_t H, W, stride, A;
  float cell[FSG_MAX_CELL_ANCHORS][4];
} fsg_anchor_level;
FSG_API int fsg_grid_anchors(const fsg_anchor_level* h_levels, int num_levels, float* anchors, int64_t R,
                     fsg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Head-layout adapter (retinanet.py:24-54, gambler_heads.py:34-101)
 * ---------------------------------------------------------------------------------------- */

/* One FPN level: nchw (N, C, HW) with C = A*K  <->  flat[n*flat_image_stride + flat_offset + hw*C + c],
 * i.e. rows (h*W+w)*A+a of the (N, sum HWA, K) layout when flat_offset = level_offset*K and
 * flat_image_stride = R*K.  to_nchw == 0 reads nchw and writes flat; 1 is the inverse (gradients). */
FSG_API int fsg_permute_level(float* nchw, float* flat, int N, int C, int64_t HW,
                              int64_t flat_image_stride, int64_t flat_offset, int to_nchw,
                              fsg_stream_t stream);

/* The (N,R)-sized per-anchor maps in one launch for all levels and up to three tensors at once:
 * per-level (N, A, H_l, W_l) maps (the gambler's betting maps, NAKHW_loss, d/d bets;
 * gambler_heads.py:91-101,291-318)  <->  flat (N, R) with r = level_offset + (h*W+w)*A + a.
 * h_level_ptrs[t*num_levels + l]: tensor t, level l; h_flat_ptrs[t]: tensor t flat; h_HW[l] = H_l*W_l.
 * to_levels == 0 gathers levels -> flat, 1 scatters flat -> levels. */
FSG_API int fsg_anchor_maps(float* const* h_level_ptrs, float* const* h_flat_ptrs, int ntensors,
                    const int32_t* h_HW, int num_levels, int A, int N, int to_levels, fsg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2 -- fused gambler-weighted sigmoid-focal / smooth-L1 loss, forward + backward
 * ---------------------------------------------------------------------------------------- */

enum fsg_cls_mode { FSG_CLS_FOCAL = 0, FSG_CLS_SIGMOID = 1 };         /* GAMBLER_LOSS_MODE      */
enum fsg_norm_mode { FSG_NORM_NONE = 0, FSG_NORM_IMAGE = 1, FSG_NORM_BATCH = 2 };
/* NORMALIZE=False / L_BAHW (gambler_heads.py:311) / L_BAHW_extendtobatch (:308-309) */

typedef struct fsg_loss_params {
  int32_t num_classes;      /* K                                                          */
  int32_t gambler_mode;     /* fsg_cls_mode for the gambler term (loss_cls is always focal) */
  int32_t norm_mode;        /* fsg_norm_mode                                               */
  int32_t reserved;
  float focal_alpha;        /* < 0: no alpha weighting                                     */
  float focal_gamma;        /* 2.0 takes the fast path                                     */
  float smooth_l1_beta;     /* retinanet.py:241-246                                        */
  float temperature;        /* GAMBLER_TEMPERATURE (added to bet*mask)                     */
  float gambler_gamma;      /* GAMBLER_GAMMA: G = -sum w_hat^gamma * l                     */
  float c_cls, c_reg, c_gam; /* the backward is that of c_cls*loss_cls + c_reg*loss_box_reg
                                + c_gam*gambler_loss (train_net.py:1089-1098)              */
  float box_weights[4];     /* Box2BoxTransform weights for the fused encode               */
} fsg_loss_params;

/* stats layout (double): [0] num_foreground  [1] S_batch = sum_n S[n]  [2+n] S[n] = sum_r (bet*mask+T)
 * In a sharded (multi-GPU) run, all-reduce(SUM) stats[0..1] over ranks between the pre-pass and
 * fsg_loss_main; S[n] stays local (per-image normaliser). */
#define FSG_STATS_HEADER 2
FSG_API size_t fsg_loss_prepass_workspace_bytes(int N, int64_t R);
FSG_API int fsg_loss_prepass(const int64_t* gt_classes, const int64_t* mask, const float* bets, int N,
                     int64_t R, int num_classes, float temperature, double* stats, void* workspace,
                     size_t workspace_bytes, fsg_stream_t stream);

/* scalars layout (double), written by fsg_loss_main:
 *  [0] sum_valid focal          (loss_cls * max(1, nf))
 *  [1] sum_fg smooth-L1         (loss_box_reg * max(1, nf))
 *  [2] sum_n A[n]               (= -gambler_loss), A[n] = sum_r w_hat^gamma * l
 *  [3] sum l                    (loss_before_weighting numerator)
 *  [4] sum_n max_r l[n,r]       (get_loss_upper_bound, gambler_heads.py:17-31)
 *  [5] loss_cls  [6] loss_box_reg  [7] gambler_loss   (local-rank values, normalised with stats[0])
 *  [8] c_cls*[5] + c_reg*[6] + c_gam*[7]   [9] num_foreground used
 *  [10+n] A[n]
 * In a sharded run with FSG_NORM_BATCH, all-reduce(SUM) scalars[2] before fsg_loss_post. */
#define FSG_SCALARS_HEADER 10
FSG_API size_t fsg_loss_main_workspace_bytes(int N, int64_t R, int K);

/* The main pass.  Reads logits (N,R,K) once and writes grad_logits once.
 * Regression targets: either gt_deltas (N,R,4) explicit, or (gt_deltas == NULL) encoded on the fly
 * from anchors + gt_boxes[gt_offsets[n] + matched_idx32[n,r]] (Box2BoxTransform.get_deltas fused).
 * bets (N,R) may be NULL when c_gam == 0 and no gambler outputs are requested.
 * Outputs (each may be NULL): grad_logits (N,R,K), grad_deltas (N,R,4), per_anchor_loss l (N,R)
 * (the NAKHW_loss values, gambler_heads.py:218), weights_out w_hat (N,R). */
FSG_API int fsg_loss_main(const float* logits, const float* pred_deltas, const float* gt_deltas,
                  const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                  const int32_t* gt_offsets, const int32_t* matched_idx32,
                  const int64_t* gt_classes, const int64_t* mask, const float* bets, int N,
                  int64_t R, const fsg_loss_params* h_params, const double* stats,
                  float* grad_logits, float* grad_deltas, float* per_anchor_loss,
                  float* weights_out, double* scalars, void* workspace, size_t workspace_bytes,
                  fsg_stream_t stream);

/* The main pass on the head's NATIVE layout: per-level conv outputs are read in place
 * and the gradients are written in the same layout, so permute_to_N_HWA_K + cat (retinanet.py:24-54,217-219;
 * gambler_heads.py:34-101,538-540) and their inverses in the backward are never materialised.
 * Level l: logits (N, A*K, H_l, W_l), pred_deltas (N, A*4, H_l, W_l) (channel a*K+k / a*4+j, retinanet.py:40-43),
 * grad_* of the same shapes (each may be NULL).  Levels are concatenated in the order given; R must equal
 * sum_l H_l*W_l*A.  Everything (N,R)-sized keeps the flattened anchor order r = level_offset + (h*W+w)*A + a:
 * gt_classes, mask, bets, matched_idx32, gt_deltas (N,R,4), per_anchor_loss, weights_out.
 * All other arguments, stats and scalars exactly as fsg_loss_main. */
typedef struct fsg_head_level {
  const float* logits;
  float* grad_logits;
  const float* pred_deltas;
  float* grad_deltas;
  const float* bets;        /* (N, A, H, W) betting map of the level, or NULL: the flat `bets` argument is used   */
  float* per_anchor_loss;   /* (N, A, H, W) NAKHW_loss of the level, or NULL: the flat `per_anchor_loss` argument */
  int32_t H, W;
} fsg_head_level;
FSG_API size_t fsg_loss_main_levels_workspace_bytes(int N, const fsg_head_level* h_levels, int num_levels, int A);
FSG_API int fsg_loss_main_levels(const fsg_head_level* h_levels, int num_levels, int A, const float* gt_deltas,
                         const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                         const int32_t* gt_offsets, const int32_t* matched_idx32, const int64_t* gt_classes,
                         const int64_t* mask, const float* bets, int N, int64_t R,
                         const fsg_loss_params* h_params, const double* stats, float* per_anchor_loss,
                         float* weights_out, double* scalars, void* workspace, size_t workspace_bytes,
                         fsg_stream_t stream);

/* d(c_gam * gambler_loss)/d bets  (SURVEY App. A item 12):
 *   -(m/S) * gamma * (w_hat^(gamma-1) * l - A)     for FSG_NORM_IMAGE / _BATCH
 *   -m * gamma * w^(gamma-1) * l                   for FSG_NORM_NONE */
FSG_API int fsg_loss_post(const float* bets, const int64_t* mask, const float* per_anchor_loss, int N,
                  int64_t R, const fsg_loss_params* h_params, const double* stats,
                  const double* scalars, float* grad_bets, fsg_stream_t stream);

/* fsg_loss_post on the gambler's own layout: betting maps, NAKHW_loss and d/d bets per level as (N, A, H, W);
 * mask stays the flat (N, R) K1 output (may be NULL). */
typedef struct fsg_post_level {
  const float* bets;
  const float* per_anchor_loss;
  float* grad_bets;
  int32_t H, W;
} fsg_post_level;
FSG_API int fsg_loss_post_levels(const fsg_post_level* h_levels, int num_levels, int A, const int64_t* mask, int N,
                         int64_t R, const fsg_loss_params* h_params, const double* stats, const double* scalars,
                         fsg_stream_t stream);

/* The whole fused training step in one call: K1 (as one persistent launch when the batch averages <= 32 GT per
 * image) -> K2 main pass -> K2 post pass, the last two under programmatic dependent launch (each kernel issues its
 * first loads while the kernel in front of it drains and waits on the device for its completion), one memset node
 * in front.  Same arithmetic and outputs as fsg_match_anchors + fsg_loss_main + fsg_loss_post on the flattened
 * layout (the regression targets are encoded on the fly).  Capturable in a CUDA graph.
 * h_peer (batch sharded by image over ranks): K1 only posts this rank's [num_foreground, S_batch] into the peers'
 * mailboxes; every CTA of the main pass polls its own mailbox and sums in rank order, so the NVLink round trip
 * hides behind the first logit loads.  stats[0..1] hold the global sums afterwards.  FSG_NORM_BATCH needs a second
 * exchange before the post pass and is not supported together with h_peer (status FSG_ERR_UNSUPPORTED). */
typedef struct fsg_step_io {
  const float* logits;      /* (N,R,K) */
  const float* pred_deltas; /* (N,R,4) */
  const float* bets;        /* (N,R)   */
  const float* anchors;     /* (R,4), or (N,R,4) with anchor_image_stride = R*4 */
  int64_t anchor_image_stride;
  const float* gt_boxes;    /* packed (sum_M,4) */
  const int64_t* gt_class_ids;
  const int32_t* gt_offsets; /* (N+1) */
  int64_t sum_M;
  /* outputs */
  int64_t* gt_classes;      /* (N,R) */
  int64_t* mask;            /* (N,R) */
  int32_t* matched_idx32;   /* (N,R) */
  double* stats;            /* [2+N] */
  double* scalars;          /* [10+N] */
  float* grad_logits;       /* (N,R,K) or NULL (gambler phase, detach_pred) */
  float* grad_deltas;       /* (N,R,4) or NULL */
  float* grad_bets;         /* (N,R) */
  float* per_anchor_loss;   /* (N,R) */
  float* weights_out;       /* (N,R) or NULL */
} fsg_step_io;
typedef struct fsg_match_config {
  float thresholds[4];
  float picky_thresholds[4];
  int8_t labels[8];         /* num_thresholds + 1 used */
  int8_t picky_labels[8];
  int32_t num_thresholds, num_picky_thresholds; /* num_picky_thresholds == 0: invalid here (the mask is needed) */
  int32_t allow_low_quality_matches;
  int32_t workspace_is_clean; /* 0: the call clears what it needs (one memset node, ~4 us of the step inside a CUDA
                               * graph).  1: the caller vouches that `workspace` was zero-filled once after allocation
                               * and since then only used by fsg_dense_step / fsg_dense_step_levels calls with the same
                               * (N, R, levels, workspace_bytes) -- the number of GT may change: every call leaves it
                               * clean again, and no memset is enqueued. */
} fsg_match_config;
FSG_API size_t fsg_dense_step_workspace_bytes(int N, int64_t R, int K, int64_t sum_M);
FSG_API int fsg_dense_step(const fsg_step_io* h_io, int N, int64_t R, const fsg_match_config* h_match,
                   const fsg_loss_params* h_params, const fsg_peer_ctx* h_peer, void* workspace,
                   size_t workspace_bytes, fsg_stream_t stream);

/* Bet / weight statistics of GANTrainer.calc_log_metrics (ImbalanceDetection/train_net.py:1104-1121) without a
 * host sync or a sort: out[0..2] = sum, max, mean of the MASKED betting maps bet*mask (the maps as gambler_loss
 * leaves them, gambler_heads.py:568-569; the max starts at 0 like the reference's running maximum), out[3..5] = sum,
 * max, mean of the normalised weights w_hat (N*R values, computed as the loss kernel computes them), out[6] =
 * torch.median(w_hat) (the lower median, by a 3-pass radix select that never materialises w_hat), out[7] reserved.
 * bets: flat (N,R), or h_bet_levels for the per-level (N, A, H, W) maps read in place -- exactly one of the two.
 * mask (N,R) int64 may be NULL (= 1).  stats: the step's [num_foreground, S_batch, S[n]...].  Three launches. */
FSG_API size_t fsg_bet_stats_workspace_bytes(void);
FSG_API int fsg_bet_stats(const float* bets, const fsg_bet_levels* h_bet_levels, const int64_t* mask, int N, int64_t R,
                  const fsg_loss_params* h_params, const double* stats, double* out /* 8 */, void* workspace,
                  size_t workspace_bytes, fsg_stream_t stream);

/* fsg_dense_step on the head's NATIVE layout: logits / deltas / betting maps and their gradients, and the NAKHW_loss,
 * all per level as the heads and the gambler produce them (h_levels: logits, grad_logits, pred_deltas, grad_deltas,
 * bets, per_anchor_loss; h_post: bets, per_anchor_loss, grad_bets).  Same kernel chain and peer handling. */
typedef struct fsg_step_levels_io {
  const float* anchors;     /* (R,4), or (N,R,4) with anchor_image_stride = R*4 */
  int64_t anchor_image_stride;
  const float* gt_boxes;
  const int64_t* gt_class_ids;
  const int32_t* gt_offsets;
  int64_t sum_M;
  int64_t* gt_classes;      /* (N,R) */
  int64_t* mask;            /* (N,R) */
  int32_t* matched_idx32;   /* (N,R) */
  double* stats;            /* [2+N] */
  double* scalars;          /* [10+N] */
  float* weights_out;       /* (N,R) or NULL */
} fsg_step_levels_io;
FSG_API size_t fsg_dense_step_levels_workspace_bytes(int N, const fsg_head_level* h_levels, int num_levels, int A,
                                             int64_t sum_M);
FSG_API int fsg_dense_step_levels(const fsg_step_levels_io* h_io, const fsg_head_level* h_levels,
                          const fsg_post_level* h_post, int num_levels, int A, int N, int64_t R,
                          const fsg_match_config* h_match, const fsg_loss_params* h_params,
                          const fsg_peer_ctx* h_peer, void* workspace, size_t workspace_bytes, fsg_stream_t stream);

/* in-place x *= *scale_dev or x *= scale_host (backward with a non-unit upstream gradient) */
FSG_API int fsg_scale_inplace(float* x, int64_t n, const float* scale_dev, float scale_host,
                      fsg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3 -- decode + score threshold + top-k + batched NMS
 * ---------------------------------------------------------------------------------------- */

/* torchvision.ops.nms semantics (layers/nms.py:6): stable score-descending greedy NMS, suppress
 * when IoU > iou_threshold (compared as the reference does, fp32 IoU against the double
 * threshold).  class_ids == NULL: plain nms; else per-class (batched_nms, layers/nms.py:9-26,
 * un-offset per-class algorithm).  keep (n) int64 receives the kept indices in score-descending
 * order (ties: lower index first); *num_keep (device int32) their count.
 * workspace: fsg_nms_workspace_bytes(n). */
FSG_API size_t fsg_nms_workspace_bytes(int64_t n);
FSG_API int fsg_nms(const float* boxes, const float* scores, const int64_t* class_ids, int64_t n,
            double iou_threshold, int64_t* keep, int32_t* num_keep, void* workspace,
            size_t workspace_bytes, fsg_stream_t stream);

/* Batched RetinaNet.inference (retinanet.py:431-520) for N images at once.
 * logits (N,R,K), deltas (N,R,4) in the flattened layout; anchors as in fsg_match_anchors;
 * h_level_offsets[num_levels+1]: anchor ranges of the FPN levels (top-k is per level, :487).
 * Outputs: out_boxes (N,max_det,4), out_scores (N,max_det), out_classes (N,max_det) int64,
 * out_count (N) int32; rows >= out_count[n] are zero-filled.
 * Optional pre-NMS candidates in the reference's concatenation order (level-major, score
 * descending inside a level): cand_boxes (N,cap,4), cand_scores (N,cap), cand_classes (N,cap)
 * int64, cand_count (N) int32, keep_idx (N,max_det) int64 with cap = num_levels*topk.
 * postprocess (device, (N,4) fp32 rows [scale_x, scale_y, clip_w, clip_h], 16-byte aligned, or NULL):
 * detector_postprocess (modeling/postprocessing.py:8-52) fused into the NMS epilogue -- the max_det survivors
 * are scaled (Boxes.scale), clipped to the output size (Boxes.clip) and the ones that became empty
 * (Boxes.nonempty) are dropped, stable; out_count / keep_idx then describe the filtered list. */
FSG_API size_t fsg_detect_workspace_bytes(int N, int64_t R, int K, int num_levels, int topk);
/* Diagnostics: after fsg_detect / fsg_detect_levels, the (N, num_levels) int32 words at workspace + this offset say
 * how each (image, level) slab was selected: 1 = sampled bar + one scan, 2 = the sampled bar could not be proven
 * safe (score ties at the top-k boundary, a plateau overflowing the candidate list, an unlucky sample) and the exact
 * streaming top-k kernel redid the slab.  The result is the same either way. */
FSG_API size_t fsg_detect_status_offset(int N, int num_levels, int topk);
FSG_API int fsg_detect(const float* logits, const float* deltas, const float* anchors,
               int64_t anchor_image_stride, int N, int64_t R, int K,
               const int64_t* h_level_offsets, int num_levels, float score_threshold, int topk,
               double nms_threshold, int max_det, const float* h_box_weights, float scale_clamp,
               float* out_boxes, float* out_scores, int64_t* out_classes, int32_t* out_count,
               float* cand_boxes, float* cand_scores, int64_t* cand_classes, int32_t* cand_count,
               int64_t* keep_idx, const float* postprocess, void* workspace, size_t workspace_bytes,
               fsg_stream_t stream);

/* fsg_detect on the head's NATIVE layout (SURVEY section 8f row 2 for K3): level l's logits (N, A*K, H_l, W_l) and
 * deltas (N, A*4, H_l, W_l) are read in place (channel a*K+k / a*4+j, retinanet.py:40-43) -- the permute + cat that
 * RetinaNet.inference does per level (retinanet.py:444-447) is never materialised.  Results are identical to
 * fsg_detect on the permuted copy: candidates are ranked by (score, index in the reference's (h, w, a, k) order).
 * anchors: flat (R,4) / (N,R,4) in the reference's order, R = sum_l H_l*W_l*A.  Workspace: fsg_detect_workspace_bytes. */
typedef struct fsg_detect_level {
  const float* logits;
  const float* deltas;
  int32_t H, W;
} fsg_detect_level;
FSG_API int fsg_detect_levels(const fsg_detect_level* h_levels, int num_levels, int A, int K, const float* anchors,
                      int64_t anchor_image_stride, int N, int64_t R, float score_threshold, int topk,
                      double nms_threshold, int max_det, const float* h_box_weights, float scale_clamp,
                      float* out_boxes, float* out_scores, int64_t* out_classes, int32_t* out_count,
                      float* cand_boxes, float* cand_scores, int64_t* cand_classes, int32_t* cand_count,
                      int64_t* keep_idx, const float* postprocess, void* workspace, size_t workspace_bytes,
                      fsg_stream_t stream);

/* detector_postprocess on an arbitrary box list (e.g. an Instances of another detector head):
 * out_boxes[i] = clip(scale(boxes[i])), keep[i] = 1 iff the result is non-empty (width > 0 and height > 0). */
FSG_API int fsg_postprocess_boxes(const float* boxes, int64_t n, float scale_x, float scale_y, float clip_w,
                          float clip_h, float* out_boxes, uint8_t* keep, fsg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Other callers of the same ops: RPN proposal selection
 * ---------------------------------------------------------------------------------------- */

/* find_top_rpn_proposals (proposal_generator/rpn_outputs.py:52-151) for the whole batch in two launches.
 * Level l: proposals (N, S_l, 4) decoded boxes and objectness logits (N, S_l), S_l = h_level_sizes[l]
 * (host arrays of device pointers).  image_sizes: device (N, 2) fp32 rows [height, width].
 * Per (image, level): the min(pre_nms_topk, S_l) best logits (ties: lower index first), Boxes.clip to the image,
 * boxes with a side <= min_box_side_len dropped; then per-level NMS (batched_nms with the level as class id,
 * un-offset per-class algorithm) and the post_nms_topk best survivors by logit.
 * Outputs: out_boxes (N, post_nms_topk, 4), out_logits (N, post_nms_topk), out_levels (N, post_nms_topk) int64
 * (may be NULL), out_count (N) int32; rows >= out_count[n] are zero-filled.
 * Up to 8192 candidates per NMS CTA (every FPN setting) the per-level NMS runs in shared memory: two launches for
 * the batch.  Beyond that -- pre_nms_topk up to 16384 per level, e.g. RPN.PRE_NMS_TOPK_TRAIN = 12000 of the C4
 * models (config/defaults.py:219) -- the same select kernel is followed by the general-n NMS per image (3 launches
 * each) and a gather.  Limits: pre_nms_topk <= 16384 per level, num_levels * pre_nms_topk <= 262144. */
FSG_API size_t fsg_rpn_proposals_workspace_bytes(int N, const int64_t* h_level_sizes, int num_levels,
                                         int pre_nms_topk, int post_nms_topk);
FSG_API int fsg_rpn_proposals(const float* const* h_level_proposals, const float* const* h_level_logits,
                      const int64_t* h_level_sizes, int num_levels, int N, const float* image_sizes,
                      int pre_nms_topk, int post_nms_topk, double nms_threshold, float min_box_side_len,
                      float* out_boxes, float* out_logits, int64_t* out_levels, int32_t* out_count,
                      void* workspace, size_t workspace_bytes, fsg_stream_t stream);

/* Candidate stage of fast_rcnn_inference_single_image (roi_heads/fast_rcnn.py:76-105): scores (R, K+1) with the
 * background column last, boxes (R, C*4) with C = num_bbox_reg_classes in {1, K}.  Boxes are clipped to the
 * image; every (r, k) with scores[r,k] > score_thresh is emitted in row-major order (torch.nonzero order):
 * out_boxes (cap,4), out_scores (cap), out_classes (cap) int64 = k, out_rows (cap) int64 = r, *out_count (device
 * int32); cap = R*K in the worst case.  Follow with fsg_nms(class_ids = out_classes) for the per-class NMS. */
FSG_API int fsg_score_filter(const float* boxes, int num_bbox_reg_classes, const float* scores, int64_t R, int K,
                     float image_height, float image_width, float score_thresh, float* out_boxes,
                     float* out_scores, int64_t* out_classes, int64_t* out_rows, int32_t* out_count,
                     fsg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FSG_DENSE_H_ */
