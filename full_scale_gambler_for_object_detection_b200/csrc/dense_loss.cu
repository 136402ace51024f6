// K2: fused Box2Box encode + gambler-weighted sigmoid-focal / smooth-L1 loss, forward + backward.
//
// Reference semantics (paths relative to the reference tree):
//   detectron2/modeling/meta_arch/retinanet.py:201-248      RetinaNet.losses
//   ImbalanceDetection/imbalancedetection/gambler_heads.py:104-128  calc_cls_loss
//                                                    :131-253  calc_gambler_loss (L_BAHW, L_BAHW_extendtobatch)
//                                                    :291-318  bet normalisation
//                                                    :502-602  LayeredUnetGambler.gambler_loss
//   ImbalanceDetection/train_net.py:1089-1098               loss combination
//
// One pass over the (N, R, K) logits: each anchor row is owned by a group of G lanes that read it
// with 16-byte streaming loads, evaluate focal loss and its derivative sharing one exp2 and one
// reciprocal per element (log1p by a degree-7 polynomial, so the per-element relative error stays
// ~3e-7 instead of MUFU.LG2's 1e-5 near 1), write the gradient once, and reduce the per-anchor loss
// with shuffles.  The gradient scale needs max(1, num_foreground) and the per-image bet normaliser
// S[n] up front (pre-pass over (N,R)-sized data, fused into K1's second pass), and d/d bets needs
// A[n] = sum_r w_hat*l afterwards (post pass over (N,R)-sized data).
#include "common.cuh"
#include "loss_math.cuh"
#include "step_internal.cuh"

namespace fsg {

#ifndef LOSS_GENERIC_MINB
#define LOSS_GENERIC_MINB 4
#endif
constexpr int kAnchorsPerGroup = 4;

enum LossVariant { kFastWrite = 0, kFastNoWrite = 1, kGeneric = 2 };

struct LossArgs {
  const float* logits;
  const float* pred_deltas;
  const float* gt_deltas;
  const float4* anchors;
  int64_t anchor_stride4;
  const float4* gt_boxes;
  const int32_t* gt_offsets;
  const int32_t* matched;
  const int64_t* gt_classes;
  const int64_t* mask;
  const float* bets;
  int N;
  int64_t R;
  int K;
  int nvec;       // K / V
  int G;          // lanes per anchor (power of two)
  int logG;
  int tiles_per_image;
  int anchors_per_tile;
  float a0, a1;   // alpha_t for t=0 / t=1 (1,1 when alpha < 0)
  float gamma, beta, T, ggamma;
  int gmode, nmode;
  float c_cls, c_reg, c_gam;
  float wx, wy, ww, wh;
  const double* stats;
  float* grad_logits;
  float* grad_deltas;
  float* ell;
  float* wout;
  float* partials;
  unsigned* counter;
  double* scalars;
  // sharded batch, fsg_dense_step: K1 only POSTED this rank's [num_foreground, S_batch] into the peers' mailboxes;
  // every CTA of this kernel polls its own rank's mailbox (local memory, L2 hits once the peers have arrived) and
  // sums the slots in rank order -- the NVLink latency hides behind the first logit loads already in flight
  PeerPoll peer;
};


// ------------------------------------------------------------------------------------------
// pre-pass (stand-alone form; K1's pass B carries the same reduction fused)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_prepass_kernel(const int64_t* __restrict__ gt_classes,
                                                           const int64_t* __restrict__ mask,
                                                           const float* __restrict__ bets, int N, int64_t R,
                                                           int num_classes, float temperature,
                                                           int* __restrict__ part_cnt, float* __restrict__ part_s,
                                                           unsigned* __restrict__ counter, double* __restrict__ stats) {
  __shared__ float s_red[8];
  __shared__ int s_redi[8];
  __shared__ double s_tc[8], s_ts[8];
  __shared__ bool s_last;
  const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t r = (int64_t)blockIdx.x * 256 + tid;
  int fg = 0;
  float w = 0.f;
  if (r < R) {
    const int64_t o = (int64_t)n * R + r;
    const int64_t c = gt_classes[o];
    fg = (c >= 0 && c != num_classes) ? 1 : 0;
    if (bets) w = __fadd_rn(__fmul_rn(bets[o], mask ? (float)mask[o] : 1.f), temperature);
  }
  int fgw = __reduce_add_sync(kFull, fg);
  float sw = warp_sum(w);
  if (lane == 0) { s_redi[wid] = fgw; s_red[wid] = sw; }
  __syncthreads();
  const int nb = gridDim.x;
  if (tid == 0) {
    int ci = 0; float cs = 0.f;
    for (int k = 0; k < 8; ++k) { ci += s_redi[k]; cs += s_red[k]; }
    part_cnt[n * nb + blockIdx.x] = ci;
    part_s[n * nb + blockIdx.x] = cs;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == (unsigned)(nb * N) - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tc = 0.0, ts = 0.0;
  for (int img = wid; img < N; img += 8) {
    double c = 0.0, s = 0.0;
    for (int b = lane; b < nb; b += 32) {
      c += (double)__ldcg(&part_cnt[img * nb + b]);
      s += (double)__ldcg(&part_s[img * nb + b]);
    }
    c = warp_sum_d(c); s = warp_sum_d(s);
    if (lane == 0) { stats[FSG_STATS_HEADER + img] = s; tc += c; ts += s; }
  }
  if (lane == 0) { s_tc[wid] = tc; s_ts[wid] = ts; }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, s = 0.0;
    for (int k = 0; k < 8; ++k) { c += s_tc[k]; s += s_ts[k]; }
    stats[0] = c; stats[1] = s;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// main pass
// ------------------------------------------------------------------------------------------
// PEER: the sharded form (fsg_dense_step with a peer context).  A separate instantiation, because even dead peer code
// in the prologue and in finish_tile cost the single-GPU kernel 1.6 us (133.6 -> 135.2 us) through register allocation.
template <int V, int BATCH, int VARIANT, int GT, bool PEER = false>
__global__ void __launch_bounds__(kLossBlock, (GT > 0 ? 4 : LOSS_GENERIC_MINB)) loss_main_kernel(const LossArgs A) {
  // GT > 0: "exact" instantiation, G == GT lanes per anchor and K == V*GT*BATCH (no predication in the
  // element loop); GT == 0: G and K are run-time values.
  constexpr bool kWrite = (VARIANT != kFastNoWrite);
  constexpr bool kFast = (VARIANT != kGeneric);
  grid_launch_dependents();   // the post pass may be scheduled as this grid's CTAs retire; it waits for the whole grid
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int G = GT > 0 ? GT : A.G;
  const int logG = GT > 0 ? (GT == 1 ? 0 : GT == 2 ? 1 : GT == 4 ? 2 : GT == 8 ? 3 : GT == 16 ? 4 : 5) : A.logG;
  const int nvec = GT > 0 ? GT * BATCH : A.nvec;
  const int gl = tid & (G - 1);           // lane within the anchor group
  const int grp = tid >> logG;            // group within the block
  const int ngrp = kLossBlock >> logG;
  const int64_t tile_base = (int64_t)blockIdx.x * A.anchors_per_tile;

  // Sharded run: this CTA is about to wait for the peers' sums to cross NVLink.  The logits are not produced by K1,
  // so it pulls its tile's rows into L2 first and the HBM latency of the first wave hides behind the exchange.
  // (Not done on one GPU: with nothing to hide, the extra L2 requests cost the main pass 2 %.)
  // Only the first wave of CTAs waits (the later ones find the sums published), so only it prefetches.
  if (PEER && A.peer.world > 1 && blockIdx.y * gridDim.x + blockIdx.x < 148 * 4) {
    const char* rows = reinterpret_cast<const char*>(A.logits + ((int64_t)n * A.R + tile_base) * A.K);
    int64_t bytes = (min((int64_t)A.anchors_per_tile, A.R - tile_base)) * A.K * (int64_t)sizeof(float);
    if (bytes > 96 * 1024) bytes = 96 * 1024;
    for (int64_t off = (int64_t)tid * 128; off < bytes; off += kLossBlock * 128) prefetch_l2(rows + off);
  }
  // Everything K1 produced is read from here on.
  grid_dependency_sync();
  // normalisers: every thread derives them itself from two broadcast loads (no block barrier behind one
  // thread's fp64 divide).  num_foreground is an integer < 2^24, so the fp32 reciprocal of max(1, nf) is the
  // reference's own fp32 division; S[n] is rounded to fp32 like the reference's fp32 sum.
  double nf_d = A.stats[0];
  float s_batch = (A.nmode == FSG_NORM_BATCH) ? (float)A.stats[1] : 1.f;
  if (PEER && A.peer.world > 1) {   // sharded batch: num_foreground of the whole batch arrives through the peer mailboxes
    const double nf_local = nf_d, sb_local = A.stats[1];
    peer_poll_nf(A.peer, nf_d);
    // (S_batch is not a normaliser of a sharded step -- FSG_NORM_BATCH is refused there; the ranks' complete records
    //  are for the step's statistics and are posted off the critical path, by one CTA, after its own wait)
    if (blockIdx.x == 0 && blockIdx.y == 0) peer_post_full(A.peer, nf_local, sb_local);
  }
  const float inv_nf = __frcp_rn(fmaxf((float)nf_d, 1.f));
  float inv_S = 1.f;
  if (A.nmode == FSG_NORM_IMAGE) inv_S = __frcp_rn((float)A.stats[FSG_STATS_HEADER + n]);
  else if (A.nmode == FSG_NORM_BATCH) inv_S = __frcp_rn(s_batch);
  const int m0 = A.gt_offsets ? A.gt_offsets[n] : 0;
  const bool write_grad = kWrite && (A.grad_logits != nullptr);

  float acc_cls = 0.f, acc_reg = 0.f, acc_wl = 0.f, acc_l = 0.f, max_l = 0.f;

#pragma unroll 1
  for (int it = 0; it < kAnchorsPerGroup; ++it) {
    const int64_t r = tile_base + (int64_t)it * ngrp + grp;
    const bool live = r < A.R;   // uniform within the group
    const int64_t o = (int64_t)n * A.R + (live ? r : 0);
    const float* xrow = A.logits + o * A.K + gl * V;
    float* grow = write_grad ? A.grad_logits + o * A.K + gl * V : nullptr;

    // ---- issue the row loads first, then the per-anchor metadata (same address for the G lanes of a
    //      group: one broadcast request); the arithmetic below hides under both
    Vec<V> x[BATCH];
    if (GT > 0) {
      if (live) {
#pragma unroll
        for (int b = 0; b < BATCH; ++b) x[b].load(xrow + b * (GT * V));
      }
    }
    int cls = -1;
    float w_hat = 0.f;
    if (live) {
      cls = (int)A.gt_classes[o];
      if (A.bets) {
        const float m = A.mask ? (float)A.mask[o] : 1.f;
        const float w = __fadd_rn(__fmul_rn(A.bets[o], m), A.T);   // gambler_heads.py:569,304
        w_hat = w * inv_S;                                        // :308-311
      }
    }
    const bool valid = cls >= 0;
    const bool fgc = valid && cls != A.K;
    const float wg = (A.ggamma == 1.f) ? w_hat : powf(w_hat, A.ggamma);
    // d total / d x = f'(x) * coef_f + bce'(x) * coef_b
    float coef_f = 0.f, coef_b = 0.f;
    if (valid) {
      coef_f = A.c_cls * inv_nf;
      if (A.gmode == FSG_CLS_FOCAL) coef_f = fmaf(-A.c_gam, wg, coef_f);
      else coef_b = -A.c_gam * wg;
    }
    const int jpos = fgc ? cls / V : -1;   // vector that holds the positive class
    const int kpos = fgc ? cls - jpos * V : 0;

    float sum_f = 0.f, sum_b = 0.f;       // per-lane partial of sum_k focal / sum_k bce (unscaled for t=0)
    float pos_fix = 0.f;                  // (alpha_1 f_1 - alpha_0 f_0) of the positive element

    if (live && !valid) {
      // ignored anchor: zero loss, zero gradient (gambler_heads.py:554-555, retinanet.py:233)
      if (write_grad) {
        Vec<V> z;
#pragma unroll
        for (int k = 0; k < V; ++k) z.v[k] = 0.f;
        for (int j = gl; j < nvec; j += G) z.store(grow + (int64_t)(j - gl) * V);
      }
    } else if (live) {
      if (GT > 0 && kFast) {
        // exact fast path: every element treated as a negative, the one positive is patched afterwards
        const float cf = coef_f * A.a0;
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          Vec<V> g;
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float l, d;
            focal_neg_g2(x[b].v[k], l, d);
            sum_f += l;
            g.v[k] = d * cf;
          }
          if (write_grad) g.store(grow + b * (GT * V));
        }
        if (fgc && (jpos & (GT - 1)) == gl) {   // rare: this lane owns the positive class
          const float xv = A.logits[o * A.K + cls];
          float l0, d0, f1, fg1, b1, bg1;
          focal_neg_g2(xv, l0, d0);
          cls_elem_general(xv, true, 2.f, f1, fg1, b1, bg1);
          pos_fix = f1 * A.a1 - l0 * A.a0;
          if (write_grad) A.grad_logits[o * A.K + cls] = fg1 * (coef_f * A.a1);
        }
      } else if (kFast) {
        // any K, fast variants: whole batches without predication, one predicated tail batch, the positive class
        // patched afterwards (no per-vector test, no register double buffer: 4 CTAs/SM hide the load latency)
        const float cf = coef_f * A.a0;
        const int span = G * BATCH;
        // y holds the batch being evaluated, z the next one (in flight meanwhile); a batch is "full" when every
        // lane of the group has all BATCH vectors in range, only the last one is predicated
        Vec<V> y[BATCH], z[BATCH];
        int j0 = gl;
        auto load_batch = [&](Vec<V>* dst, int jb) {
          if (jb + span - gl <= nvec) {
#pragma unroll
            for (int b = 0; b < BATCH; ++b) dst[b].load(xrow + (int64_t)(jb - gl + b * G) * V);
          } else {
#pragma unroll
            for (int b = 0; b < BATCH; ++b)
              if (jb + b * G < nvec) dst[b].load(xrow + (int64_t)(jb - gl + b * G) * V);
          }
        };
        load_batch(y, j0);
#pragma unroll 1
        for (; j0 - gl < nvec; j0 += span) {
          if (j0 - gl + span < nvec) load_batch(z, j0 + span);
          const bool full = (j0 + span - gl <= nvec);
#pragma unroll
          for (int b = 0; b < BATCH; ++b) {
            if (full || j0 + b * G < nvec) {
              Vec<V> g;
#pragma unroll
              for (int k = 0; k < V; ++k) {
                float l, d;
                focal_neg_g2(y[b].v[k], l, d);
                sum_f += l;
                g.v[k] = d * cf;
              }
              if (write_grad) g.store(grow + (int64_t)(j0 - gl + b * G) * V);
            }
          }
#pragma unroll
          for (int b = 0; b < BATCH; ++b) y[b] = z[b];
        }
        if (fgc && (jpos & (G - 1)) == gl) {   // rare: this lane owns the positive class
          const float xv = A.logits[o * A.K + cls];
          float l0, d0, f1, fg1, b1, bg1;
          focal_neg_g2(xv, l0, d0);
          cls_elem_general(xv, true, 2.f, f1, fg1, b1, bg1);
          pos_fix = f1 * A.a1 - l0 * A.a0;
          if (write_grad) A.grad_logits[o * A.K + cls] = fg1 * (coef_f * A.a1);
        }
      } else {
        // double-buffered: the next batch of row pieces is in flight while this one is evaluated
        Vec<V> y[BATCH], z[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          const int j = gl + b * G;
          if (j < nvec) y[b].load(xrow + (int64_t)(j - gl) * V);
        }
        for (int j0 = gl; j0 < nvec; j0 += G * BATCH) {
#pragma unroll
          for (int b = 0; b < BATCH; ++b) {
            const int j = j0 + G * BATCH + b * G;
            if (j < nvec) z[b].load(xrow + (int64_t)(j - gl) * V);
          }
#pragma unroll
          for (int b = 0; b < BATCH; ++b) {
            const int j = j0 + b * G;
            if (j < nvec) {
              Vec<V> g;
              if (kFast) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                  float l, d;
                  focal_neg_g2(y[b].v[k], l, d);
                  sum_f += l;
                  g.v[k] = d * (coef_f * A.a0);
                }
                if (j == jpos) {   // rare: patch the single positive element of this anchor
#pragma unroll
                  for (int k = 0; k < V; ++k) {
                    if (k == kpos) {
                      float l0, d0, f1, fg1, b1, bg1;
                      focal_neg_g2(y[b].v[k], l0, d0);
                      cls_elem_general(y[b].v[k], true, 2.f, f1, fg1, b1, bg1);
                      pos_fix = f1 * A.a1 - l0 * A.a0;
                      g.v[k] = fg1 * (coef_f * A.a1);
                    }
                  }
                }
              } else {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                  const bool t = (j == jpos) && (k == kpos);
                  float f, fgd, bc, bgd;
                  cls_elem_general(y[b].v[k], t, A.gamma, f, fgd, bc, bgd);
                  const float at = t ? A.a1 : A.a0;
                  sum_f += f * at;
                  sum_b += bc;
                  g.v[k] = fmaf(fgd * at, coef_f, bgd * coef_b);
                }
              }
              if (write_grad) g.store(grow + (int64_t)(j - gl) * V);
            }
          }
#pragma unroll
          for (int b = 0; b < BATCH; ++b) y[b] = z[b];
        }
      }
    }
    if (kFast) sum_f = fmaf(sum_f, A.a0, pos_fix);
    // ---- reduce over the group's lanes
    for (int s = G >> 1; s > 0; s >>= 1) {
      sum_f += __shfl_xor_sync(kFull, sum_f, s);
      if (!kFast) sum_b += __shfl_xor_sync(kFull, sum_b, s);
    }
    if (live && gl == 0) {
      const float lf = valid ? sum_f : 0.f;                                   // gambler_heads.py:554-555
      const float lg = (A.gmode == FSG_CLS_FOCAL) ? lf : (valid ? sum_b : 0.f);
      if (A.ell) A.ell[o] = lg;
      if (A.wout) A.wout[o] = w_hat;
      acc_cls += lf;
      acc_wl = fmaf(wg, lg, acc_wl);
      acc_l += lg;
      max_l = fmaxf(max_l, lg);
    }
    // ---- regression: one lane per anchor (the group's last lane, to spread the work)
    if (live && gl == G - 1 && A.pred_deltas) {
      float4 gout = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fgc) {
        const float4 pd = reinterpret_cast<const float4*>(A.pred_deltas)[o];
        float4 gd;
        if (A.gt_deltas) gd = reinterpret_cast<const float4*>(A.gt_deltas)[o];
        else gd = encode_deltas_loss(A.anchors[(int64_t)n * A.anchor_stride4 + r],
                                     A.gt_boxes[m0 + A.matched[o]], A.wx, A.wy, A.ww, A.wh);
        const float sc = A.c_reg * inv_nf;
        float l, g;
        smooth_l1_elem(pd.x, gd.x, A.beta, l, g); acc_reg += l; gout.x = g * sc;
        smooth_l1_elem(pd.y, gd.y, A.beta, l, g); acc_reg += l; gout.y = g * sc;
        smooth_l1_elem(pd.z, gd.z, A.beta, l, g); acc_reg += l; gout.z = g * sc;
        smooth_l1_elem(pd.w, gd.w, A.beta, l, g); acc_reg += l; gout.w = g * sc;
      }
      if (A.grad_deltas) reinterpret_cast<float4*>(A.grad_deltas)[o] = gout;
    }
  }

  // ---- tile partials -> scalars (deterministic: one slot per tile, fixed reduction order)
  acc_cls = warp_sum(acc_cls); acc_reg = warp_sum(acc_reg); acc_wl = warp_sum(acc_wl);
  acc_l = warp_sum(acc_l); max_l = warp_max(max_l);
  finish_tile<kLossBlock, PEER>(acc_cls, acc_reg, acc_wl, acc_l, max_l, n, blockIdx.x, A.tiles_per_image, A.N, A.partials, A.counter,
              A.scalars, nf_d, A.c_cls, A.c_reg, A.c_gam, A.peer);
}

// ------------------------------------------------------------------------------------------
// post pass: d(c_gam * G)/d bets
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float post_one(float b, float m, float l, float T, float ggamma, int nmode,
                                          float inv_S, float Asum) {
  const float w = __fadd_rn(__fmul_rn(b, m), T);
  if (nmode == FSG_NORM_NONE) {
    const float pw = (ggamma == 1.f) ? 1.f : powf(w, ggamma - 1.f);
    return -m * ggamma * pw * l;
  }
  const float w_hat = w * inv_S;
  const float pw = (ggamma == 1.f) ? 1.f : powf(w_hat, ggamma - 1.f);
  return -(m * inv_S) * ggamma * (pw * l - Asum);
}

__global__ void __launch_bounds__(256) loss_post_kernel(const float* __restrict__ bets,
                                                        const int64_t* __restrict__ mask,
                                                        const float* __restrict__ ell, int N, int64_t R, float T,
                                                        float ggamma, int nmode, float c_gam,
                                                        const double* __restrict__ stats,
                                                        const double* __restrict__ scalars,
                                                        float* __restrict__ grad_bets) {
  const int n = blockIdx.y;
  grid_dependency_sync();   // (programmatic dependent launch behind the main pass)
  // every thread derives the two per-image constants itself (broadcast loads, one fp32 reciprocal): a block
  // barrier behind one thread's fp64 divide cost more than the whole rest of this kernel
  float inv_S = 1.f, Asum = 0.f;
  if (nmode != FSG_NORM_NONE) {
    const double S = (nmode == FSG_NORM_IMAGE) ? stats[FSG_STATS_HEADER + n] : stats[1];
    const double A = (nmode == FSG_NORM_IMAGE) ? scalars[FSG_SCALARS_HEADER + n] : scalars[2];
    inv_S = __frcp_rn((float)S);
    Asum = (float)A;
  }
  const int64_t r0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (r0 >= R) return;
  const int64_t o = (int64_t)n * R + r0;
  if (r0 + 4 <= R && (o & 3) == 0) {
    const float4 b = *reinterpret_cast<const float4*>(bets + o);
    const float4 l = *reinterpret_cast<const float4*>(ell + o);
    float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
    if (mask) {
      const longlong2 ma = *reinterpret_cast<const longlong2*>(mask + o);
      const longlong2 mb = *reinterpret_cast<const longlong2*>(mask + o + 2);
      m0 = (float)ma.x; m1 = (float)ma.y; m2 = (float)mb.x; m3 = (float)mb.y;
    }
    float4 g;
    g.x = c_gam * post_one(b.x, m0, l.x, T, ggamma, nmode, inv_S, Asum);
    g.y = c_gam * post_one(b.y, m1, l.y, T, ggamma, nmode, inv_S, Asum);
    g.z = c_gam * post_one(b.z, m2, l.z, T, ggamma, nmode, inv_S, Asum);
    g.w = c_gam * post_one(b.w, m3, l.w, T, ggamma, nmode, inv_S, Asum);
    *reinterpret_cast<float4*>(grad_bets + o) = g;
  } else {
    for (int k = 0; k < 4 && r0 + k < R; ++k) {
      const float m = mask ? (float)mask[o + k] : 1.f;
      grad_bets[o + k] = c_gam * post_one(bets[o + k], m, ell[o + k], T, ggamma, nmode, inv_S, Asum);
    }
  }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, int64_t n4, int64_t n,
                                                    const float* __restrict__ sdev, float shost) {
  const float s = sdev ? *sdev * shost : shost;
  if (s == 1.f) return;  // unit upstream gradient: nothing to do (checked on the device, no host sync)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) x[i] *= s;
  }
}

struct LossPlan {
  int V, G, logG, batch, nvec, anchors_per_tile, tiles_per_image;
};

static LossPlan plan_loss(int64_t R, int K) {
  LossPlan p;
  p.V = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);
  p.nvec = K / p.V;
  // pick lanes-per-anchor G (power of two) and the unroll batch to minimise predicated-off work;
  // prefer G >= 4 so a group's request covers at least two full 32-byte sectors
  double best = 1e30;
  p.G = 32; p.batch = 4;
  for (int g = 1; g <= 32; g <<= 1) {
    if (g < 4 && p.nvec >= 4) continue;
    for (int b = 4; b <= 5; ++b) {
      const int64_t span = (int64_t)g * b;
      const double waste = (double)(ceil_div(p.nvec, span) * span) / p.nvec;
      const double cost = waste + 0.002 * g + (ceil_div(p.nvec, span) > 8 ? 0.5 : 0.0);
      if (cost < best - 1e-9) { best = cost; p.G = g; p.batch = b; }
    }
  }
  p.logG = 0;
  while ((1 << p.logG) < p.G) ++p.logG;
  p.anchors_per_tile = (kLossBlock / p.G) * kAnchorsPerGroup;
  p.tiles_per_image = (int)ceil_div(R > 0 ? R : 1, p.anchors_per_tile);
  return p;
}

struct LossWs {
  size_t off_counter, off_partials, total;
};
static LossWs loss_ws_layout(int N, const LossPlan& p) {
  LossWs w;
  size_t o = 0;
  w.off_counter = o; o += 16;
  w.off_partials = o; o += align_up(sizeof(float) * kPartialStride * (size_t)N * p.tiles_per_image, 16);
  w.total = o;
  return w;
}

template <int V, int BATCH, int GT>
static void launch_main(int variant, dim3 grid, cudaStream_t s, const LossArgs& a, bool pdl) {
  if (a.peer.world > 1) {
    if (variant == kFastWrite) launch_pdl(loss_main_kernel<V, BATCH, kFastWrite, GT, true>, grid, dim3(kLossBlock), 0, s, pdl, a);
    else if (variant == kFastNoWrite) launch_pdl(loss_main_kernel<V, BATCH, kFastNoWrite, GT, true>, grid, dim3(kLossBlock), 0, s, pdl, a);
    else launch_pdl(loss_main_kernel<V, BATCH, kGeneric, GT, true>, grid, dim3(kLossBlock), 0, s, pdl, a);
    return;
  }
  if (variant == kFastWrite) launch_pdl(loss_main_kernel<V, BATCH, kFastWrite, GT>, grid, dim3(kLossBlock), 0, s, pdl, a);
  else if (variant == kFastNoWrite) launch_pdl(loss_main_kernel<V, BATCH, kFastNoWrite, GT>, grid, dim3(kLossBlock), 0, s, pdl, a);
  else launch_pdl(loss_main_kernel<V, BATCH, kGeneric, GT>, grid, dim3(kLossBlock), 0, s, pdl, a);
}

}  // namespace fsg

using namespace fsg;

extern "C" size_t fsg_loss_prepass_workspace_bytes(int N, int64_t R) {
  if (N <= 0 || R < 0) return 0;
  const size_t nb = (size_t)ceil_div(R > 0 ? R : 1, 256);
  return 16 + align_up(sizeof(int) * N * nb, 16) + align_up(sizeof(float) * N * nb, 16);
}

extern "C" int fsg_loss_prepass(const int64_t* gt_classes, const int64_t* mask, const float* bets, int N,
                                int64_t R, int num_classes, float temperature, double* stats, void* workspace,
                                size_t workspace_bytes, fsg_stream_t stream) {
  if (N <= 0 || R <= 0 || !gt_classes || !stats) return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < fsg_loss_prepass_workspace_bytes(N, R) || ((uintptr_t)workspace & 15))
    return FSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nb = (size_t)ceil_div(R, 256);
  char* ws = (char*)workspace;
  unsigned* counter = (unsigned*)ws;
  int* pc = (int*)(ws + 16);
  float* ps = (float*)(ws + 16 + align_up(sizeof(int) * N * nb, 16));
  FSG_CUDA_TRY(cudaMemsetAsync(counter, 0, 16, s));
  dim3 grid((unsigned)nb, (unsigned)N);
  loss_prepass_kernel<<<grid, 256, 0, s>>>(gt_classes, mask, bets, N, R, num_classes, temperature, pc, ps,
                                           counter, stats);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" size_t fsg_loss_main_workspace_bytes(int N, int64_t R, int K) {
  if (N <= 0 || R < 0 || K <= 0) return 0;
  return loss_ws_layout(N, plan_loss(R, K)).total;
}

namespace fsg {
int loss_main_enqueue(const float* logits, const float* pred_deltas, const float* gt_deltas,
                      const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                      const int32_t* gt_offsets, const int32_t* matched_idx32, const int64_t* gt_classes,
                      const int64_t* mask, const float* bets, int N, int64_t R,
                      const fsg_loss_params* hp, double* stats, float* grad_logits,
                      float* grad_deltas, float* per_anchor_loss, float* weights_out, double* scalars,
                      void* workspace, size_t workspace_bytes, const fsg_peer_ctx* h_peer, int flags,
                      fsg_stream_t stream) {
  if (!hp || N <= 0 || R <= 0 || hp->num_classes <= 0) return FSG_ERR_INVALID_ARG;
  if (!logits || !gt_classes || !stats || !scalars) return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  if (anchor_image_stride % 4 != 0) return FSG_ERR_INVALID_ARG;
  if (pred_deltas && !gt_deltas && (!anchors || !gt_boxes || !gt_offsets || !matched_idx32))
    return FSG_ERR_INVALID_ARG;
  if (grad_deltas && !pred_deltas) return FSG_ERR_INVALID_ARG;
  if (!bets && (hp->c_gam != 0.f || weights_out)) return FSG_ERR_INVALID_ARG;
  if (hp->gambler_mode != FSG_CLS_FOCAL && hp->gambler_mode != FSG_CLS_SIGMOID) return FSG_ERR_INVALID_ARG;
  if (hp->norm_mode < FSG_NORM_NONE || hp->norm_mode > FSG_NORM_BATCH) return FSG_ERR_INVALID_ARG;
  const int K = hp->num_classes;
  const LossPlan p = plan_loss(R, K);
  const LossWs w = loss_ws_layout(N, p);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  if (p.V == 4 && (((uintptr_t)logits | (uintptr_t)grad_logits) & 15)) return FSG_ERR_INVALID_ARG;
  if (p.V == 2 && (((uintptr_t)logits | (uintptr_t)grad_logits) & 7)) return FSG_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;

  LossArgs a;
  a.logits = logits; a.pred_deltas = pred_deltas; a.gt_deltas = gt_deltas;
  a.anchors = (const float4*)anchors; a.anchor_stride4 = anchor_image_stride / 4;
  a.gt_boxes = (const float4*)gt_boxes; a.gt_offsets = gt_offsets; a.matched = matched_idx32;
  a.gt_classes = gt_classes; a.mask = mask; a.bets = bets;
  a.N = N; a.R = R; a.K = K; a.nvec = p.nvec; a.G = p.G; a.logG = p.logG;
  a.tiles_per_image = p.tiles_per_image; a.anchors_per_tile = p.anchors_per_tile;
  a.a0 = hp->focal_alpha >= 0.f ? 1.f - hp->focal_alpha : 1.f;
  a.a1 = hp->focal_alpha >= 0.f ? hp->focal_alpha : 1.f;
  a.gamma = hp->focal_gamma; a.beta = hp->smooth_l1_beta; a.T = hp->temperature; a.ggamma = hp->gambler_gamma;
  a.gmode = hp->gambler_mode; a.nmode = bets ? hp->norm_mode : FSG_NORM_NONE;
  a.c_cls = hp->c_cls; a.c_reg = hp->c_reg; a.c_gam = hp->c_gam;
  a.wx = hp->box_weights[0]; a.wy = hp->box_weights[1]; a.ww = hp->box_weights[2]; a.wh = hp->box_weights[3];
  a.stats = stats; a.grad_logits = grad_logits; a.grad_deltas = grad_deltas; a.ell = per_anchor_loss;
  a.wout = weights_out; a.partials = (float*)(ws + w.off_partials); a.counter = (unsigned*)(ws + w.off_counter);
  a.scalars = scalars;
  a.peer = make_peer_poll(h_peer, stats);

  const bool pdl = (flags & kLossPdl) != 0;
  if (!(flags & kLossCounterZeroed)) FSG_CUDA_TRY(cudaMemsetAsync(a.counter, 0, 16, s));
  const bool fast = (hp->focal_gamma == 2.f) && (hp->gambler_mode == FSG_CLS_FOCAL);
  const int variant = fast ? (grad_logits ? kFastWrite : kFastNoWrite) : kGeneric;
  dim3 grid((unsigned)p.tiles_per_image, (unsigned)N);
  const bool exact = (p.nvec == p.G * p.batch);
  if (p.V == 4 && p.G == 4 && p.batch == 5 && exact && variant != kGeneric) launch_main<4, 5, 4>(variant, grid, s, a, pdl);  // K = 80
  else if (p.V == 4) { if (p.batch == 5) launch_main<4, 5, 0>(variant, grid, s, a, pdl); else launch_main<4, 4, 0>(variant, grid, s, a, pdl); }
  else if (p.V == 2) { if (p.batch == 5) launch_main<2, 5, 0>(variant, grid, s, a, pdl); else launch_main<2, 4, 0>(variant, grid, s, a, pdl); }
  else { if (p.batch == 5) launch_main<1, 5, 0>(variant, grid, s, a, pdl); else launch_main<1, 4, 0>(variant, grid, s, a, pdl); }
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

int loss_post_enqueue(const float* bets, const int64_t* mask, const float* per_anchor_loss, int N,
                      int64_t R, const fsg_loss_params* hp, const double* stats, const double* scalars,
                      float* grad_bets, int flags, fsg_stream_t stream) {
  if (!hp || N <= 0 || R <= 0 || !bets || !per_anchor_loss || !stats || !scalars || !grad_bets)
    return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(R, 1024), (unsigned)N);
  launch_pdl(loss_post_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (flags & kLossPdl) != 0, bets, mask,
             per_anchor_loss, N, R, hp->temperature, hp->gambler_gamma, (int)hp->norm_mode, hp->c_gam, stats, scalars,
             grad_bets);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

size_t loss_main_ws_bytes(int N, int64_t R, int K) { return loss_ws_layout(N, plan_loss(R, K)).total; }
}  // namespace fsg

extern "C" int fsg_loss_main(const float* logits, const float* pred_deltas, const float* gt_deltas,
                             const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                             const int32_t* gt_offsets, const int32_t* matched_idx32, const int64_t* gt_classes,
                             const int64_t* mask, const float* bets, int N, int64_t R,
                             const fsg_loss_params* hp, const double* stats, float* grad_logits,
                             float* grad_deltas, float* per_anchor_loss, float* weights_out, double* scalars,
                             void* workspace, size_t workspace_bytes, fsg_stream_t stream) {
  return loss_main_enqueue(logits, pred_deltas, gt_deltas, anchors, anchor_image_stride, gt_boxes, gt_offsets,
                           matched_idx32, gt_classes, mask, bets, N, R, hp, const_cast<double*>(stats), grad_logits,
                           grad_deltas, per_anchor_loss, weights_out, scalars, workspace, workspace_bytes, nullptr, 0,
                           stream);
}

extern "C" int fsg_loss_post(const float* bets, const int64_t* mask, const float* per_anchor_loss, int N,
                             int64_t R, const fsg_loss_params* hp, const double* stats, const double* scalars,
                             float* grad_bets, fsg_stream_t stream) {
  return loss_post_enqueue(bets, mask, per_anchor_loss, N, R, hp, stats, scalars, grad_bets, 0, stream);
}

extern "C" int fsg_scale_inplace(float* x, int64_t n, const float* scale_dev, float scale_host,
                                 fsg_stream_t stream) {
  if (n < 0) return FSG_ERR_INVALID_ARG;
  if (n == 0) return FSG_OK;
  if (!x) return FSG_ERR_INVALID_ARG;
  const bool aligned = (((uintptr_t)x) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  int64_t blocks = ceil_div(n4 > 0 ? n4 : 1, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n4, n, scale_dev, scale_host);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
