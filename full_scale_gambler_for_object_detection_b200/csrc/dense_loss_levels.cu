// K2 on the head's native layout: the fused loss main pass reading the per-level conv outputs
// (N, A*K, H, W) / (N, A*4, H, W) directly and writing the gradients in the same layout, so the
// permute_to_N_HWA_K + cat copies (detectron2/modeling/meta_arch/retinanet.py:24-54,217-219 and again in
// ImbalanceDetection/imbalancedetection/gambler_heads.py:34-101,538-540) and their inverses in the backward
// disappear.  Same arithmetic, same scalars/partials as dense_loss.cu.
//
// Mapping: one thread per anchor (n, level, a, hw); consecutive lanes hold consecutive hw, so for every class
// channel k a warp reads / writes one contiguous 128-byte piece of the (a*K + k) plane -- fully coalesced
// without any shared-memory transpose, and the K-reduction of the per-anchor loss is a private register sum.
// The (N,R)-sized side arrays (gt_classes, mask, bets, l, w_hat) keep the flattened anchor order
// r = level_offset + hw*A + a of the rest of the library.
#include "common.cuh"
#include "loss_math.cuh"
#include "step_internal.cuh"

namespace fsg {

constexpr int kLvMax = FSG_MAX_LEVELS;

#ifndef LV_MINB
#define LV_MINB 4
#endif
#ifndef LV_VEC4
#define LV_VEC4 0
#endif
#ifndef LV_BLOCK
#define LV_BLOCK 256
#endif
constexpr int kLvBlock = LV_BLOCK;   // threads per CTA; a CTA covers kLvBlock * vec consecutive hw of one anchor slot

struct LevelTable {
  const float* logits[kLvMax];
  float* grad_logits[kLvMax];
  const float* pred_deltas[kLvMax];
  float* grad_deltas[kLvMax];
  const float* bets[kLvMax];  // per-level (N, A, H, W) betting maps / NAKHW_loss, or NULL: the flat (N, R) arrays
  float* ell[kLvMax];
  int64_t off[kLvMax + 1];   // first anchor index of the level
  int HW[kLvMax];
  int vec[kLvMax];           // hw positions per thread (2 when every plane of the level is 8-byte aligned)
  int chunks[kLvMax];        // ceil(HW / (kLvBlock * vec))
  int tile_base[kLvMax + 1]; // tiles of a level: A * chunks
  int num_levels, A;
};

struct LevelLossArgs {
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
  int tiles_per_image;
  float a0, a1, gamma, beta, T, ggamma;
  int gmode, nmode;
  float c_cls, c_reg, c_gam;
  float wx, wy, ww, wh;
  const double* stats;
  float* ell;
  float* wout;
  float* partials;
  unsigned* counter;
  double* scalars;
  PeerPoll peer;   // sharded batch, fsg_dense_step_levels: the exchange is polled here (see dense_loss.cu)
};

struct TileSums {
  float cls, reg, wl, l, mx;
};

// One thread: VEC consecutive hw positions of anchor slot `a` (VEC anchors), all K class planes.
// BATCH planes are in flight while the previous BATCH are evaluated; EXACT: K % BATCH == 0 (no predication in
// the element loop); FAST: focal gamma == 2 in focal gambler mode; WRITE: gradient planes are written.
template <int VEC, int BATCH, bool EXACT, bool FAST, bool WRITE>
__device__ __forceinline__ void level_body(const LevelLossArgs& A, const LevelTable& LT, int l, int a, int hw0, int n,
                                           float inv_nf, float inv_S, TileSums& S) {
  const int HW = LT.HW[l];
  const int K = A.K;
  const int64_t r0 = LT.off[l] + (int64_t)hw0 * LT.A + a;
  const int64_t o0 = (int64_t)n * A.R + r0;
  const int64_t plane0 = ((int64_t)n * LT.A + a) * K;
  const float* xcol = LT.logits[l] + plane0 * HW + hw0;
  float* gcol = WRITE ? LT.grad_logits[l] + plane0 * HW + hw0 : nullptr;

  Vec<VEC> x[BATCH];
#pragma unroll
  for (int b = 0; b < BATCH; ++b)
    if (EXACT || b < K) x[b].load(xcol + (int64_t)b * HW);

  // Per-anchor state that must live through the element loop is kept to the minimum (class id, the two gradient
  // coefficients, the running sums); the normalised weight is cheap to recompute afterwards from the same two
  // L1-resident loads, and not holding it across the loop is what keeps this kernel at 64 registers without spills.
  auto weight_of = [&](int v, int64_t o) -> float {
    if (!(A.bets || LT.bets[l])) return 0.f;
    const float m = A.mask ? (float)A.mask[o] : 1.f;
    const float b = LT.bets[l] ? LT.bets[l][((int64_t)n * LT.A + a) * HW + hw0 + v] : A.bets[o];
    return __fadd_rn(__fmul_rn(b, m), A.T) * inv_S;              // gambler_heads.py:569,304 and :308-311
  };
  int cls[VEC];
  float cf[VEC], cb[VEC], sum_f[VEC], sum_b[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const bool in = (VEC == 1) || (hw0 + v < HW);   // HW % VEC == 0 whenever VEC > 1, kept for clarity
    const int64_t o = o0 + (int64_t)v * LT.A;
    cls[v] = in ? (int)A.gt_classes[o] : -1;
    const float w_hat = in ? weight_of(v, o) : 0.f;
    const float wg = (A.ggamma == 1.f) ? w_hat : powf(w_hat, A.ggamma);
    float coef_f = 0.f, coef_b = 0.f;
    if (cls[v] >= 0) {
      coef_f = A.c_cls * inv_nf;
      if (A.gmode == FSG_CLS_FOCAL) coef_f = fmaf(-A.c_gam, wg, coef_f);
      else coef_b = -A.c_gam * wg;
    }
    cf[v] = FAST ? coef_f * A.a0 : coef_f;   // ignored anchors: zero coefficient -> zero gradient
    cb[v] = coef_b;
    sum_f[v] = 0.f;
    sum_b[v] = 0.f;
  }

  // Rotating pipeline: slot b is re-loaded with plane k + BATCH as soon as plane k has been taken out of it, so
  // BATCH planes stay in flight with ONE set of buffer registers (a second set for the "next batch" cost 20
  // registers and pushed this kernel into spills at 64 registers / 4 CTAs per SM).
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += BATCH) {
    const bool more = k0 + BATCH < K;
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int k = k0 + b;
      if (EXACT || k < K) {
        const Vec<VEC> cur = x[b];
        if (more && (EXACT || k + BATCH < K)) x[b].load(xcol + (int64_t)(k + BATCH) * HW);
        Vec<VEC> g;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (FAST) {
            float lo, d;
            focal_neg_g2(cur.v[v], lo, d);
            sum_f[v] += lo;
            g.v[v] = d * cf[v];
          } else {
            const bool t = (k == cls[v]);   // cls == K (background) or -1 (ignored) never matches
            float f, fgd, bc, bgd;
            cls_elem_general(cur.v[v], t, A.gamma, f, fgd, bc, bgd);
            const float at = t ? A.a1 : A.a0;
            sum_f[v] += f * at;
            sum_b[v] += bc;
            g.v[v] = fmaf(fgd * at, cf[v], bgd * cb[v]);
          }
        }
        if (WRITE) g.store(gcol + (int64_t)k * HW);
      }
    }
  }

#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int64_t o = o0 + (int64_t)v * LT.A;
    const bool valid = cls[v] >= 0;
    const bool fgc = valid && cls[v] != K;
    const float w_hat = ((VEC == 1) || (hw0 + v < HW)) ? weight_of(v, o) : 0.f;
    const float wg = (A.ggamma == 1.f) ? w_hat : powf(w_hat, A.ggamma);
    if (FAST) {
      sum_f[v] *= A.a0;
      if (fgc) {   // patch the single positive class of a foreground anchor
        const float xv = xcol[(int64_t)cls[v] * HW + v];
        float l0, d0, f1, fg1, b1, bg1;
        focal_neg_g2(xv, l0, d0);
        cls_elem_general(xv, true, 2.f, f1, fg1, b1, bg1);
        sum_f[v] += f1 * A.a1 - l0 * A.a0;
        const float coef_f = fmaf(-A.c_gam, wg, A.c_cls * inv_nf);   // (FAST implies the focal gambler mode)
        if (WRITE) gcol[(int64_t)cls[v] * HW + v] = fg1 * (coef_f * A.a1);
      }
    }
    const float lf = valid ? sum_f[v] : 0.f;                                   // gambler_heads.py:554-555
    const float lg = (A.gmode == FSG_CLS_FOCAL) ? lf : (valid ? sum_b[v] : 0.f);
    if (LT.ell[l]) LT.ell[l][((int64_t)n * LT.A + a) * HW + hw0 + v] = lg;
    else if (A.ell) A.ell[o] = lg;
    if (A.wout) A.wout[o] = w_hat;
    S.cls += lf;
    S.wl = fmaf(wg, lg, S.wl);
    S.l += lg;
    S.mx = fmaxf(S.mx, lg);
  }

  // ---- regression planes (a*4 + j)
  if (LT.pred_deltas[l]) {
    const int64_t dplane = ((int64_t)n * LT.A + a) * 4;
    const float* pcol = LT.pred_deltas[l] + dplane * HW + hw0;
    Vec<VEC> g4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int v = 0; v < VEC; ++v) g4[j].v[v] = 0.f;
    const float sc = A.c_reg * inv_nf;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (cls[v] >= 0 && cls[v] != K) {
        const int64_t o = o0 + (int64_t)v * LT.A;
        float4 gd;
        if (A.gt_deltas) gd = reinterpret_cast<const float4*>(A.gt_deltas)[o];
        else gd = encode_deltas_loss(A.anchors[(int64_t)n * A.anchor_stride4 + r0 + (int64_t)v * LT.A],
                                     A.gt_boxes[A.gt_offsets[n] + A.matched[o]], A.wx, A.wy, A.ww, A.wh);
        const float gdv[4] = {gd.x, gd.y, gd.z, gd.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, g;
          smooth_l1_elem(pcol[(int64_t)j * HW + v], gdv[j], A.beta, lo, g);
          S.reg += lo;
          g4[j].v[v] = g * sc;
        }
      }
    }
    if (LT.grad_deltas[l]) {
      float* dcol = LT.grad_deltas[l] + dplane * HW + hw0;
#pragma unroll
      for (int j = 0; j < 4; ++j) g4[j].store(dcol + (int64_t)j * HW);
    }
  }
}

template <int BATCH, bool EXACT, bool FAST, bool WRITE, bool PEER = false>   // PEER: see loss_main_kernel
__global__ void __launch_bounds__(kLvBlock, LV_MINB) loss_main_levels_kernel(const LevelLossArgs A, const LevelTable LT) {
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  int l = 0;
  while (l + 1 < LT.num_levels && (int)blockIdx.x >= LT.tile_base[l + 1]) ++l;
  const int local = (int)blockIdx.x - LT.tile_base[l];
  const int a = local / LT.chunks[l];
  const int vec = LT.vec[l];
  const int hw0 = ((local - a * LT.chunks[l]) * kLvBlock + tid) * vec;

  grid_launch_dependents();   // the post pass may be scheduled as this grid's CTAs retire; it waits for the whole grid
  grid_dependency_sync();     // everything K1 produced is read from here on
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

  TileSums S = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (hw0 < LT.HW[l]) {
#if LV_VEC4
    if (vec == 4) level_body<4, BATCH / 2, EXACT, FAST, WRITE>(A, LT, l, a, hw0, n, inv_nf, inv_S, S);
    else
#endif
    if (vec == 2) level_body<2, BATCH, EXACT, FAST, WRITE>(A, LT, l, a, hw0, n, inv_nf, inv_S, S);
    else level_body<1, BATCH, EXACT, FAST, WRITE>(A, LT, l, a, hw0, n, inv_nf, inv_S, S);
  }
  S.cls = warp_sum(S.cls); S.reg = warp_sum(S.reg); S.wl = warp_sum(S.wl);
  S.l = warp_sum(S.l); S.mx = warp_max(S.mx);
  finish_tile<kLvBlock, PEER>(S.cls, S.reg, S.wl, S.l, S.mx, n, blockIdx.x, A.tiles_per_image, A.N, A.partials, A.counter,
                        A.scalars, nf_d, A.c_cls, A.c_reg, A.c_gam, A.peer);
}

// d(c_gam * G)/d bets on the gambler's own layout (see fsg_loss_post for the formula): thread = one (n, a, hw).
struct PostTable {
  const float* bets[kLvMax];
  const float* ell[kLvMax];
  float* grad[kLvMax];
  int64_t off[kLvMax + 1];
  int HW[kLvMax];
  int tile_base[kLvMax + 1];   // tiles of 256 cells (all A anchor slots of a cell in one thread)
  int num_levels, A;
};

__global__ void __launch_bounds__(256) loss_post_levels_kernel(const PostTable PT, const int64_t* __restrict__ mask,
                                                               int64_t R, float T, float ggamma, int nmode, float c_gam,
                                                               const double* __restrict__ stats,
                                                               const double* __restrict__ scalars) {
  // thread = one cell (n, level, hw), all A anchor slots of it: the A mask entries of a cell are adjacent in the
  // (N, R) order, so a warp reads them as one contiguous piece, and each of the A planes is walked coalesced
  const int n = blockIdx.y;
  grid_dependency_sync();   // (programmatic dependent launch behind the main pass)
  int l = 0;
  while (l + 1 < PT.num_levels && (int)blockIdx.x >= PT.tile_base[l + 1]) ++l;
  const int hw = ((int)blockIdx.x - PT.tile_base[l]) * 256 + threadIdx.x;
  if (hw >= PT.HW[l]) return;
  float inv_S = 1.f, Asum = 0.f;
  if (nmode != FSG_NORM_NONE) {
    const double S = (nmode == FSG_NORM_IMAGE) ? stats[FSG_STATS_HEADER + n] : stats[1];
    const double Av = (nmode == FSG_NORM_IMAGE) ? scalars[FSG_SCALARS_HEADER + n] : scalars[2];
    inv_S = __frcp_rn((float)S);
    Asum = (float)Av;
  }
  const int64_t* mrow = mask ? mask + (int64_t)n * R + PT.off[l] + (int64_t)hw * PT.A : nullptr;
  constexpr int kG = 3;   // anchor slots per round: all loads of a round are issued before its arithmetic
  for (int a0 = 0; a0 < PT.A; a0 += kG) {
    float m[kG], b[kG], e[kG];
#pragma unroll
    for (int j = 0; j < kG; ++j) {
      const bool in = a0 + j < PT.A;
      const int64_t i = ((int64_t)n * PT.A + a0 + j) * PT.HW[l] + hw;
      m[j] = (in && mrow) ? (float)mrow[a0 + j] : 1.f;
      b[j] = in ? PT.bets[l][i] : 0.f;
      e[j] = in ? PT.ell[l][i] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kG; ++j) {
      if (a0 + j >= PT.A) break;
      const int64_t i = ((int64_t)n * PT.A + a0 + j) * PT.HW[l] + hw;
      const float w = __fadd_rn(__fmul_rn(b[j], m[j]), T);
      float g;
      if (nmode == FSG_NORM_NONE) {
        const float pw = (ggamma == 1.f) ? 1.f : powf(w, ggamma - 1.f);
        g = -m[j] * ggamma * pw * e[j];
      } else {
        const float w_hat = w * inv_S;
        const float pw = (ggamma == 1.f) ? 1.f : powf(w_hat, ggamma - 1.f);
        g = -(m[j] * inv_S) * ggamma * (pw * e[j] - Asum);
      }
      PT.grad[l][i] = c_gam * g;
    }
  }
}

static int build_table(const fsg_head_level* h, int num_levels, int A, LevelTable* t, int64_t* R_out) {
  if (!h || num_levels <= 0 || num_levels > kLvMax || A <= 0) return FSG_ERR_INVALID_ARG;
  t->num_levels = num_levels;
  t->A = A;
  int64_t off = 0;
  int tiles = 0;
  for (int l = 0; l < kLvMax; ++l) {
    if (l < num_levels) {
      if (h[l].H <= 0 || h[l].W <= 0) return FSG_ERR_INVALID_ARG;
      const int64_t hw = (int64_t)h[l].H * h[l].W;
      if (hw > (1 << 30)) return FSG_ERR_UNSUPPORTED;
      t->logits[l] = h[l].logits; t->grad_logits[l] = h[l].grad_logits;
      t->pred_deltas[l] = h[l].pred_deltas; t->grad_deltas[l] = h[l].grad_deltas;
      t->bets[l] = h[l].bets; t->ell[l] = h[l].per_anchor_loss;
      // two hw per thread when every (channel) plane of the level starts 8-byte aligned
      const bool al8 = (hw % 2 == 0) && !(((uintptr_t)h[l].logits | (uintptr_t)h[l].grad_logits |
                                           (uintptr_t)h[l].pred_deltas | (uintptr_t)h[l].grad_deltas) & 7);
      const bool al16 = LV_VEC4 && (hw % 4 == 0) && !(((uintptr_t)h[l].logits | (uintptr_t)h[l].grad_logits |
                                                         (uintptr_t)h[l].pred_deltas | (uintptr_t)h[l].grad_deltas) & 15);
      t->vec[l] = al16 ? 4 : (al8 ? 2 : 1);
      t->off[l] = off; t->HW[l] = (int)hw; t->chunks[l] = (int)ceil_div(hw, (int64_t)kLvBlock * t->vec[l]);
      t->tile_base[l] = tiles;
      off += hw * A;
      tiles += A * t->chunks[l];
    } else {
      t->logits[l] = nullptr; t->grad_logits[l] = nullptr; t->pred_deltas[l] = nullptr; t->grad_deltas[l] = nullptr;
      t->bets[l] = nullptr; t->ell[l] = nullptr;
      t->off[l] = off; t->HW[l] = 0; t->vec[l] = 1; t->chunks[l] = 1; t->tile_base[l] = tiles;
    }
  }
  t->off[kLvMax] = off;
  t->tile_base[kLvMax] = tiles;
  for (int l = num_levels; l < kLvMax; ++l) { t->off[l] = off; t->tile_base[l] = tiles; }
  *R_out = off;
  return FSG_OK;
}

template <int BATCH, bool EXACT>
static void launch_levels2(bool fast, bool write, dim3 grid, cudaStream_t s, const LevelLossArgs& a,
                           const LevelTable& t, bool pdl) {
  const dim3 blk(kLvBlock);
  if (a.peer.world > 1) {
    if (fast) {
      if (write) launch_pdl(loss_main_levels_kernel<BATCH, EXACT, true, true, true>, grid, blk, 0, s, pdl, a, t);
      else launch_pdl(loss_main_levels_kernel<BATCH, EXACT, true, false, true>, grid, blk, 0, s, pdl, a, t);
    } else {
      if (write) launch_pdl(loss_main_levels_kernel<BATCH, EXACT, false, true, true>, grid, blk, 0, s, pdl, a, t);
      else launch_pdl(loss_main_levels_kernel<BATCH, EXACT, false, false, true>, grid, blk, 0, s, pdl, a, t);
    }
    return;
  }
  if (fast) {
    if (write) launch_pdl(loss_main_levels_kernel<BATCH, EXACT, true, true>, grid, blk, 0, s, pdl, a, t);
    else launch_pdl(loss_main_levels_kernel<BATCH, EXACT, true, false>, grid, blk, 0, s, pdl, a, t);
  } else {
    if (write) launch_pdl(loss_main_levels_kernel<BATCH, EXACT, false, true>, grid, blk, 0, s, pdl, a, t);
    else launch_pdl(loss_main_levels_kernel<BATCH, EXACT, false, false>, grid, blk, 0, s, pdl, a, t);
  }
}
static void launch_levels(int K, bool fast, bool write, dim3 grid, cudaStream_t s, const LevelLossArgs& a,
                          const LevelTable& t, bool pdl) {
#if defined(LV_BATCH16)
  if (K % 16 == 0) launch_levels2<16, true>(fast, write, grid, s, a, t, pdl);
  else if (K % 8 == 0) launch_levels2<8, true>(fast, write, grid, s, a, t, pdl);
  else if (K % 10 == 0) launch_levels2<10, true>(fast, write, grid, s, a, t, pdl);
#elif defined(LV_PREFER10)
  if (K % 10 == 0) launch_levels2<10, true>(fast, write, grid, s, a, t, pdl);
  else if (K % 8 == 0) launch_levels2<8, true>(fast, write, grid, s, a, t, pdl);
#else
  if (K % 8 == 0) launch_levels2<8, true>(fast, write, grid, s, a, t, pdl);      // K = 80: 64 registers, no spills
  else if (K % 10 == 0) launch_levels2<10, true>(fast, write, grid, s, a, t, pdl);  // K = 1230
#endif
  else launch_levels2<8, false>(fast, write, grid, s, a, t, pdl);
}

}  // namespace fsg

using namespace fsg;

extern "C" size_t fsg_loss_main_levels_workspace_bytes(int N, const fsg_head_level* h_levels, int num_levels,
                                                       int A) {
  LevelTable t;
  int64_t R;
  if (N <= 0 || build_table(h_levels, num_levels, A, &t, &R) != FSG_OK) return 0;
  return 16 + align_up(sizeof(float) * kPartialStride * (size_t)N * t.tile_base[kLvMax], 16);
}

namespace fsg {
int loss_main_levels_enqueue(const fsg_head_level* h_levels, int num_levels, int A, const float* gt_deltas,
                             const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                             const int32_t* gt_offsets, const int32_t* matched_idx32,
                             const int64_t* gt_classes, const int64_t* mask, const float* bets, int N,
                             int64_t R, const fsg_loss_params* hp, double* stats,
                             float* per_anchor_loss, float* weights_out, double* scalars, void* workspace,
                             size_t workspace_bytes, const fsg_peer_ctx* h_peer, int flags, fsg_stream_t stream) {
  if (!hp || N <= 0 || R <= 0 || hp->num_classes <= 0) return FSG_ERR_INVALID_ARG;
  if (!gt_classes || !stats || !scalars) return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  if (anchor_image_stride % 4 != 0) return FSG_ERR_INVALID_ARG;
  LevelTable t;
  int64_t Rt = 0;
  const int st = build_table(h_levels, num_levels, A, &t, &Rt);
  if (st != FSG_OK) return st;
  if (Rt != R) return FSG_ERR_INVALID_ARG;
  bool any_pred = false;
  for (int l = 0; l < num_levels; ++l) {
    if (!t.logits[l]) return FSG_ERR_INVALID_ARG;
    if (t.grad_deltas[l] && !t.pred_deltas[l]) return FSG_ERR_INVALID_ARG;
    any_pred = any_pred || t.pred_deltas[l];
  }
  if (any_pred && !gt_deltas && (!anchors || !gt_boxes || !gt_offsets || !matched_idx32)) return FSG_ERR_INVALID_ARG;
  int native_bets = 0;
  for (int l = 0; l < num_levels; ++l) native_bets += t.bets[l] ? 1 : 0;
  if (native_bets != 0 && (native_bets != num_levels || bets)) return FSG_ERR_INVALID_ARG;   // all levels, or flat
  const bool have_bets = bets || native_bets;
  if (!have_bets && (hp->c_gam != 0.f || weights_out)) return FSG_ERR_INVALID_ARG;
  if (hp->gambler_mode != FSG_CLS_FOCAL && hp->gambler_mode != FSG_CLS_SIGMOID) return FSG_ERR_INVALID_ARG;
  if (hp->norm_mode < FSG_NORM_NONE || hp->norm_mode > FSG_NORM_BATCH) return FSG_ERR_INVALID_ARG;
  const size_t need = 16 + align_up(sizeof(float) * kPartialStride * (size_t)N * t.tile_base[kLvMax], 16);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;

  LevelLossArgs a;
  a.gt_deltas = gt_deltas; a.anchors = (const float4*)anchors; a.anchor_stride4 = anchor_image_stride / 4;
  a.gt_boxes = (const float4*)gt_boxes; a.gt_offsets = gt_offsets; a.matched = matched_idx32;
  a.gt_classes = gt_classes; a.mask = mask; a.bets = bets;
  a.N = N; a.R = R; a.K = hp->num_classes; a.tiles_per_image = t.tile_base[kLvMax];
  a.a0 = hp->focal_alpha >= 0.f ? 1.f - hp->focal_alpha : 1.f;
  a.a1 = hp->focal_alpha >= 0.f ? hp->focal_alpha : 1.f;
  a.gamma = hp->focal_gamma; a.beta = hp->smooth_l1_beta; a.T = hp->temperature; a.ggamma = hp->gambler_gamma;
  a.gmode = hp->gambler_mode; a.nmode = have_bets ? hp->norm_mode : FSG_NORM_NONE;
  a.c_cls = hp->c_cls; a.c_reg = hp->c_reg; a.c_gam = hp->c_gam;
  a.wx = hp->box_weights[0]; a.wy = hp->box_weights[1]; a.ww = hp->box_weights[2]; a.wh = hp->box_weights[3];
  a.stats = stats; a.ell = per_anchor_loss; a.wout = weights_out;
  a.partials = (float*)(ws + 16); a.counter = (unsigned*)ws; a.scalars = scalars;
  a.peer = make_peer_poll(h_peer, stats);

  if (!(flags & kLossCounterZeroed)) FSG_CUDA_TRY(cudaMemsetAsync(a.counter, 0, 16, s));
  const bool fast = (hp->focal_gamma == 2.f) && (hp->gambler_mode == FSG_CLS_FOCAL);
  dim3 grid((unsigned)t.tile_base[kLvMax], (unsigned)N);
  bool write = false;
  for (int l = 0; l < num_levels; ++l) write = write || t.grad_logits[l];
  for (int l = 0; l < num_levels; ++l)
    if (write && !t.grad_logits[l]) return FSG_ERR_INVALID_ARG;   // gradients for all levels or for none
  launch_levels(a.K, fast, write, grid, s, a, t, (flags & kLossPdl) != 0);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

size_t loss_main_levels_ws_bytes(int N, const fsg_head_level* h_levels, int num_levels, int A) {
  return fsg_loss_main_levels_workspace_bytes(N, h_levels, num_levels, A);
}
}  // namespace fsg

extern "C" int fsg_loss_main_levels(const fsg_head_level* h_levels, int num_levels, int A, const float* gt_deltas,
                                    const float* anchors, int64_t anchor_image_stride, const float* gt_boxes,
                                    const int32_t* gt_offsets, const int32_t* matched_idx32,
                                    const int64_t* gt_classes, const int64_t* mask, const float* bets, int N,
                                    int64_t R, const fsg_loss_params* hp, const double* stats,
                                    float* per_anchor_loss, float* weights_out, double* scalars, void* workspace,
                                    size_t workspace_bytes, fsg_stream_t stream) {
  return loss_main_levels_enqueue(h_levels, num_levels, A, gt_deltas, anchors, anchor_image_stride, gt_boxes,
                                  gt_offsets, matched_idx32, gt_classes, mask, bets, N, R, hp,
                                  const_cast<double*>(stats), per_anchor_loss, weights_out, scalars, workspace,
                                  workspace_bytes, nullptr, 0, stream);
}

namespace fsg {
int loss_post_levels_enqueue(const fsg_post_level* h_levels, int num_levels, int A, const int64_t* mask, int N,
                             int64_t R, const fsg_loss_params* hp, const double* stats, const double* scalars,
                             int flags, fsg_stream_t stream) {
  if (!hp || !h_levels || num_levels <= 0 || num_levels > kLvMax || A <= 0 || N <= 0 || !stats || !scalars)
    return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  PostTable t;
  int64_t off = 0;
  int tiles = 0;
  for (int l = 0; l < kLvMax; ++l) {
    t.off[l] = off;
    t.tile_base[l] = tiles;
    if (l < num_levels) {
      const int64_t hw = (int64_t)h_levels[l].H * h_levels[l].W;
      if (h_levels[l].H <= 0 || h_levels[l].W <= 0 || hw > (1 << 30)) return FSG_ERR_INVALID_ARG;
      if (!h_levels[l].bets || !h_levels[l].per_anchor_loss || !h_levels[l].grad_bets) return FSG_ERR_INVALID_ARG;
      t.bets[l] = h_levels[l].bets; t.ell[l] = h_levels[l].per_anchor_loss; t.grad[l] = h_levels[l].grad_bets;
      t.HW[l] = (int)hw;
      off += hw * A;
      tiles += (int)ceil_div(hw, 256);
    } else {
      t.bets[l] = nullptr; t.ell[l] = nullptr; t.grad[l] = nullptr; t.HW[l] = 0;
    }
  }
  t.off[kLvMax] = off; t.tile_base[kLvMax] = tiles;
  t.num_levels = num_levels; t.A = A;
  if (off != R) return FSG_ERR_INVALID_ARG;
  launch_pdl(loss_post_levels_kernel, dim3((unsigned)tiles, (unsigned)N), dim3(256), 0, (cudaStream_t)stream,
             (flags & kLossPdl) != 0, t, mask, R, hp->temperature, hp->gambler_gamma, (int)hp->norm_mode, hp->c_gam,
             stats, scalars);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
}  // namespace fsg

extern "C" int fsg_loss_post_levels(const fsg_post_level* h_levels, int num_levels, int A, const int64_t* mask, int N,
                                    int64_t R, const fsg_loss_params* hp, const double* stats, const double* scalars,
                                    fsg_stream_t stream) {
  return loss_post_levels_enqueue(h_levels, num_levels, A, mask, N, R, hp, stats, scalars, 0, stream);
}
