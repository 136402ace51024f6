// K1: pairwise IoU, Matcher, and the fused IoU+Matcher+GT-assignment path (sm_100a).
//
// Reference semantics (paths relative to the reference tree):
//   detectron2/structures/boxes.py:243-275   pairwise_iou (fp32 op order reproduced exactly:
//                                            no FMA contraction, IEEE division)
//   detectron2/modeling/matcher.py:55-132    Matcher.__call__ / set_low_quality_matches_
//   detectron2/modeling/box_regression.py:34-67  get_deltas
//   detectron2/modeling/meta_arch/retinanet.py:339-363, 400-425  GT assignment + picky mask
//
// Layout: anchors (R,4) float4-aligned, streamed with one coalesced 16-byte load per anchor;
// the image's GT boxes are staged into shared memory by a TMA bulk copy (cp.async.bulk +
// mbarrier); every thread owns U anchors and walks the GT table with shared-memory broadcast
// reads; per-GT maxima are reduced with redux.sync and merged with atomicMax on the
// float-as-uint pattern (valid because IoU >= 0).
#include <stdlib.h>

#include "common.cuh"
#include "step_internal.cuh"

namespace fsg {

constexpr int kMaxThr = 4;       // thresholds per matcher
constexpr int kMatchBlock = 256;
constexpr int kGtChunk = 1024;   // GT boxes staged per shared-memory chunk

struct MatcherBands {
  float lo[kMaxThr + 1];
  float hi[kMaxThr + 1];
  int8_t lab[kMaxThr + 1];
  int n;  // number of bands (= thresholds + 1); 0 = matcher disabled
};

__device__ __forceinline__ float box_area(float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));  // boxes.py:119
}

// boxes.py:261-274, b1 = ground truth (area1), b2 = anchor (area2)
__device__ __forceinline__ float iou_exact(float4 a, float area_a, float4 b, float area_b) {
  float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  float inter = __fmul_rn(w, h);
  float r = 0.f;
  if (inter > 0.f) r = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return r;
}

// intersection area and the union denominator of boxes.py:261-272 (same op order as iou_exact)
__device__ __forceinline__ float inter_area(float4 a, float4 b) {
  float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  return __fmul_rn(w, h);
}
// Conservative pre-test for "fl(inter/uni) >= bound": a correctly rounded quotient q satisfies
// inter/uni >= q(1-2^-24), so q >= bound implies inter >= bound*uni*(1-2^-24); the factor 1-2^-21 below
// leaves room for the two roundings of the product.  Never rejects a pair that could reach `bound`;
// pairs it rejects are strictly below it, so skipping their IEEE division cannot change any max/argmax.
__device__ __forceinline__ bool may_reach(float inter, float uni, float bound) {
  return inter >= bound * uni * 0.99999952f;
}

// Pass A inner step for one (GT, anchor) pair.  Non-overlapping pairs (the vast majority) cost four
// min/max, two subtractions and two compares; the clamp-to-zero of boxes.py:265 is implied by the
// "both extents positive" test (a clamped extent of 0 gives inter == 0 -> IoU exactly 0, which can never
// beat a running best that starts at 0 nor raise a per-GT maximum).
__device__ __forceinline__ void pair_update(float4 G, float ga, float4 a, float aa, float known, int gidx,
                                            float& bv, int& bi, float& m, int& ov) {
  const float w = __fsub_rn(fminf(G.z, a.z), fmaxf(G.x, a.x));
  const float h = __fsub_rn(fminf(G.w, a.w), fmaxf(G.y, a.y));
  if (w > 0.f && h > 0.f) {
    const float inter = __fmul_rn(w, h);
    if (inter > 0.f) {   // the product of two tiny positives can underflow to 0
      const float uni = __fsub_rn(__fadd_rn(ga, aa), inter);
      // (a pair the filter rejects is strictly below both the anchor's best and the GT's running maximum)
      if (may_reach(inter, uni, fminf(bv, known))) {
        const float v = __fdiv_rn(inter, uni);
        // ov bit 0: the pair that holds the anchor's best IoU reached its GT's running maximum when it was evaluated;
        // bit 1: some OTHER evaluated pair of the anchor did.  IoU(g, a) can equal the final maximum of g only if it
        // reached the running one, so bit 1 (ov >= 2) tells pass B whether a GT other than the argmax can promote the
        // anchor at all -- for nearly every anchor it cannot, and pass B is done after the argmax test.  (Config 2:
        // no difference to flagging "overlaps two or more GT"; a crowded image in the batch: fewer warp-serial walks.)
        const int top = (v >= known) ? 1 : 0;
        if (v > bv) { bv = v; bi = gidx; ov = ((ov & 1) << 1) | (ov & 2) | top; }
        else ov |= top << 1;
        m = fmaxf(m, v);
      }
    }
  }
}

__device__ __forceinline__ int8_t band_label(const MatcherBands& mb, float v) {
  int8_t l = 1;  // matcher.py:88
#pragma unroll
  for (int i = 0; i <= kMaxThr; ++i)
    if (i < mb.n && v >= mb.lo[i] && v < mb.hi[i]) l = mb.lab[i];
  return l;
}

// The same bands in registers: thresholds ascending, so the band index is the number of thresholds <= v and
// the label comes out of a byte-packed word.  (For a NaN quality the reference keeps label 1, matcher.py:88;
// IoU produced by this library is never NaN.)
struct BandsReg {
  float t[kMaxThr];          // thresholds; +inf beyond the last one
  unsigned long long packed; // labels of bands 0..kMaxThr, one byte each
};
__host__ __device__ inline BandsReg make_bands_reg(const MatcherBands& mb) {
  BandsReg b;
  b.packed = 0ull;
#pragma unroll
  for (int i = 0; i < kMaxThr; ++i) b.t[i] = (i + 1 < mb.n) ? mb.hi[i] : __builtin_huge_valf();
#pragma unroll
  for (int i = 0; i <= kMaxThr; ++i)
    b.packed |= (unsigned long long)(unsigned char)((i < mb.n) ? mb.lab[i] : (int8_t)1) << (8 * i);
  return b;
}
__device__ __forceinline__ int8_t band_label_reg(const BandsReg& b, float v) {
  int idx = 0;
#pragma unroll
  for (int i = 0; i < kMaxThr; ++i) idx += (v >= b.t[i]) ? 1 : 0;
  return (int8_t)(unsigned char)(b.packed >> (8 * idx));
}

// box_regression.py:49-63 (mul-then-add order kept; 0.5*w is exact so the fused form is identical)
__device__ __forceinline__ float4 encode_deltas(float4 s, float4 t, float wx, float wy, float ww, float wh) {
  float sw = __fsub_rn(s.z, s.x), sh = __fsub_rn(s.w, s.y);
  float sx = __fadd_rn(s.x, __fmul_rn(0.5f, sw)), sy = __fadd_rn(s.y, __fmul_rn(0.5f, sh));
  float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
  float tx = __fadd_rn(t.x, __fmul_rn(0.5f, tw)), ty = __fadd_rn(t.y, __fmul_rn(0.5f, th));
  float4 d;
  d.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tx, sx)), sw);
  d.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(ty, sy)), sh);
  d.z = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
  d.w = __fmul_rn(wh, logf(__fdiv_rn(th, sh)));
  return d;
}

// ------------------------------------------------------------------------------------------
// pairwise_iou drop-in: materialises the (n1, n2) matrix
// ------------------------------------------------------------------------------------------
constexpr int kIouRows = 8;
__global__ void __launch_bounds__(256) pairwise_iou_kernel(const float4* __restrict__ b1, int64_t n1,
                                                           const float4* __restrict__ b2, int64_t n2,
                                                           float* __restrict__ out) {
  __shared__ float4 rows[kIouRows];
  __shared__ float areas[kIouRows];
  const int64_t i0 = (int64_t)blockIdx.y * kIouRows;
  if (threadIdx.x < kIouRows && i0 + threadIdx.x < n1) {
    float4 r = b1[i0 + threadIdx.x];
    rows[threadIdx.x] = r;
    areas[threadIdx.x] = box_area(r);
  }
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n2) return;
  const float4 c = b2[j];
  const float ac = box_area(c);
#pragma unroll
  for (int k = 0; k < kIouRows; ++k) {
    if (i0 + k < n1) out[(i0 + k) * n2 + j] = iou_exact(rows[k], areas[k], c, ac);
  }
}

// ------------------------------------------------------------------------------------------
// Matcher drop-in on a materialised matrix
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matrix_rowmax_kernel(const float* __restrict__ q, int64_t M, int64_t N,
                                                            unsigned* __restrict__ rowmax) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g0 = (int64_t)blockIdx.y * 64;
  const int64_t g1 = min(M, g0 + 64);
  for (int64_t g = g0; g < g1; ++g) {
    float v = (n < N) ? q[g * N + n] : 0.f;
    unsigned m = __reduce_max_sync(kFull, __float_as_uint(v));
    if ((threadIdx.x & 31) == 0 && m > 0u) atomicMax(&rowmax[g], m);
  }
}

__global__ void __launch_bounds__(256) matrix_match_kernel(const float* __restrict__ q, int64_t M, int64_t N,
                                                           MatcherBands mb, int allow_lq,
                                                           const float* __restrict__ rowmax,
                                                           int64_t* __restrict__ matches,
                                                           int8_t* __restrict__ labels) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (M == 0) {  // matcher.py:70-80
    matches[n] = 0;
    labels[n] = mb.lab[0];
    return;
  }
  float best = q[n];
  int64_t bi = 0;
  bool lq = allow_lq && (best == rowmax[0]);
  for (int64_t g = 1; g < M; ++g) {
    float v = q[g * N + n];
    if (v > best) { best = v; bi = g; }
    if (allow_lq && v == rowmax[g]) lq = true;
  }
  matches[n] = bi;
  labels[n] = lq ? (int8_t)1 : band_label(mb, best);
}

// ------------------------------------------------------------------------------------------
// Fused path, pass A: per-anchor max/argmax over the image's GT + per-GT max over anchors
// ------------------------------------------------------------------------------------------
constexpr int kMultiOverlap = 1 << 30;   // best_idx flag: a GT other than the argmax may promote the anchor (pair_update)
constexpr int kSmallM = 32;  // images with at most this many GT: boxes staged by one parallel load, warp-level culling

struct MatchOut {
  int64_t* matches;
  int8_t* match_labels;
  int8_t* picky_labels;
  int64_t* gt_classes;
  int64_t* mask;
  float4* gt_deltas;
  int32_t* matched_idx32;
};

// the gambler's betting maps in their own per-level (N, A, H, W) layout (read in place)
struct BetLevels {
  const float* ptr[FSG_MAX_LEVELS];
  int64_t off[FSG_MAX_LEVELS + 1];
  int HW[FSG_MAX_LEVELS];
  int A, num_levels;
};
__device__ __forceinline__ float bet_at(const BetLevels& lv, int n, int64_t r) {
  int l = 0;
  while (l + 1 < lv.num_levels && r >= lv.off[l + 1]) ++l;
  const int local = (int)(r - lv.off[l]);
  const int hw = local / lv.A;
  const int a = local - hw * lv.A;
  return lv.ptr[l][((int64_t)n * lv.A + a) * lv.HW[l] + hw];
}

constexpr int kWarpsPerBlock = kMatchBlock / 32;

// Every thread of pass A has seen ALL ground truth of its image when its loop ends, so its best IoU / argmax are
// final and the threshold bands (matcher.py:88-92), the class relabel (retinanet.py:354-360), the picky mask
// (:417-425), get_deltas and the loss pre-pass sums can be produced right there.  What pass A cannot know is the
// low-quality rule (matcher.py:99-132: needs the per-GT maxima over all anchors): pass B revisits only the anchors
// that can be affected and patches their labels and the sums.
struct PassAEpi {
  const int64_t* gt_class_ids;
  int num_classes;
  int8_t lab0, plab0;        // labels of an image without ground truth (matcher.py:70-80)
  int has_picky;
  BandsReg br, pbr;
  float wx, wy, ww, wh;
  MatchOut out;
  const float* bets;
  float temperature;
  int* part_cnt;             // (N, gridDim.x) or NULL: no pre-pass sums
  float* part_s;
};

// order-preserving float <-> int (its own inverse), so that redux.sync reduces a float min/max in one instruction
__device__ __forceinline__ int float_key(float f) {
  const int k = __float_as_int(f);
  return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

// STEP: the output set of the fused training step (gt_classes, mask, matched_idx32; class ids given) known at compile
// time -- the per-anchor tests of seven output pointers were 11 % of the kernel's instructions.
template <int U, bool STEP>
__global__ void __launch_bounds__(kMatchBlock, 4) match_pass_a_kernel(
    const float4* __restrict__ anchors, int64_t R, int64_t anchor_stride4,
    const float4* __restrict__ gt_boxes, const int32_t* __restrict__ gt_offsets,
    float* __restrict__ best_val, int32_t* __restrict__ best_idx, unsigned* __restrict__ gt_max,
    const PassAEpi E, const BetLevels lv) {
  __shared__ __align__(16) float4 s_gt[kGtChunk];
  __shared__ float s_area[kGtChunk];
  __shared__ unsigned s_max[kGtChunk];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int64_t s_cls[kSmallM];
  grid_launch_dependents();   // the next kernel may be scheduled as this grid's CTAs retire (it waits for the whole grid)

  const int n = blockIdx.y;
  const int m0 = gt_offsets[n];
  const int M = gt_offsets[n + 1] - m0;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int64_t base = (int64_t)blockIdx.x * (kMatchBlock * U);
  const float4* a_img = anchors + (int64_t)n * anchor_stride4;

  float4 a[U];
  float aa[U], bv[U], bet[U];
  int bi[U], ov[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    ov[u] = 0;
    int64_t r = base + u * kMatchBlock + tid;
    a[u] = (r < R) ? a_img[r] : make_float4(0.f, 0.f, 0.f, 0.f);
    // the bet is only needed by the epilogue; loading it now hides its HBM latency behind the GT loop
    bet[u] = (E.bets != nullptr && r < R) ? E.bets[(int64_t)n * R + r] : 0.f;
    aa[u] = box_area(a[u]);
    bv[u] = 0.f;   // running best starts at (IoU 0, GT 0): exactly torch's argmax of an all-zero column,
    bi[u] = 0;     // and strict '>' keeps the lowest GT index among ties (matcher.py:86)
  }

  unsigned live = 0xffffffffu;   // few-GT images: the GT whose box reaches into this warp's anchors
  const bool small = M <= kSmallM;
  if (small) {
    // ---- few GT (the detection-training case).  Lane g holds GT g (one parallel load; M broadcast loads inside
    //      the loop would each expose a full L2 latency) and hands it out by shuffles; per-GT maxima are merged per
    //      CTA in shared memory first, because same-address atomics serialise in L2 (~27 cycles each).
    const bool has_g = lane < M;
    float4 myG = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_g) myG = gt_boxes[m0 + lane];
    if (tid < kSmallM) {
      s_max[tid] = 0u;
      if (tid < M) s_cls[tid] = E.gt_class_ids ? E.gt_class_ids[m0 + tid] : 0;
    }
    // warp-level culling: consecutive anchors are neighbouring cells of one pyramid level, so the warp's anchors
    // span a small rectangle, and a ground truth that does not reach into it has IoU exactly 0 with every lane --
    // it can neither beat a running best (which starts at 0) nor raise a per-GT maximum.  Exact, not a heuristic.
    // 32 GT are tested at once (one per lane) and only the survivors are walked.
    int kx0 = 0x7fffffff, ky0 = 0x7fffffff, kx1 = (int)0x80000000, ky1 = (int)0x80000000;
  #pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * kMatchBlock + tid < R) {
        kx0 = min(kx0, float_key(a[u].x)); ky0 = min(ky0, float_key(a[u].y));
        kx1 = max(kx1, float_key(a[u].z)); ky1 = max(ky1, float_key(a[u].w));
      }
    }
    const float bx0 = key_float(__reduce_min_sync(kFull, kx0)), by0 = key_float(__reduce_min_sync(kFull, ky0));
    const float bx1 = key_float(__reduce_max_sync(kFull, kx1)), by1 = key_float(__reduce_max_sync(kFull, ky1));
    // (NaN never culls; a warp without a valid anchor has a NaN rectangle and culls nothing, harmlessly)
    const float my_ga = box_area(myG);
    const bool hit = has_g && !(myG.z <= bx0 || myG.x >= bx1 || myG.w <= by0 || myG.y >= by1);
    live = __ballot_sync(kFull, hit);
    __syncthreads();
    for (unsigned rem = live; rem != 0u; rem &= rem - 1u) {   // ascending g: ties keep the lowest GT index
      const int g = __ffs(rem) - 1;
      float4 G;
      G.x = __shfl_sync(kFull, myG.x, g); G.y = __shfl_sync(kFull, myG.y, g);
      G.z = __shfl_sync(kFull, myG.z, g); G.w = __shfl_sync(kFull, myG.w, g);
      const float ga = __shfl_sync(kFull, my_ga, g);
      const float known = __uint_as_float(s_max[g]);  // what this CTA found so far: only makes the filter looser
      float m = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) pair_update(G, ga, a[u], aa[u], known, g, bv[u], bi[u], m, ov[u]);
      if (__any_sync(kFull, m > known)) {
        const unsigned wm = __reduce_max_sync(kFull, __float_as_uint(m));
        if (lane == 0) atomicMax(&s_max[g], wm);
      }
    }
    __syncthreads();
    if (tid < M) {
      const unsigned v = s_max[tid];
      if (v > 0u) atomicMax(&gt_max[m0 + tid], v);
    }
  } else {
    // ---- a crowded image inside a few-GT batch (a whole crowded batch goes to match_pass_a_crowded_kernel): detection
    //      training, i.e. anchors on a regular grid, so the warp-level cull of the few-GT path applies here as well --
    //      a staged GT that does not reach into the rectangle of the warp's anchors is skipped by the whole warp (exact:
    //      its IoU is 0 with every lane).  At 33..100 GT on an 800 x 1333 image that leaves a handful per warp.
    int kx0 = 0x7fffffff, ky0 = 0x7fffffff, kx1 = (int)0x80000000, ky1 = (int)0x80000000;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * kMatchBlock + tid < R) {
        kx0 = min(kx0, float_key(a[u].x)); ky0 = min(ky0, float_key(a[u].y));
        kx1 = max(kx1, float_key(a[u].z)); ky1 = max(ky1, float_key(a[u].w));
      }
    }
    const float bx0 = key_float(__reduce_min_sync(kFull, kx0)), by0 = key_float(__reduce_min_sync(kFull, ky0));
    const float bx1 = key_float(__reduce_max_sync(kFull, kx1)), by1 = key_float(__reduce_max_sync(kFull, ky1));
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int c = 0; c < M; c += kGtChunk) {
      const int cnt = min(kGtChunk, M - c);
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&s_bar, (uint32_t)cnt * 16u);
        tma_bulk_g2s(s_gt, gt_boxes + m0 + c, (uint32_t)cnt * 16u, &s_bar);
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
      for (int g = tid; g < cnt; g += kMatchBlock) {
        s_area[g] = box_area(s_gt[g]);
        s_max[g] = gt_max[m0 + c + g];   // what other CTAs found so far (stale is fine)
      }
      __syncthreads();

      for (int g = 0; g < cnt; ++g) {
        const float4 G = s_gt[g];
        if (G.z <= bx0 || G.x >= bx1 || G.w <= by0 || G.y >= by1) continue;   // warp-uniform; NaN never culls
        const float ga = s_area[g];
        const float known = __uint_as_float(s_max[g]);
        float m = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) pair_update(G, ga, a[u], aa[u], known, c + g, bv[u], bi[u], m, ov[u]);
        // padding lanes (r >= R) hold a zero box: no overlap, harmless for the max
        if (__any_sync(kFull, m > known)) {
          const unsigned wm = __reduce_max_sync(kFull, __float_as_uint(m));
          if (lane == 0) atomicMax(&s_max[g], wm);
        }
      }
      __syncthreads();
      for (int g = tid; g < cnt; g += kMatchBlock) {
        const unsigned v = s_max[g];
        if (v > 0u) atomicMax(&gt_max[m0 + c + g], v);
      }
      __syncthreads();  // s_gt / s_max are rewritten by the next chunk
    }
  }


  // ---- epilogue: everything that depends only on this anchor's own best IoU
  const bool need_deltas = !STEP && (E.out.gt_deltas != nullptr);
  const bool o_matches = !STEP && E.out.matches != nullptr, o_labels = !STEP && E.out.match_labels != nullptr;
  const bool o_picky = !STEP && E.out.picky_labels != nullptr, o_cls = STEP || E.out.gt_classes != nullptr;
  const bool o_mask = STEP || E.out.mask != nullptr, o_idx32 = STEP || E.out.matched_idx32 != nullptr;
  const bool has_ids = STEP || E.gt_class_ids != nullptr, has_bets = E.bets != nullptr;
  int fg = 0;
  float w_part = 0.f;
  int8_t l1 = 0, l2 = 0;
  int64_t cls = 0, msk = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t r = base + u * kMatchBlock + tid;
    if (r >= R) continue;
    const int64_t o = (int64_t)n * R + r;
    const float val = (M > 0) ? bv[u] : 0.f;
    best_val[o] = val;
    best_idx[o] = bi[u] | (ov[u] >= 2 ? kMultiOverlap : 0);
    int id = bi[u];
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (M > 0) {
      // a warp no ground truth reaches into (most of them) has IoU 0 / argmax 0 everywhere: classify once
      if (u == 0 || live != 0u) {
        l1 = band_label_reg(E.br, val);
        l2 = E.has_picky ? band_label_reg(E.pbr, val) : (int8_t)0;
        cls = small ? s_cls[id] : (has_ids ? E.gt_class_ids[m0 + id] : 0);
        if (l1 == 0) cls = E.num_classes;   // retinanet.py:356
        if (l1 == -1) cls = -1;             // :360
        msk = (l2 == 1) ? 1 : 0;            // :417-423
      }
      if (need_deltas) d = encode_deltas(a[u], gt_boxes[m0 + id], E.wx, E.wy, E.ww, E.wh);
    } else {                              // matcher.py:70-80, retinanet.py:362-363, :425
      l1 = E.lab0;
      l2 = E.plab0;
      cls = E.num_classes;
      msk = E.num_classes;
      id = 0;
    }
    if (o_matches) E.out.matches[o] = id;
    if (o_labels) E.out.match_labels[o] = l1;
    if (o_picky) E.out.picky_labels[o] = l2;
    if (o_cls) E.out.gt_classes[o] = cls;
    if (o_mask) E.out.mask[o] = msk;
    if (need_deltas) E.out.gt_deltas[o] = d;
    if (o_idx32) E.out.matched_idx32[o] = id;
    fg += (cls >= 0 && cls != E.num_classes) ? 1 : 0;
    if (has_bets) w_part += __fadd_rn(__fmul_rn(bet[u], (float)msk), E.temperature);  // gambler_heads.py:569,304
    else if (lv.num_levels > 0 && msk != 0) w_part += __fmul_rn(bet_at(lv, n, r), (float)msk);  // + R*T at the fold
  }
  // one partial per WARP in a fixed slot (no barrier, no serial tail in any of the 2000 CTAs); the fold kernel sums
  // the slots in a fixed order => run-to-run deterministic
  if (E.part_cnt != nullptr) {
    const int cw = __reduce_add_sync(kFull, fg);
    const float sw = warp_sum(w_part);
    if (lane == 0) {
      const int64_t slot = ((int64_t)n * gridDim.x + blockIdx.x) * kWarpsPerBlock + (tid >> 5);
      E.part_cnt[slot] = cw;
      E.part_s[slot] = sw;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Pass A for CROWDED batches (more than kSmallM ground truth per image on average; BASELINE config 5: 200 GT x 1 M
// free-form anchors).  Same results as match_pass_a_kernel's generic branch, bit for bit; different schedule.
//
// The brute-force loop spends its time in divergence: ~6 % of the (anchor, GT) pairs overlap, so with 32 lanes nearly
// every warp-level step enters the expensive branch (exact IEEE division, max / argmax, per-GT maximum) for one or two
// active lanes -- 33 issued instructions per pair where a non-overlapping pair needs under ten.  Here the two kinds of
// work are separated:
//   * screen: one anchor per lane against the staged GT, a superset test of "both extents positive"
//     (fl(p - q) > 0 <=> p > q without flush-to-zero) done in packed half precision with outward rounding -- two
//     2-wide compares per pair -- and hits are appended to the lane's own queue column in shared memory (ascending
//     GT index);
//   * drain: when some lane's column is nearly full (and after the last GT) every lane takes its pending pairs in
//     FIFO order and does the exact arithmetic of pair_update -- dense: all lanes with work run the same code.
// Ascending GT order per anchor is preserved, so strict '>' keeps the lowest GT index among ties (matcher.py:86).
// ------------------------------------------------------------------------------------------
// The queue is written (screen_push) and read through explicit 32-bit shared-window addresses: nvcc re-derives the
// window base (S2R CgaCtaId, LEA) inside every predicated store otherwise, which doubles the cost of a push.  Both
// accessors are volatile asm, so they keep their program order with respect to each other.
__device__ __forceinline__ int queue_at(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (int)v;
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t opaque(uint32_t v) {   // keeps nvcc from re-deriving a shared-window base per use
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void reds_max_u32(uint32_t addr, unsigned v) {
  asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ---- the screen in packed half precision.  A box is kept as two f16x2 words, lo = (x1, y1) rounded DOWN and
// hi = (x2, y2) rounded UP, so that "G.hi > a.lo && a.hi > G.lo" (two HSETP2 + one PLOP3 per pair, on the fp16 pipe
// instead of four FSETP on the ALU pipe that bounds this kernel) holds whenever the fp32 test does: a superset, and
// the drain repeats the test exactly.  Out-of-range coordinates saturate towards the permissive side (+-inf / +-65504).
__device__ __forceinline__ uint32_t pack_half2_rd(float x, float y) {
  unsigned short hx, hy;
  asm("cvt.rm.f16.f32 %0, %1;" : "=h"(hx) : "f"(x));
  asm("cvt.rm.f16.f32 %0, %1;" : "=h"(hy) : "f"(y));
  return (uint32_t)hx | ((uint32_t)hy << 16);
}
__device__ __forceinline__ uint32_t pack_half2_ru(float x, float y) {
  unsigned short hx, hy;
  asm("cvt.rp.f16.f32 %0, %1;" : "=h"(hx) : "f"(x));
  asm("cvt.rp.f16.f32 %0, %1;" : "=h"(hy) : "f"(y));
  return (uint32_t)hx | ((uint32_t)hy << 16);
}
// if the boxes (glo, ghi) and (alo, ahi) may overlap: append entry `g` to the queue column at qp and advance it
__device__ __forceinline__ void screen_push(uint32_t& qp, int g, uint32_t glo, uint32_t ghi, uint32_t alo, uint32_t ahi,
                                            uint32_t slot_bytes) {
  asm volatile(
      "{\n\t.reg .pred p, q, r, s;\n\t"
      "setp.gt.f16x2 p|q, %2, %3;\n\t"
      "setp.gt.and.f16x2 r|s, %4, %5, p;\n\t"
      "and.pred r, r, s;\n\t"
      "and.pred r, r, q;\n\t"
      "@r st.shared.u16 [%0], %1;\n\t"
      "@r add.u32 %0, %0, %6;\n\t}"
      : "+r"(qp)
      : "h"((unsigned short)g), "r"(ghi), "r"(alo), "r"(ahi), "r"(glo), "r"(slot_bytes));
}
// (the screen's loads: no "memory" clobber, so that a group of them can be scheduled together; volatile keeps them
//  behind the barrier that follows the staging of the chunk)
__device__ __forceinline__ uint4 lds_u4_screen(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

constexpr int kCrU = 4;      // anchors per thread, one after the other (same grid as match_pass_a_kernel<4, *>)
constexpr int kCrQ = 48;     // pending pairs per lane
constexpr int kCrChunk = 512;   // GT staged per shared-memory chunk
constexpr int kCrStep = 8;   // GT screened between two fill checks (kCrChunk % kCrStep == 0)

template <bool STEP>
__global__ void __launch_bounds__(kMatchBlock, 4) match_pass_a_crowded_kernel(
    const float4* __restrict__ anchors, int64_t R, int64_t anchor_stride4,
    const float4* __restrict__ gt_boxes, const int32_t* __restrict__ gt_offsets,
    float* __restrict__ best_val, int32_t* __restrict__ best_idx, unsigned* __restrict__ gt_max,
    const PassAEpi E, const BetLevels lv) {
  static_assert(kCrChunk % kCrStep == 0 && kCrChunk <= 65536, "queue entries are 16-bit chunk-local GT indices");
  __shared__ __align__(16) float4 s_gt[kCrChunk];
  __shared__ float s_area[kCrChunk];
  __shared__ unsigned s_max[kCrChunk];
  __shared__ unsigned short s_q[kCrQ * kMatchBlock];   // [slot][thread]: a lane's column, conflict-free per slot
  __shared__ __align__(16) uint2 s_gth[kCrChunk];      // the staged GT for the screen: (lo, hi) f16x2 words
  __shared__ __align__(8) uint64_t s_bar;
  grid_launch_dependents();

  // CTAs are dealt out round-robin over the images (the hardware starts them in linear order): the CTAs that run
  // first -- with no per-GT maxima to start from, so that every pair looks like a new maximum -- are then a thin
  // slice of every image instead of most of the first one
  const unsigned linear = blockIdx.y * gridDim.x + blockIdx.x;
  const int n = (int)(linear % gridDim.y);
  const unsigned bx = linear / gridDim.y;
  const int m0 = gt_offsets[n];
  const int M = gt_offsets[n + 1] - m0;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int64_t base = (int64_t)bx * (kMatchBlock * kCrU);
  const float4* a_img = anchors + (int64_t)n * anchor_stride4;
  const float inf = __int_as_float(0x7f800000);

  float4 a[kCrU];
  float bv[kCrU];
  int bi[kCrU];
  // bit 0: the pair that currently holds the anchor's best IoU reached its GT's running maximum when it was evaluated;
  // bit 1: some OTHER evaluated pair of the anchor did.  IoU(g, a) can equal the final maximum of g only if it reached
  // the running one, so bit 1 is what pass B needs to know whether a GT other than the argmax can promote the anchor
  // (it takes the place of the "overlaps two or more GT" flag of match_pass_a_kernel: same meaning, far fewer hits).
  int tops[kCrU];
#pragma unroll
  for (int u = 0; u < kCrU; ++u) {
    const int64_t r = base + u * kMatchBlock + tid;
    // a lane without an anchor screens an inverted box: no ground truth passes
    a[u] = (r < R) ? a_img[r] : make_float4(inf, inf, -inf, -inf);
    bv[u] = 0.f;   // (IoU 0, GT 0): torch's argmax of an all-zero column
    bi[u] = 0;
    tops[u] = 0;
  }

  // ---- every lane takes its own anchors in the order of their expected number of overlapping GT.  A drain round lasts
  //      as long as its busiest lane, and the per-anchor pair counts are spread widely (0 to 50 on config 5, mean 12.6:
  //      an anchor near the border of the GT extent overlaps almost nothing, a large one in the middle a quarter of the
  //      GT).  With every lane's largest first, its smallest last, a round holds anchors of more similar reach (drain
  //      rounds per warp 120 -> 97 in a simulation of config 5).  The estimate: the anchor grown by half the mean GT
  //      size and clipped to the extent of the GT.  Results do not depend on the order.
  int place[kCrU];   // which of the lane's anchors (u of the coalesced load) is processed in round u
  {
    float x0 = inf, y0 = inf, x1 = -inf, y1 = -inf, sw = 0.f, sh = 0.f;
    const int ms = min(M, 256);
    for (int g = lane; g < ms; g += 32) {
      const float4 G = gt_boxes[m0 + g];
      x0 = fminf(x0, G.x); y0 = fminf(y0, G.y); x1 = fmaxf(x1, G.z); y1 = fmaxf(y1, G.w);
      sw += G.z - G.x; sh += G.w - G.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x0 = fminf(x0, __shfl_xor_sync(kFull, x0, o)); y0 = fminf(y0, __shfl_xor_sync(kFull, y0, o));
      x1 = fmaxf(x1, __shfl_xor_sync(kFull, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(kFull, y1, o));
    }
    const float hw = 0.5f * warp_sum(sw) / (float)max(ms, 1), hh = 0.5f * warp_sum(sh) / (float)max(ms, 1);
    float key[kCrU];
#pragma unroll
    for (int u = 0; u < kCrU; ++u) {
      place[u] = u;
      const float rx = fmaxf(fminf(a[u].z + hw, x1) - fmaxf(a[u].x - hw, x0), 0.f);
      const float ry = fmaxf(fminf(a[u].w + hh, y1) - fmaxf(a[u].y - hh, y0), 0.f);
      key[u] = rx * ry;
    }
    auto order = [&](int i, int j) {   // compare-exchange: the larger key first (any outcome is a valid order)
      if (key[j] > key[i]) {
        const float kf = key[i]; key[i] = key[j]; key[j] = kf;
        const float4 af = a[i]; a[i] = a[j]; a[j] = af;
        const int pf = place[i]; place[i] = place[j]; place[j] = pf;
      }
    };
    static_assert(kCrU == 4, "sorting network for 4");
    order(0, 1); order(2, 3); order(0, 2); order(1, 3); order(1, 2);
  }

  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t q0 = opaque((uint32_t)__cvta_generic_to_shared(s_q + tid));   // the lane's queue column
  constexpr uint32_t kSlot = kMatchBlock * 2;                                      // bytes between two slots of a column
  const uint32_t gt0 = opaque((uint32_t)__cvta_generic_to_shared(s_gt));
  const uint32_t gth0 = opaque((uint32_t)__cvta_generic_to_shared(s_gth));
  const uint32_t area0 = opaque((uint32_t)__cvta_generic_to_shared(s_area));
  const uint32_t max0 = opaque((uint32_t)__cvta_generic_to_shared(s_max));
  uint32_t phase = 0;
  for (int c = 0; c < M; c += kCrChunk) {
    const int cnt = min(kCrChunk, M - c);
    const int cnt_up = (cnt + kCrStep - 1) / kCrStep * kCrStep;
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&s_bar, (uint32_t)cnt * 16u);
      tma_bulk_g2s(s_gt, gt_boxes + m0 + c, (uint32_t)cnt * 16u, &s_bar);
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1u;
    for (int g = tid; g < cnt_up; g += kMatchBlock) {
      if (g < cnt) {
        const float4 G = s_gt[g];
        s_area[g] = box_area(G);
        s_max[g] = gt_max[m0 + c + g];   // what other CTAs found so far (stale is fine)
        s_gth[g] = make_uint2(pack_half2_rd(G.x, G.y), pack_half2_ru(G.z, G.w));
      } else {
        s_gth[g] = make_uint2(pack_half2_rd(inf, inf), pack_half2_ru(-inf, -inf));   // padding of the unrolled screen: never hit
      }
    }
    __syncthreads();

#pragma unroll
    for (int u = 0; u < kCrU; ++u) {
      const float4 au = a[u];
      const float aa = box_area(au);
      const uint32_t alo = pack_half2_rd(au.x, au.y), ahi = pack_half2_ru(au.z, au.w);
      float bvu = bv[u];
      int biu = bi[u], tpu = tops[u];
      uint32_t qp = q0;   // the lane's next free queue slot
      // the exact per-pair arithmetic of pair_update, for the pairs the screen let through
      auto drain = [&]() {
        const uint32_t most = __reduce_max_sync(kFull, qp);
        for (uint32_t qr = q0; qr < most; qr += kSlot) {
          if (qr < qp) {
            const int g = queue_at(qr);
            const float4 G = lds_f4(gt0 + 16u * g);
            const float w = __fsub_rn(fminf(G.z, au.z), fmaxf(G.x, au.x));
            const float h = __fsub_rn(fminf(G.w, au.w), fmaxf(G.y, au.y));
            if (w > 0.f && h > 0.f) {   // (the screen is a superset test: degenerate boxes pass it and stop here)
              const float inter = __fmul_rn(w, h);
              if (inter > 0.f) {
                const float uni = __fsub_rn(__fadd_rn(lds_f32(area0 + 4u * g), aa), inter);
                const float known = lds_f32(max0 + 4u * g);
                // (a pair the filter rejects is strictly below both the anchor's best and the GT's running maximum)
                if (may_reach(inter, uni, fminf(bvu, known))) {
                  const float v = __fdiv_rn(inter, uni);
                  const int top = (v >= known) ? 1 : 0;
                  if (v > bvu) { bvu = v; biu = c + g; tpu = ((tpu & 1) << 1) | (tpu & 2) | top; }
                  else tpu |= top << 1;
                  if (v > known) reds_max_u32(max0 + 4u * g, __float_as_uint(v));   // (IoU >= 0: bits order like uints)
                }
              }
            }
          }
        }
        qp = q0;
      };
      for (int g0 = 0; g0 < cnt_up; g0 += kCrStep) {
        uint4 G2[kCrStep / 2];   // two GT per 16-byte load
#pragma unroll
        for (int i = 0; i < kCrStep / 2; ++i) G2[i] = lds_u4_screen(gth0 + 8u * g0 + 16u * i);
#pragma unroll
        for (int i = 0; i < kCrStep / 2; ++i) {
          screen_push(qp, g0 + 2 * i, G2[i].x, G2[i].y, alo, ahi, kSlot);
          screen_push(qp, g0 + 2 * i + 1, G2[i].z, G2[i].w, alo, ahi, kSlot);
        }
        if (__any_sync(kFull, qp > q0 + (kCrQ - kCrStep) * kSlot)) drain();
      }
      drain();
      bv[u] = bvu; bi[u] = biu; tops[u] = tpu;
      if (u + 1 < kCrU) {
        // exchange the per-GT maxima with the other CTAs before the next anchor (no barrier: every step is a max).
        // Fresher maxima make may_reach reject more pairs and keep the "reached the running maximum" flag rare.
        for (int g = tid; g < cnt; g += kMatchBlock) {
          const unsigned mine = s_max[g], seen = __ldcg(&gt_max[m0 + c + g]);
          if (mine > seen) atomicMax(&gt_max[m0 + c + g], mine);
          else if (seen > mine) atomicMax(&s_max[g], seen);
        }
      }
    }
    __syncthreads();
    for (int g = tid; g < cnt; g += kMatchBlock) {
      const unsigned v = s_max[g];
      if (v > 0u) atomicMax(&gt_max[m0 + c + g], v);
    }
    __syncthreads();  // s_gt / s_max are rewritten by the next chunk
  }

  // ---- epilogue: as in match_pass_a_kernel
  const bool need_deltas = !STEP && (E.out.gt_deltas != nullptr);
  const bool o_matches = !STEP && E.out.matches != nullptr, o_labels = !STEP && E.out.match_labels != nullptr;
  const bool o_picky = !STEP && E.out.picky_labels != nullptr, o_cls = STEP || E.out.gt_classes != nullptr;
  const bool o_mask = STEP || E.out.mask != nullptr, o_idx32 = STEP || E.out.matched_idx32 != nullptr;
  const bool has_ids = STEP || E.gt_class_ids != nullptr, has_bets = E.bets != nullptr;
  int fg = 0;
  float w_part = 0.f;
#pragma unroll
  for (int u = 0; u < kCrU; ++u) {
    const int64_t r = base + place[u] * kMatchBlock + tid;
    if (r >= R) continue;
    const int64_t o = (int64_t)n * R + r;
    const float val = (M > 0) ? bv[u] : 0.f;
    best_val[o] = val;
    best_idx[o] = bi[u] | ((tops[u] & 2) ? kMultiOverlap : 0);
    int id = bi[u];
    int8_t l1, l2;
    int64_t cls, msk;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (M > 0) {
      l1 = band_label_reg(E.br, val);
      l2 = E.has_picky ? band_label_reg(E.pbr, val) : (int8_t)0;
      cls = has_ids ? E.gt_class_ids[m0 + id] : 0;
      if (l1 == 0) cls = E.num_classes;   // retinanet.py:356
      if (l1 == -1) cls = -1;             // :360
      msk = (l2 == 1) ? 1 : 0;            // :417-423
      if (need_deltas) d = encode_deltas(a[u], gt_boxes[m0 + id], E.wx, E.wy, E.ww, E.wh);
    } else {                              // matcher.py:70-80, retinanet.py:362-363, :425
      l1 = E.lab0;
      l2 = E.plab0;
      cls = E.num_classes;
      msk = E.num_classes;
      id = 0;
    }
    if (o_matches) E.out.matches[o] = id;
    if (o_labels) E.out.match_labels[o] = l1;
    if (o_picky) E.out.picky_labels[o] = l2;
    if (o_cls) E.out.gt_classes[o] = cls;
    if (o_mask) E.out.mask[o] = msk;
    if (need_deltas) E.out.gt_deltas[o] = d;
    if (o_idx32) E.out.matched_idx32[o] = id;
    fg += (cls >= 0 && cls != E.num_classes) ? 1 : 0;
    if (has_bets) w_part += __fadd_rn(__fmul_rn(E.bets[o], (float)msk), E.temperature);  // gambler_heads.py:569,304
    else if (lv.num_levels > 0 && msk != 0) w_part += __fmul_rn(bet_at(lv, n, r), (float)msk);  // + R*T at the fold
  }
  if (E.part_cnt != nullptr) {   // one partial per warp in a fixed slot (see match_pass_a_kernel)
    const int cw = __reduce_add_sync(kFull, fg);
    const float sw = warp_sum(w_part);
    if (lane == 0) {
      const int64_t slot = ((int64_t)n * gridDim.x + bx) * kWarpsPerBlock + (tid >> 5);
      E.part_cnt[slot] = cw;
      E.part_s[slot] = sw;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused path, pass B: the low-quality rule (matcher.py:99-132) as a patch pass, and the fold of the pre-pass sums.
// An anchor can equal some GT's maximum only if its own best IoU reaches the smallest per-GT maximum of its image
// (IoU(g,a) <= best(a)), so all but a handful of anchors are dismissed after one 4-byte load.
// ------------------------------------------------------------------------------------------
// Pass B for one run of 4 consecutive anchors per thread in a CROWDED image (more than kSmallM ground truth),
// warp-level and barrier-free.  An anchor is a candidate when its best IoU reaches the smallest per-GT maximum of the
// image.  Its own argmax GT promotes it iff best(a) == max(argmax) -- two loads, no arithmetic.  Any other GT g needs
// IoU(g, a) == max(g), which pass A rules out for all but the anchors it flagged (kMultiOverlap); for those few the
// roles flip: the candidate is broadcast and the LANES hold the ground truth (32 at a time), each testing the GT
// whose maximum the anchor's best IoU reaches (IoU(g, a) <= best(a)).  Nothing is staged or sorted.  Kept out of line
// so that its registers do not weigh on the few-GT path of the kernel.
// a_run0 / idx_run0: the image's anchors / pass A's argmax words, offset so that thread t's run starts at [4 * t];
// gt_boxes / gt_max: this image's.  Returns the 4-bit mask of promoted anchors; live4: which of the 4 exist.
__device__ __noinline__ unsigned pass_b_crowded_run(float4 val4, unsigned live4, const float4* a_run0,
                                                    const int32_t* idx_run0, const float4* gt_boxes,
                                                    const unsigned* gt_max, int M) {
  const int tid = threadIdx.x, lane = tid & 31;
  const float inf = __int_as_float(0x7f800000);
  float mn = inf;
  for (int g = lane; g < M; g += 32) mn = fminf(mn, __uint_as_float(gt_max[g]));
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) mn = fminf(mn, __shfl_xor_sync(kFull, mn, sft));
  const float val[4] = {val4.x, val4.y, val4.z, val4.w};
  unsigned lq4 = 0u;
  if (mn == 0.f) {   // a GT that overlaps no anchor promotes every anchor (matcher.py:114-116)
#pragma unroll
    for (int q = 0; q < 4; ++q) lq4 |= (((live4 >> q) & 1u) && val[q] >= 0.f) ? (1u << q) : 0u;
    return lq4;
  }
  bool hard[4];
  int bix[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool cand = ((live4 >> q) & 1u) && val[q] >= mn;
    bix[q] = cand ? idx_run0[tid * 4 + q] : 0;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool cand = ((live4 >> q) & 1u) && val[q] >= mn;
    const float gmb = cand ? __uint_as_float(gt_max[bix[q] & (kMultiOverlap - 1)]) : inf;
    const bool own = cand && val[q] == gmb;
    lq4 |= own ? (1u << q) : 0u;
    hard[q] = cand && !own && (bix[q] & kMultiOverlap) != 0;
  }
  int nhard = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) nhard += __popc(__ballot_sync(kFull, hard[q]));
  if (nhard == 0) return lq4;
  if (nhard * 12 > M) {   // (a walk costs about M / 12 of the scan below)
    // Many flagged candidates in one warp (a crowded image whose pass A CTAs all started without per-GT maxima to
    // compare against, e.g. the one crowded image of a detection batch): every lane keeps its own anchors and the
    // warp walks the ground truth once -- one broadcast load of a maximum per GT, and the box only for a GT that some
    // lane's candidate can reach.
    float4 aq[4];
    float vmax = -1.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      aq[q] = hard[q] ? a_run0[tid * 4 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (hard[q]) vmax = fmaxf(vmax, val[q]);
    }
    for (int g = 0; g < M; ++g) {
      const float gm = __uint_as_float(gt_max[g]);
      if (!__any_sync(kFull, gm <= vmax)) continue;
      const float4 G = gt_boxes[g];
      const float ga = box_area(G);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (hard[q] && gm <= val[q] && iou_exact(G, ga, aq[q], box_area(aq[q])) == gm) lq4 |= 1u << q;
    }
    return lq4;
  }
  // lane l keeps the maxima of GT l, l + 32, ... (the first kLaneGt * 32 of the image) in registers for all the
  // candidates of the warp: a candidate then costs kLaneGt compares, and loads only for the GT it can reach
  constexpr int kLaneGt = 8;
  float gmr[kLaneGt];
#pragma unroll
  for (int k = 0; k < kLaneGt; ++k) {
    const int g = lane + 32 * k;
    gmr[k] = (g < M) ? __uint_as_float(gt_max[g]) : inf;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    for (unsigned rem = __ballot_sync(kFull, hard[q]); rem != 0u; rem &= rem - 1u) {
      const int src = __ffs(rem) - 1;
      const float cval = __shfl_sync(kFull, val[q], src);
      const float4 ac = a_run0[((tid & ~31) + src) * 4 + q];   // lane src's anchor: one address for the whole warp
      const float aac = box_area(ac);
      bool hit = false;
#pragma unroll
      for (int k = 0; k < kLaneGt; ++k) {
        if (gmr[k] <= cval) {        // (padding entries are +inf)
          const float4 G = gt_boxes[lane + 32 * k];
          hit = hit || (iou_exact(G, box_area(G), ac, aac) == gmr[k]);
        }
      }
      for (int g = lane + 32 * kLaneGt; g < M; g += 32) {
        const float gm = __uint_as_float(gt_max[g]);
        if (gm <= cval) {
          const float4 G = gt_boxes[g];
          hit = hit || (iou_exact(G, box_area(G), ac, aac) == gm);
        }
      }
      if (__any_sync(kFull, hit) && lane == src) lq4 |= 1u << q;
    }
  }
  return lq4;
}

constexpr int kPassBU = 4;  // anchors per thread in pass B (8 was measured slower: 17 vs 13 us on config 2)

template <int U>
__global__ void __launch_bounds__(kMatchBlock, 4) match_pass_b_kernel(
    const float4* __restrict__ anchors, int64_t R, int64_t anchor_stride4,
    const float4* __restrict__ gt_boxes, const int64_t* __restrict__ gt_class_ids,
    const int32_t* __restrict__ gt_offsets, int N, int allow_lq, int has_picky, const float* best_val,
    const int32_t* best_idx, const unsigned* gt_max, MatchOut out,
    const float* __restrict__ bets, float temperature, int* __restrict__ part_cnt, float* __restrict__ part_s,
    const BandsReg br, const BandsReg pbr, const BetLevels lv) {
  static_assert(U % 4 == 0, "runs of 4 consecutive anchors per thread");
  grid_launch_dependents();

  const int n = blockIdx.y;
  const int m0 = gt_offsets[n];
  const int M = gt_offsets[n + 1] - m0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * (kMatchBlock * U);
  const int64_t img = (int64_t)n * R;
  const float4* a_img = anchors + (int64_t)n * anchor_stride4;
  // few GT: lane g holds GT g's box and class.  These are inputs of the step, not results of pass A, so under
  // programmatic dependent launch they are fetched while pass A still drains.
  float4 myG = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t my_cls = 0;
  if (allow_lq && M <= kSmallM && lane < M) {
    myG = gt_boxes[m0 + lane];
    my_cls = gt_class_ids ? gt_class_ids[m0 + lane] : 0;
  }
  grid_dependency_sync();     // (programmatic dependent launch behind pass A in fsg_dense_step)
  best_val = produced_by_dependency(best_val);   // pass A's results: no load of these may move above the wait
  best_idx = produced_by_dependency(best_idx);
  gt_max = produced_by_dependency(gt_max);

  // thread t owns two runs of 4 consecutive anchors: base + j*1024 + 4t .. +3 (16-byte loads of the best IoUs)
  auto r_of = [&](int u) -> int64_t { return base + (int64_t)(u >> 2) * (kMatchBlock * 4) + tid * 4 + (u & 3); };
  float val[U];
  bool live[U], lq[U];
#pragma unroll
  for (int j = 0; j < U / 4; ++j) {
    const int64_t r0 = r_of(4 * j);
    if (r0 + 4 <= R && ((img + r0) & 3) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(best_val + img + r0);
      val[4 * j] = v.x; val[4 * j + 1] = v.y; val[4 * j + 2] = v.z; val[4 * j + 3] = v.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) live[4 * j + q] = true;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        live[4 * j + q] = r0 + q < R;
        val[4 * j + q] = live[4 * j + q] ? best_val[img + r0 + q] : -1.f;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) lq[u] = false;

  int dfg = 0;       // what the promotions do to num_foreground and to the bet normaliser
  float dw = 0.f;
  if (allow_lq && M > 0 && M <= kSmallM) {
    // ---- few GT: lane g holds GT g's maximum, box and class (one parallel load, then shuffles), warp-level control
    //      flow, and every load issued as early as its address is known: the warps that do find candidates are the
    //      critical path of the whole kernel (a CTA's slot is held until its slowest warp retires).
    float my_gm = __int_as_float(0x7f800000);
    if (lane < M) my_gm = __uint_as_float(gt_max[m0 + lane]);
    float mn = my_gm;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mn = fminf(mn, __shfl_xor_sync(kFull, mn, sft));
    bool any_cand = false;
    bool cand[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cand[u] = live[u] && val[u] >= mn;
      any_cand |= cand[u];
    }
    if (__any_sync(kFull, any_cand)) {
      // the candidates' argmax (+ overlap flag), box and bet: independent loads, all in flight together
      int bix[U];
      float4 a[U];
      float betv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r_of(u);
        bix[u] = cand[u] ? best_idx[img + r] : 0;
        a[u] = cand[u] ? a_img[r] : make_float4(0.f, 0.f, 0.f, 0.f);
        betv[u] = (cand[u] && bets != nullptr) ? bets[img + r] : 0.f;
      }
      if (mn == 0.f) {
        // a GT that overlaps no anchor: every anchor has IoU 0 == its maximum 0 and is promoted (matcher.py:114-116)
#pragma unroll
        for (int u = 0; u < U; ++u) lq[u] = cand[u];
      } else {
        // IoU(g, a) <= best(a), so GT g can promote a only if max(g) <= best(a).  For g = argmax(a) that is the
        // test best(a) == max(g), no recomputation; any other g needs IoU(g, a) == max(g), which pass A rules out
        // unless it flagged the anchor (pair_update) -- only those (a fraction of the candidates) are looked at again.
        bool hard[U];
        bool any_hard = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float gmb = __shfl_sync(kFull, my_gm, bix[u] & 31);
          lq[u] = cand[u] && val[u] == gmb;
          hard[u] = cand[u] && !lq[u] && (bix[u] & kMultiOverlap) != 0;
          any_hard |= hard[u];
        }
        if (__any_sync(kFull, any_hard)) {
          const float my_ga = box_area(myG);
          for (int g = 0; g < M; ++g) {
            const float gm = __shfl_sync(kFull, my_gm, g);
            float4 G;
            G.x = __shfl_sync(kFull, myG.x, g); G.y = __shfl_sync(kFull, myG.y, g);
            G.z = __shfl_sync(kFull, myG.z, g); G.w = __shfl_sync(kFull, myG.w, g);
            const float ga = __shfl_sync(kFull, my_ga, g);
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (hard[u] && gm <= val[u] && iou_exact(G, ga, a[u], box_area(a[u])) == gm) lq[u] = true;  // :114-116
          }
        }
      }
      // patch (label 1 whatever the band said; `matches` is never changed, matcher.py:131-132)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t cls = __shfl_sync(kFull, my_cls, bix[u] & 31);
        if (!lq[u]) continue;   // (lq implies live)
        lq[u] = false;          // done here, not by the generic loop below
        const int64_t r = r_of(u);
        const int64_t o = img + r;
        if (band_label_reg(br, val[u]) != 1) {
          if (out.match_labels) out.match_labels[o] = 1;
          if (out.gt_classes) out.gt_classes[o] = cls;
          dfg += 1;             // was background or ignored, is foreground now
        }
        if (has_picky && band_label_reg(pbr, val[u]) != 1) {
          if (out.picky_labels) out.picky_labels[o] = 1;
          if (out.mask) out.mask[o] = 1;
          if (bets) dw += __fadd_rn(betv[u], temperature) - temperature;   // (bet*1 + T) - (bet*0 + T)
          else if (lv.num_levels > 0) dw += bet_at(lv, n, r);
        }
      }
    }
  } else if (allow_lq && M > 0) {
    // ---- crowded image (see pass_b_crowded_run)
#pragma unroll
    for (int j = 0; j < U / 4; ++j) {
      unsigned live4 = 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) live4 |= live[4 * j + q] ? (1u << q) : 0u;
      const unsigned lq4 = pass_b_crowded_run(make_float4(val[4 * j], val[4 * j + 1], val[4 * j + 2], val[4 * j + 3]),
                                              live4, a_img + r_of(4 * j) - tid * 4,
                                              best_idx + img + r_of(4 * j) - tid * 4, gt_boxes + m0, gt_max + m0, M);
#pragma unroll
      for (int q = 0; q < 4; ++q) lq[4 * j + q] = (lq4 >> q) & 1u;
    }
  }


  // ---- patch the anchors the low-quality rule promotes (label 1 whatever their band said; `matches` is never
  //      changed, matcher.py:131-132) and note what that does to num_foreground and to the bet normaliser
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!lq[u]) continue;   // (lq implies live)
    const int64_t r = r_of(u);
    const int64_t o = img + r;
    if (band_label_reg(br, val[u]) != 1) {
      if (out.match_labels) out.match_labels[o] = 1;
      if (out.gt_classes) out.gt_classes[o] = gt_class_ids ? gt_class_ids[m0 + (best_idx[o] & (kMultiOverlap - 1))] : 0;
      dfg += 1;             // was background or ignored, is foreground now
    }
    if (has_picky && band_label_reg(pbr, val[u]) != 1) {
      if (out.picky_labels) out.picky_labels[o] = 1;
      if (out.mask) out.mask[o] = 1;
      if (bets) dw += __fadd_rn(bets[o], temperature) - temperature;   // (bet*1 + T) - (bet*0 + T)
      else if (lv.num_levels > 0) dw += bet_at(lv, n, r);
    }
  }

  if (part_cnt == nullptr) return;
  // corrections of this WARP in its own slot, written only where there is something to add (the slots are zeroed by
  // the memset in front of pass A): almost every warp retires here without a barrier; folded by match_fold_kernel
  const int cw = __reduce_add_sync(kFull, dfg);
  const float sw = warp_sum(dw);
  if (lane == 0 && (cw != 0 || sw != 0.f)) {
    const int64_t slot = ((int64_t)n * gridDim.x + blockIdx.x) * kWarpsPerBlock + wid;
    part_cnt[slot] = cw;
    part_s[slot] = sw;
  }
}

// ------------------------------------------------------------------------------------------
// Fold of the loss pre-pass sums: num_foreground and S[n] = sum_r (bet*mask + T).  One CTA per image adds pass A's
// per-CTA partials and pass B's per-warp corrections in a fixed order (run-to-run deterministic); the last CTA to
// finish adds the images up in image order and runs the peer exchange.  A kernel of its own because 1000+ CTAs each
// ending in fence + atomic + barrier cost pass B more than all of its real work.
// ------------------------------------------------------------------------------------------
struct FoldArgs {
  const int32_t* gt_offsets;   // with gt_max: self-cleaning call, the fold zeroes what pass A / pass B dirtied
  unsigned* gt_max;
  int* clean_cnt_b;            // pass B's slots, writable (NULL: leave them)
  float* clean_s_b;
  int N;
  int64_t R;
  int nb_a;
  const int* part_cnt_a;       // pass A (N, nb_a)
  const float* part_s_a;
  const int* part_cnt_b;       // pass B (N, slots_b) or NULL: not launched
  const float* part_s_b;
  int slots_b, levels_mode;
  float temperature;
  double* img_cnt;
  unsigned* done_counter;
  double* stats;
  int peer_wait;
};

__global__ void __launch_bounds__(kMatchBlock) match_fold_kernel(const FoldArgs F, const fsg_peer_ctx peer) {
  grid_launch_dependents();
  grid_dependency_sync();
  __shared__ double s_tc[kWarpsPerBlock], s_ts[kWarpsPerBlock];
  __shared__ bool s_last;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t R = F.R;
  double* stats = F.stats;
  double* img_cnt = F.img_cnt;
  unsigned* done_counter = F.done_counter;
  const int N = F.N;
  {
    double c = 0.0, sacc = 0.0;
    for (int b = tid; b < F.nb_a; b += kMatchBlock) {
      c += (double)F.part_cnt_a[(int64_t)n * F.nb_a + b];
      sacc += (double)F.part_s_a[(int64_t)n * F.nb_a + b];
    }
    if (F.part_cnt_b != nullptr) {
      for (int b = tid; b < F.slots_b; b += kMatchBlock) {
        const int pc = F.part_cnt_b[(int64_t)n * F.slots_b + b];
        const float ps = F.part_s_b[(int64_t)n * F.slots_b + b];
        c += (double)pc;
        sacc += (double)ps;
        if (F.clean_cnt_b != nullptr && (pc != 0 || ps != 0.f)) {
          F.clean_cnt_b[(int64_t)n * F.slots_b + b] = 0;
          F.clean_s_b[(int64_t)n * F.slots_b + b] = 0.f;
        }
      }
    }
    if (F.gt_max != nullptr) {
      const int m0 = F.gt_offsets[n], m1 = F.gt_offsets[n + 1];
      for (int g = m0 + tid; g < m1; g += kMatchBlock) F.gt_max[g] = 0u;
    }
    c = warp_sum_d(c);
    sacc = warp_sum_d(sacc);
    if (lane == 0) { s_tc[wid] = c; s_ts[wid] = sacc; }
  }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, sacc = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) { c += s_tc[w]; sacc += s_ts[w]; }
    if (F.levels_mode) sacc += (double)R * (double)F.temperature;   // per-level bets: S[n] = R*T + sum bet*mask
    stats[FSG_STATS_HEADER + n] = sacc;
    img_cnt[n] = c;
    __threadfence();
    s_last = (atomicAdd(&done_counter[0], 1u) == (unsigned)N - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last image: batch totals, in image order
  __shared__ double s_loc[2];
  {
    double c = 0.0, sacc = 0.0;
    for (int im = tid; im < N; im += kMatchBlock) {
      c += __ldcg(&img_cnt[im]);
      sacc += __ldcg(&stats[FSG_STATS_HEADER + im]);
    }
    c = warp_sum_d(c);
    sacc = warp_sum_d(sacc);
    __syncthreads();   // s_tc / s_ts are reused
    if (lane == 0) { s_tc[wid] = c; s_ts[wid] = sacc; }
  }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, sacc = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) { c += s_tc[w]; sacc += s_ts[w]; }
    stats[0] = c;
    stats[1] = sacc;
    s_loc[0] = c;
    s_loc[1] = sacc;
    done_counter[0] = 0u;
  }
  if (peer.world <= 1) return;

  // ---- fused exchange over peer memory: all-reduce(SUM) of [num_foreground, S_batch] across the ranks.
  //      Mailbox of a rank: 2 (epoch parity) x 8 (sender) slots of {double v0, double v1, u64 flag, u64 fast word}.
  __shared__ double s_px[8], s_py[8];
  __syncthreads();
  unsigned long long* epoch_ptr = reinterpret_cast<unsigned long long*>(peer.epoch);
  const unsigned long long ep = *epoch_ptr + 1ull;
  if (!F.peer_wait) {
    // fsg_dense_step: what the loss kernel has to wait for is one number, num_foreground (the batch normaliser S is
    // not used when the batch is sharded).  It travels together with the epoch in ONE 8-byte word per peer -- an
    // aligned 8-byte store is atomic, so there is no payload / fence / flag sequence and no NVLink round trip on this
    // kernel's way out.  The complete record {num_foreground, S_batch} for the step's statistics is posted later, by
    // one CTA of the loss kernel (stats[0..1] stay local until that kernel's last CTA stores the global sums).
    if (tid < peer.world) {
      const int slot_out = (int)(ep & 1ull) * 8 + peer.rank;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(peer.mailbox[tid]) + slot_out * 4 + 3;
      const unsigned long long word = (ep << 32) | (unsigned long long)(unsigned)(long long)s_loc[0];
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    }
    __syncthreads();
    if (tid == 0) *epoch_ptr = ep;
    return;
  }
  if (tid < peer.world) {
    const int slot_out = (int)(ep & 1ull) * 8 + peer.rank;
    double* dst = reinterpret_cast<double*>(peer.mailbox[tid]) + slot_out * 4;
    dst[0] = s_loc[0];
    dst[1] = s_loc[1];
    __threadfence_system();
    unsigned long long* fl = reinterpret_cast<unsigned long long*>(dst + 2);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(ep) : "memory");
  }
  if (tid < peer.world) {
    const int slot_in = (int)(ep & 1ull) * 8 + tid;
    const double* src = reinterpret_cast<const double*>(peer.mailbox[peer.rank]) + slot_in * 4;
    const unsigned long long* fin = reinterpret_cast<const unsigned long long*>(src + 2);
    const long long t0 = clock64();
    unsigned long long seen = 0ull;
    const long long limit = peer.timeout_cycles > 0 ? (long long)peer.timeout_cycles : 120000000000ll;   // ~60 s
    bool arrived = true;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fin) : "memory");
      if (seen == ep) break;
      if (clock64() - t0 > limit) {   // a peer never arrived: raise the flag AND poison the sums (NaN losses and
        *reinterpret_cast<int*>(peer.error) = 1;   // gradients), so the step cannot be used silently
        arrived = false;
        break;
      }
    }
    s_px[tid] = arrived ? *reinterpret_cast<const volatile double*>(src) : __longlong_as_double(0x7ff8000000000000ll);
    s_py[tid] = arrived ? *reinterpret_cast<const volatile double*>(src + 1) : __longlong_as_double(0x7ff8000000000000ll);
  }
  __syncthreads();
  if (tid == 0) {
    double c = 0.0, sacc = 0.0;
    for (int p = 0; p < peer.world; ++p) { c += s_px[p]; sacc += s_py[p]; }   // rank order: same sum everywhere
    stats[0] = c;
    stats[1] = sacc;
    *epoch_ptr = ep;
  }
}

// ------------------------------------------------------------------------------------------
// Box2BoxTransform drop-ins
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) get_deltas_kernel(const float4* __restrict__ src,
                                                         const float4* __restrict__ tgt, int64_t n, float wx,
                                                         float wy, float ww, float wh, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = encode_deltas(src[i], tgt[i], wx, wy, ww, wh);
}

// box_regression.py:81-106
__device__ __forceinline__ float4 decode_box(float4 d, float4 b, float wx, float wy, float ww, float wh,
                                             float clampv) {
  float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
  float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
  float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
  float dw = fminf(__fdiv_rn(d.z, ww), clampv), dh = fminf(__fdiv_rn(d.w, wh), clampv);
  float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float4 o;
  o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  return o;
}

__global__ void __launch_bounds__(256) apply_deltas_kernel(const float4* __restrict__ deltas,
                                                           const float4* __restrict__ boxes, int64_t n, int k,
                                                           float wx, float wy, float ww, float wh, float clampv,
                                                           float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * k) return;
  out[i] = decode_box(deltas[i], boxes[i / k], wx, wy, ww, wh, clampv);
}

static int make_bands(const float* thr, const int8_t* lab, int nthr, MatcherBands* mb) {
  if (nthr < 1 || nthr > kMaxThr || !thr || !lab) return FSG_ERR_INVALID_ARG;
  mb->n = nthr + 1;
  for (int i = 0; i <= nthr; ++i) {
    mb->lo[i] = (i == 0) ? -__builtin_huge_valf() : thr[i - 1];
    mb->hi[i] = (i == nthr) ? __builtin_huge_valf() : thr[i];
    if (lab[i] < -1 || lab[i] > 1) return FSG_ERR_INVALID_ARG;
    mb->lab[i] = lab[i];
    if (i > 0 && i < nthr && !(thr[i - 1] <= thr[i])) return FSG_ERR_INVALID_ARG;
  }
  for (int i = nthr + 1; i <= kMaxThr; ++i) { mb->lo[i] = 0; mb->hi[i] = 0; mb->lab[i] = 0; }
  return FSG_OK;
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_pairwise_iou(const float* boxes1, int64_t n1, const float* boxes2, int64_t n2, float* iou,
                                fsg_stream_t stream) {
  if (n1 < 0 || n2 < 0) return FSG_ERR_INVALID_ARG;
  if (n1 == 0 || n2 == 0) return FSG_OK;
  if (!boxes1 || !boxes2 || !iou) return FSG_ERR_INVALID_ARG;
  dim3 grid((unsigned)ceil_div(n2, 256), (unsigned)ceil_div(n1, kIouRows));
  if (ceil_div(n1, kIouRows) > 65535) return FSG_ERR_UNSUPPORTED;
  pairwise_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes1, n1, (const float4*)boxes2,
                                                              n2, iou);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" int fsg_matcher(const float* mqm, int64_t M, int64_t N, const float* h_thresholds,
                           const int8_t* h_labels, int num_thresholds, int allow_lq, int64_t* matches,
                           int8_t* match_labels, float* ws_rowmax, fsg_stream_t stream) {
  MatcherBands mb;
  int st = make_bands(h_thresholds, h_labels, num_thresholds, &mb);
  if (st) return st;
  if (M < 0 || N < 0) return FSG_ERR_INVALID_ARG;
  if (N == 0) return FSG_OK;
  if (!matches || !match_labels || (M > 0 && !mqm)) return FSG_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (M > 0 && allow_lq) {
    if (!ws_rowmax) return FSG_ERR_WORKSPACE;
    FSG_CUDA_TRY(cudaMemsetAsync(ws_rowmax, 0, sizeof(float) * M, s));
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)ceil_div(M, 64));
    if (ceil_div(M, 64) > 65535) return FSG_ERR_UNSUPPORTED;
    matrix_rowmax_kernel<<<grid, 256, 0, s>>>(mqm, M, N, (unsigned*)ws_rowmax);
    FSG_LAUNCH_CHECK();
  }
  matrix_match_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, s>>>(mqm, M, N, mb, (M > 0 && allow_lq) ? 1 : 0,
                                                                 ws_rowmax, matches, match_labels);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

namespace {
struct MatchWs {
  size_t off_counter, off_gtmax, off_pcnt, off_ps, off_val, off_idx, off_pcnt_a, off_ps_a, off_icnt, off_zero, total;
  int nb;
};
MatchWs match_ws_layout(int N, int64_t R, int64_t sum_M) {
  MatchWs w;
  w.nb = (int)ceil_div(R > 0 ? R : 1, kMatchBlock * kPassBU);
  // per-CTA partial-sum slots of pass B and of pass A (at most: 2 anchors per thread)
  const size_t slots = (size_t)N * w.nb * kWarpsPerBlock;
  const size_t slots_a = (size_t)N * (size_t)ceil_div(R > 0 ? R : 1, kMatchBlock * 2) * kWarpsPerBlock;
  size_t o = 0;
  w.off_val = o;     o += align_up(sizeof(float) * (size_t)N * (size_t)R, 16);
  w.off_idx = o;     o += align_up(sizeof(int32_t) * (size_t)N * (size_t)R, 16);
  w.off_pcnt_a = o;  o += align_up(sizeof(int) * slots_a, 16);
  w.off_ps_a = o;    o += align_up(sizeof(float) * slots_a, 16);
  w.off_icnt = o;    o += align_up(sizeof(double) * (size_t)N, 16);
  // [off_zero, total) is zeroed by one memset before pass A.  It is the TAIL of the workspace so that the fused step
  // can extend that memset over the counter of the loss kernel, which follows in its own workspace.
  // (The per-GT maxima come last: every other offset is independent of the number of GT, so a workspace that a step
  //  left clean -- kMatchSelfClean -- is clean for the next step's GT count as well.)
  w.off_zero = o;
  w.off_counter = o; o += align_up(sizeof(unsigned) * (size_t)(N + 1), 16);
  w.off_pcnt = o;    o += align_up(sizeof(int) * slots, 16);       // pass B's slots: written only where non-zero
  w.off_ps = o;      o += align_up(sizeof(float) * slots, 16);
  w.off_gtmax = o;   o += align_up(sizeof(unsigned) * (size_t)(sum_M > 0 ? sum_M : 1), 16);
  w.total = o;
  return w;
}
}  // namespace

extern "C" size_t fsg_match_workspace_bytes(int N, int64_t R, int64_t sum_M) {
  if (N <= 0 || R < 0 || sum_M < 0) return 0;
  return match_ws_layout(N, R, sum_M).total;
}

namespace fsg {
int match_enqueue(const float* anchors, int64_t R, int64_t anchor_image_stride,
                  const float* gt_boxes, const int64_t* gt_class_ids, const int32_t* gt_offsets,
                  int N, int64_t sum_M, int num_classes, const float* h_thresholds,
                  const int8_t* h_labels, int num_thresholds, int allow_lq,
                  const float* h_picky_thresholds, const int8_t* h_picky_labels,
                  int num_picky_thresholds,
                  const float* h_box_weights, int64_t* matches, int8_t* match_labels,
                  int8_t* picky_labels, int64_t* gt_classes_out, int64_t* mask_out,
                  float* gt_deltas, int32_t* matched_idx32, const float* bets,
                  const fsg_bet_levels* h_bet_levels, float temperature,
                  double* stats, const fsg_peer_ctx* h_peer, void* workspace,
                  size_t workspace_bytes, int phases, int flags, size_t zero_tail_bytes, fsg_stream_t stream) {
  if (N <= 0 || R < 0 || sum_M < 0 || !gt_offsets) return FSG_ERR_INVALID_ARG;
  if (phases < 1 || phases > 3) return FSG_ERR_INVALID_ARG;
  BetLevels lv = {};
  if (h_bet_levels) {
    if (bets || !stats || h_bet_levels->num_levels <= 0 || h_bet_levels->num_levels > FSG_MAX_LEVELS ||
        h_bet_levels->A <= 0)
      return FSG_ERR_INVALID_ARG;
    int64_t off = 0;
    lv.A = h_bet_levels->A;
    lv.num_levels = h_bet_levels->num_levels;
    for (int l = 0; l < FSG_MAX_LEVELS; ++l) {
      lv.off[l] = off;
      if (l < lv.num_levels) {
        const int64_t hw = (int64_t)h_bet_levels->H[l] * h_bet_levels->W[l];
        if (h_bet_levels->H[l] < 0 || h_bet_levels->W[l] < 0 || hw > (1 << 30) || (hw > 0 && !h_bet_levels->bets[l]))
          return FSG_ERR_INVALID_ARG;
        lv.ptr[l] = h_bet_levels->bets[l];
        lv.HW[l] = (int)hw;
        off += hw * lv.A;
      }
    }
    lv.off[FSG_MAX_LEVELS] = off;
    if (off != R) return FSG_ERR_INVALID_ARG;
  }
  if (R == 0) return FSG_OK;
  if (!anchors || (sum_M > 0 && !gt_boxes)) return FSG_ERR_INVALID_ARG;
  if (N > 65535) return FSG_ERR_UNSUPPORTED;
  if (anchor_image_stride % 4 != 0) return FSG_ERR_INVALID_ARG;
  if (gt_classes_out && sum_M > 0 && !gt_class_ids) return FSG_ERR_INVALID_ARG;
  MatcherBands mb, pmb;
  int st = make_bands(h_thresholds, h_labels, num_thresholds, &mb);
  if (st) return st;
  if (h_picky_thresholds) {
    st = make_bands(h_picky_thresholds, h_picky_labels, num_picky_thresholds, &pmb);
    if (st) return st;
  } else {
    if (mask_out || picky_labels) return FSG_ERR_INVALID_ARG;
    pmb.n = 0;
  }
  if (bets && !stats) return FSG_ERR_INVALID_ARG;
  fsg_peer_ctx peer = {};
  peer.world = 1;
  if (h_peer) {
    if (!stats || h_peer->world < 1 || h_peer->world > 8 || h_peer->rank < 0 || h_peer->rank >= h_peer->world ||
        !h_peer->epoch || !h_peer->error)
      return FSG_ERR_INVALID_ARG;
    for (int p = 0; p < h_peer->world; ++p)
      if (!h_peer->mailbox[p]) return FSG_ERR_INVALID_ARG;
    peer = *h_peer;
  }
  const MatchWs w = match_ws_layout(N, R, sum_M);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* counter = (unsigned*)(ws + w.off_counter);
  unsigned* gtmax = (unsigned*)(ws + w.off_gtmax);
  float* bval = (float*)(ws + w.off_val);
  int32_t* bidx = (int32_t*)(ws + w.off_idx);
  // counter + gt_max + pass B's slots are contiguous: one memset node (phase 1 only: phase 2 consumes the --
  // possibly all-reduced -- per-GT maxima that phase 1 left in the workspace); zero_tail_bytes more bytes behind the
  // workspace belong to the caller (fsg_dense_step: up to and including the loss kernel's completion counter)
  // kMatchSelfClean: no memset at all -- the caller vouches that this region is zero (freshly zeroed, or left by a
  // previous self-cleaning call), and the fold kernel zeroes again what this call dirties.  A memset NODE in a CUDA
  // graph costs ~4 us of the step.
  const bool self_clean = (flags & kMatchSelfClean) != 0 && stats != nullptr && phases == 3;
  if ((phases & 1) && !self_clean)
    FSG_CUDA_TRY(cudaMemsetAsync(ws + w.off_zero, 0, w.total - w.off_zero + zero_tail_bytes, s));
  const float wx = h_box_weights ? h_box_weights[0] : 1.f, wy = h_box_weights ? h_box_weights[1] : 1.f;
  const float ww = h_box_weights ? h_box_weights[2] : 1.f, wh = h_box_weights ? h_box_weights[3] : 1.f;

  // anchors per thread in pass A: 2 for the few-GT training case (more CTAs in flight), 4 when the batch is crowded
  // (each staged GT box is reused more); the same choice sizes pass A's partial-sum slots in both phases
  const bool few_gt = sum_M <= (int64_t)N * kSmallM;
  static const bool crowded_queue = getenv("FSG_MATCH_CROWDED_BRUTE") == nullptr;   // (diagnostic switch for A/B timing)
  const int nb_a = (int)ceil_div(R, kMatchBlock * (few_gt ? 2 : 4));
  MatchOut out{matches, match_labels, picky_labels, gt_classes_out, mask_out, (float4*)gt_deltas, matched_idx32};
  if (phases & 1) {
    PassAEpi e;
    e.gt_class_ids = gt_class_ids; e.num_classes = num_classes;
    e.lab0 = mb.lab[0]; e.plab0 = pmb.n ? pmb.lab[0] : (int8_t)0; e.has_picky = pmb.n != 0;
    e.br = make_bands_reg(mb); e.pbr = make_bands_reg(pmb);
    e.wx = wx; e.wy = wy; e.ww = ww; e.wh = wh;
    e.out = out; e.bets = bets; e.temperature = temperature;
    e.part_cnt = stats ? (int*)(ws + w.off_pcnt_a) : nullptr;
    e.part_s = (float*)(ws + w.off_ps_a);
    dim3 grid_a((unsigned)nb_a, (unsigned)N);
    const bool step_set = !matches && !match_labels && !picky_labels && !gt_deltas && gt_classes_out && mask_out &&
                          matched_idx32 && gt_class_ids && pmb.n != 0;
    auto go_a = [&](auto kern) {
      kern<<<grid_a, kMatchBlock, 0, s>>>((const float4*)anchors, R, anchor_image_stride / 4, (const float4*)gt_boxes,
                                          gt_offsets, bval, bidx, gtmax, e, lv);
    };
    if (few_gt) { if (step_set) go_a(match_pass_a_kernel<2, true>); else go_a(match_pass_a_kernel<2, false>); }
    else if (crowded_queue) {
      static_assert(kCrU == 4, "pass A's partial-sum slots are sized for 4 anchors per thread in crowded batches");
      if (step_set) go_a(match_pass_a_crowded_kernel<true>); else go_a(match_pass_a_crowded_kernel<false>);
    } else      { if (step_set) go_a(match_pass_a_kernel<4, true>); else go_a(match_pass_a_kernel<4, false>); }
    FSG_LAUNCH_CHECK();
  }
  if (!(phases & 2)) return FSG_OK;
  const bool pdl = (flags & kMatchPdl) != 0 && phases == 3;
  const bool patch = allow_lq != 0;
  const int nb_b = w.nb;
  if (patch) {
    dim3 grid_b((unsigned)nb_b, (unsigned)N);
    auto go = [&](auto kern) {
      return launch_pdl(kern, grid_b, dim3(kMatchBlock), 0, s, pdl,
                        (const float4*)anchors, R, anchor_image_stride / 4, (const float4*)gt_boxes, gt_class_ids,
                        gt_offsets, N, 1, pmb.n != 0 ? 1 : 0, (const float*)bval, (const int32_t*)bidx,
                        (const unsigned*)gtmax, out, bets, temperature,
                        stats ? (int*)(ws + w.off_pcnt) : (int*)nullptr, (float*)(ws + w.off_ps),
                        make_bands_reg(mb), make_bands_reg(pmb), lv);
    };
    go(match_pass_b_kernel<kPassBU>);
    FSG_LAUNCH_CHECK();
  }
  if (!stats) return FSG_OK;
  FoldArgs f = {};
  if (self_clean) {
    f.gt_offsets = gt_offsets; f.gt_max = gtmax;
    if (patch) { f.clean_cnt_b = (int*)(ws + w.off_pcnt); f.clean_s_b = (float*)(ws + w.off_ps); }
  }
  f.N = N; f.R = R; f.nb_a = nb_a * kWarpsPerBlock;   // (pass A leaves one slot per warp)
  f.part_cnt_a = (const int*)(ws + w.off_pcnt_a); f.part_s_a = (const float*)(ws + w.off_ps_a);
  f.part_cnt_b = patch ? (const int*)(ws + w.off_pcnt) : nullptr; f.part_s_b = (const float*)(ws + w.off_ps);
  f.slots_b = nb_b * kWarpsPerBlock; f.levels_mode = lv.num_levels > 0 ? 1 : 0; f.temperature = temperature;
  f.img_cnt = (double*)(ws + w.off_icnt); f.done_counter = counter; f.stats = stats;
  f.peer_wait = (flags & kMatchPeerPolled) ? 0 : 1;
  launch_pdl(match_fold_kernel, dim3((unsigned)N), dim3(kMatchBlock), 0, s, pdl, f, peer);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
}  // namespace fsg

extern "C" int fsg_match_anchors_ex(const float* anchors, int64_t R, int64_t anchor_image_stride,
                                 const float* gt_boxes, const int64_t* gt_class_ids, const int32_t* gt_offsets,
                                 int N, int64_t sum_M, int num_classes, const float* h_thresholds,
                                 const int8_t* h_labels, int num_thresholds, int allow_lq,
                                 const float* h_picky_thresholds, const int8_t* h_picky_labels,
                                 int num_picky_thresholds,
                                 const float* h_box_weights, int64_t* matches, int8_t* match_labels,
                                 int8_t* picky_labels, int64_t* gt_classes_out, int64_t* mask_out,
                                 float* gt_deltas, int32_t* matched_idx32, const float* bets,
                                 const fsg_bet_levels* h_bet_levels, float temperature,
                                 double* stats, const fsg_peer_ctx* h_peer, void* workspace,
                                 size_t workspace_bytes, int phases, fsg_stream_t stream) {
  return match_enqueue(anchors, R, anchor_image_stride, gt_boxes, gt_class_ids, gt_offsets, N, sum_M, num_classes,
                       h_thresholds, h_labels, num_thresholds, allow_lq, h_picky_thresholds, h_picky_labels,
                       num_picky_thresholds, h_box_weights, matches, match_labels, picky_labels, gt_classes_out,
                       mask_out, gt_deltas, matched_idx32, bets, h_bet_levels, temperature, stats, h_peer, workspace,
                       workspace_bytes, phases, 0, 0, stream);
}

extern "C" int fsg_box2box_get_deltas(const float* src_boxes, const float* target_boxes, int64_t n,
                                      const float* h_weights, float* deltas, fsg_stream_t stream) {
  if (n < 0 || !h_weights) return FSG_ERR_INVALID_ARG;
  if (n == 0) return FSG_OK;
  if (!src_boxes || !target_boxes || !deltas) return FSG_ERR_INVALID_ARG;
  get_deltas_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)src_boxes, (const float4*)target_boxes, n, h_weights[0], h_weights[1], h_weights[2],
      h_weights[3], (float4*)deltas);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" int fsg_box2box_apply_deltas(const float* deltas, const float* boxes, int64_t n, int k,
                                        const float* h_weights, float scale_clamp, float* out,
                                        fsg_stream_t stream) {
  if (n < 0 || k < 1 || !h_weights) return FSG_ERR_INVALID_ARG;
  if (n == 0) return FSG_OK;
  if (!deltas || !boxes || !out) return FSG_ERR_INVALID_ARG;
  apply_deltas_kernel<<<(unsigned)ceil_div(n * k, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)deltas, (const float4*)boxes, n, k, h_weights[0], h_weights[1], h_weights[2], h_weights[3],
      scale_clamp, (float4*)out);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" int fsg_match_anchors(const float* anchors, int64_t R, int64_t anchor_image_stride,
                                 const float* gt_boxes, const int64_t* gt_class_ids, const int32_t* gt_offsets,
                                 int N, int64_t sum_M, int num_classes, const float* h_thresholds,
                                 const int8_t* h_labels, int num_thresholds, int allow_lq,
                                 const float* h_picky_thresholds, const int8_t* h_picky_labels,
                                 int num_picky_thresholds, const float* h_box_weights, int64_t* matches,
                                 int8_t* match_labels, int8_t* picky_labels, int64_t* gt_classes_out,
                                 int64_t* mask_out, float* gt_deltas, int32_t* matched_idx32, const float* bets,
                                 const fsg_bet_levels* h_bet_levels, float temperature, double* stats,
                                 const fsg_peer_ctx* h_peer, void* workspace, size_t workspace_bytes,
                                 fsg_stream_t stream) {
  return fsg_match_anchors_ex(anchors, R, anchor_image_stride, gt_boxes, gt_class_ids, gt_offsets, N, sum_M,
                              num_classes, h_thresholds, h_labels, num_thresholds, allow_lq, h_picky_thresholds,
                              h_picky_labels, num_picky_thresholds, h_box_weights, matches, match_labels,
                              picky_labels, gt_classes_out, mask_out, gt_deltas, matched_idx32, bets, h_bet_levels,
                              temperature, stats, h_peer, workspace, workspace_bytes, 3, stream);
}

extern "C" size_t fsg_match_gt_max_offset(int N, int64_t R, int64_t sum_M) {
  if (N <= 0 || R < 0 || sum_M < 0) return 0;
  return match_ws_layout(N, R, sum_M).off_gtmax;
}
