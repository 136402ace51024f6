// K3: score scan + per-level top-k + threshold + box decode + per-class NMS + top detections.
//
// Reference semantics (paths relative to the reference tree):
//   detectron2/modeling/meta_arch/retinanet.py:460-520  RetinaNet.inference_single_image
//   detectron2/modeling/box_regression.py:69-107        Box2BoxTransform.apply_deltas
//   detectron2/layers/nms.py:6,9-26                     nms / batched_nms (torchvision greedy NMS,
//                                                        per-class un-offset form)
//
// Where the reference sorts every level's H*W*A*K scores to keep 1000, the select stage here is
//   1. detect_bar_kernel       a systematic sample of every (image, level) slab (32-byte pieces, <= 32768 values per
//                              CTA, 16-bit order-preserving keys in shared memory, radix select) yields a logit bar
//                              that about 3k slab elements are expected to reach (exact k-th value for small slabs);
//   2. detect_scan_kernel      streams the logits once (coalesced 16-byte loads, no shared-memory traffic, no
//                              barrier): `logit >= bar` hits go through a per-warp staging area into the slab's
//                              candidate list in global memory;
//   3. detect_finalize_kernel  one CTA per slab: exact sigmoid, `score > threshold`, exact top-k by (score, index)
//                              radix select, sort, box decode.  It PROVES the bar was safe (at least k candidates and
//                              the k-th best score strictly above anything a non-candidate can score) -- otherwise
//   4. detect_select_kernel    the exact streaming top-k (running candidate buffer per CTA, radix-select pruning
//                              raises the bar as it goes; handles any tie pattern) redoes that slab.  On ordinary
//                              inputs every CTA of this launch exits at once.
// Both layouts are read in place: the reference's flattened (N, sum HWA, K) rows or the head's own per-level
// (N, A*K, H, W) planes (a slab is one contiguous piece of memory either way; only a candidate's index is mapped
// to the reference's (h, w, a, k) order, so ties rank exactly as in the reference's stable sort).
// One CTA per image (split by class) then runs the per-class greedy NMS entirely in shared memory (bitonic sort
// on composite keys, warp-per-class suppression) and emits the final detections -- no host round trip, no n^2
// mask in global memory.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "nms_kernel.cuh"
#include "sort_utils.cuh"

namespace fsg {

constexpr int kSelThreads = 256;
#ifndef SEL_CAP
#define SEL_CAP 4096
#endif
#ifndef SEL_CTAS
#define SEL_CTAS 4
#endif
constexpr int kSelCap = SEL_CAP;                       // candidate buffer entries (u64 keys): 32 KB -> 4 CTAs/SM
constexpr int kSelIter = kSelThreads * 8;           // elements consumed per block iteration
constexpr int kStagePerWarp = 32;                   // raw (logit, index) candidates a warp collects before it
                                                    // evaluates their sigmoids as one dense batch
// worst case a warp adds 8*32 new + one full stage per block iteration: prune while that still fits
constexpr int kSelTrigger = kSelCap - (kSelThreads / 32) * (256 + kStagePerWarp);
struct DetectLevels {
  int64_t off[kMaxLevels + 1];  // anchor offsets of the levels
  int nparts[kMaxLevels];       // CTAs per (image, level) slab
  int part_base[kMaxLevels];    // first blockIdx.x of the level
  int k[kMaxLevels];            // min(topk, HWA_l)
  int64_t part_len[kMaxLevels]; // elements per part (multiple of kSelIter)
  int scan_base[kMaxLevels + 1]; // detect_scan_kernel: first blockIdx.x of the level
  int bar_parts[kMaxLevels];    // detect_bar_kernel: sampling CTAs per slab
  int bar_rank[kMaxLevels];     // rank of the bar among the pooled sample (0: exact mode, rank = k)
  int bar_pieces[kMaxLevels];   // sampled 32-byte pieces per sampling CTA
  int num_levels;
  int total_parts;
  int max_parts;                // slot stride (parts) per level in the scratch arrays
};

// Where the scores and box deltas of a slab live.  Flat: rows (h*W+w)*A+a of (N, sum HWA, K) / (N, sum HWA, 4);
// native: the head's per-level (N, A*K, H, W) / (N, A*4, H, W) conv outputs (retinanet.py:40-43).
struct DetectSrc {
  const float* cls[kMaxLevels];
  const float* reg[kMaxLevels];
  int64_t cls_stride[kMaxLevels];   // floats between consecutive images
  int64_t reg_stride[kMaxLevels];
  int HW[kMaxLevels];
  int A, K, native;
};

// position inside a slab -> index in the reference's flattened (HWA, K) order (retinanet.py:24-33, :498-499)
__device__ __forceinline__ uint32_t ref_index(const DetectSrc& S, int l, uint32_t pos) {
  if (!S.native) return pos;
  const uint32_t HW = (uint32_t)S.HW[l], K = (uint32_t)S.K;
  const uint32_t c = pos / HW, hw = pos - c * HW;
  const uint32_t a = c / K, k = c - a * K;
  return (hw * (uint32_t)S.A + a) * K + k;
}
__device__ __forceinline__ float4 load_delta(const DetectSrc& S, int l, int n, uint32_t a_idx) {
  const float* base = S.reg[l] + (int64_t)n * S.reg_stride[l];
  if (!S.native) return reinterpret_cast<const float4*>(base)[a_idx];
  const uint32_t hw = a_idx / (uint32_t)S.A, a = a_idx - hw * (uint32_t)S.A;
  const int64_t HW = S.HW[l];
  const float* p = base + (int64_t)(a * 4) * HW + hw;
  return make_float4(p[0], p[HW], p[2 * HW], p[3 * HW]);
}

constexpr int kStatusOk = 1, kStatusFallback = 2;

__device__ __forceinline__ float sigmoid_score(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

// key: score bits in the high word, inverted slab index in the low word => descending key order is
// (score descending, index ascending), the reference's stable descending sort (retinanet.py:489).
__device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  return ((uint64_t)__float_as_uint(score) << 32) | (uint64_t)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(uint64_t k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return 0xffffffffu - (uint32_t)k; }

// ---- block-wide exact k-th largest over 64-bit keys in shared memory (11-bit radix select) ---------
// returns T such that exactly k keys are >= T (keys are distinct).  Requires count >= k >= 1.
// Histogram increments are aggregated with match.any first: score keys share their exponent bits, so the
// top digits put almost every key in one bin and un-aggregated shared atomics would serialise 32-fold.
constexpr int kDigitBits = 10;
constexpr int kBins = 1 << kDigitBits;
// slack > 0: an in-stream prune may stop after two digits once the selected bin holds at most `slack` keys more than
// needed -- the threshold then keeps a few extra keys (at most k + slack), which the next prune sorts out.
template <int NT>
__device__ uint64_t select_kth(const uint64_t* buf, int count, int k, unsigned* hist /*kBins*/, int* s_tmp /*4*/,
                               int slack = 0) {
  const int tid = threadIdx.x, lane = tid & 31;
  uint64_t prefix = 0;
  int need = k;
  const int rounds = (count + NT - 1) / NT;
  for (int top = 64; top > 0; top -= kDigitBits) {
    const int width = top >= kDigitBits ? kDigitBits : top;   // 10,10,10,10,10,10,4
    const int shift = top - width;
    for (int b = tid; b < kBins; b += NT) hist[b] = 0u;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      const int i = rd * NT + tid;
      unsigned bin = 0u;
      bool ok = false;
      if (i < count) {
        const uint64_t key = buf[i];
        ok = (top == 64) || ((key >> top) == (prefix >> top));
        bin = (unsigned)(key >> shift) & ((1u << width) - 1u);
      }
      warp_hist_add(hist, bin, ok);
    }
    __syncthreads();
    if (tid < 32) {
      // lane L owns the 64 bins kBins-1-64L .. kBins-64-64L (descending)
      constexpr int PER = kBins / 32;
      unsigned sum = 0;
      for (int b = 0; b < PER; ++b) sum += hist[kBins - 1 - PER * lane - b];
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned before = inc - sum;
      if (before < (unsigned)need && inc >= (unsigned)need) {
        unsigned cum = before;
        for (int b = 0; b < PER; ++b) {
          const unsigned h = hist[kBins - 1 - PER * lane - b];
          if (cum < (unsigned)need && cum + h >= (unsigned)need) {
            s_tmp[0] = kBins - 1 - PER * lane - b;  // digit
            s_tmp[1] = (int)cum;                    // keys strictly above this digit (within prefix)
            s_tmp[2] = (int)h;                      // keys in this digit's bin
          }
          cum += h;
        }
      }
    }
    __syncthreads();
    const int digit = s_tmp[0], above = s_tmp[1], inbin = s_tmp[2];
    need -= above;
    prefix |= (uint64_t)digit << shift;
    __syncthreads();
    if (inbin == need) break;  // the whole bin is taken: low bits of the threshold stay zero
    if (slack > 0 && top <= 64 - kDigitBits && inbin - need <= slack) break;   // good enough for now
  }
  return prefix;
}

// keep only keys >= T (order not preserved).  count <= kSelCap.
template <int NT>
__device__ void compact_ge(uint64_t* buf, int* s_count, uint64_t T) {
  constexpr int PER = (kSelCap + NT - 1) / NT;
  const int tid = threadIdx.x;
  const int count = *s_count;
  uint64_t mine[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * NT;
    mine[j] = (i < count) ? buf[i] : 0ull;
  }
  __syncthreads();
  if (tid == 0) *s_count = 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * NT;
    if (i < count && mine[j] >= T) buf[atomicAdd(s_count, 1)] = mine[j];
  }
  __syncthreads();
}

// block-wide minimum of buf[0..count); result valid in every thread
template <int NT>
__device__ uint64_t block_min_u64(const uint64_t* buf, int count, uint64_t* s_red /*NT/32*/) {
  uint64_t mn = ~0ull;
  for (int i = threadIdx.x; i < count; i += NT) mn = min(mn, buf[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, (uint64_t)__shfl_xor_sync(kFull, (unsigned long long)mn, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mn;
  __syncthreads();
  mn = s_red[0];
  for (int w = 1; w < NT / 32; ++w) mn = min(mn, s_red[w]);
  __syncthreads();
  return mn;
}

template <int NT>
__device__ void prune_topk(uint64_t* buf, int* s_count, int k, unsigned* hist, int* s_tmp, int slack = 0) {
  const int count = *s_count;  // caller synchronised
  if (count <= k) return;
  const uint64_t T = select_kth<NT>(buf, count, k, hist, s_tmp, slack);
  compact_ge<NT>(buf, s_count, T);
}

// descending bitonic sort of m (power of two) keys in shared memory
__device__ __forceinline__ float4 decode_box3(float4 d, float4 b, float wx, float wy, float ww, float wh,
                                              float clampv) {  // box_regression.py:81-106
  float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
  float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
  float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
  float dw = fminf(__fdiv_rn(d.z, ww), clampv), dh = fminf(__fdiv_rn(d.w, wh), clampv);
  float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  return make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                     __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));
}

struct SelectArgs {
  const float4* anchors;
  int64_t anchor_stride4;
  int64_t R;
  int K;
  int topk;       // slot stride per level in the candidate arrays
  float thr;      // SCORE_THRESH_TEST
  float xpre;     // conservative logit pre-filter for thr
  float wx, wy, ww, wh, clampv;
  uint64_t* part_keys;   // (N, L, max_parts, topk)
  int* part_count;       // (N, L, max_parts)
  unsigned* done;        // (N, L)
  float4* cand_box;      // (N, L*topk)
  float* cand_score;     // (N, L*topk)
  int64_t* cand_class;   // (N, L*topk)
  int* lvl_count;        // (N, L)
  // sampled-bar path (NULL status: the streaming top-k kernel runs on every slab)
  unsigned* bar_inv;     // (N, L) ~ordered bits of the slab's bar, merged with atomicMax (zero-initialised)
  unsigned* cand_count;  // (N, L) candidates appended so far (may exceed cand_cap: overflow)
  uint64_t* cand_list;   // (N, L, cand_cap) (logit bits << 32 | slab position)
  uint16_t* pool;        // (N, L, kBarPoolParts * kBarPoolKeys) best sample keys of each sampling CTA
  unsigned* pool_done;   // (N, L)
  int* status;           // (N, L) kStatusOk / kStatusFallback, written by detect_finalize_kernel
  int cand_cap;
};

// Evaluate the warp's staged raw candidates as dense batches: sigmoid (expf + IEEE divide), the exact
// `score > ts` test and the append to the CTA's key buffer.  Called by all 32 lanes.
__device__ __forceinline__ void flush_stage(const float2* stage, int n, float ts, uint64_t* buf, int* s_count,
                                            const DetectSrc& S, int l) {
  const int lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool p = i < n;
    float sc = 0.f;
    uint32_t id = 0u;
    if (p) {
      const float2 e = stage[i];
      sc = sigmoid_score(e.x);
      id = __float_as_uint(e.y);
      p = sc > ts;
    }
    const unsigned m = __ballot_sync(kFull, p);
    if (m != 0u) {
      int base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(s_count, __popc(m));
      base = __shfl_sync(kFull, base, leader);
      if (p) buf[base + __popc(m & ((1u << lane) - 1u))] = make_key(sc, ref_index(S, l, id));
    }
  }
}

// Collect the lanes' logits that pass the (cheap, conservative) logit pre-filter into the warp's staging
// queue; the expensive part runs later on full warps (flush_stage).  v[0..3] sit at idx0.., v[4..7] at idx1..
__device__ __forceinline__ void stage_candidates(const float* v, uint32_t idx0, uint32_t idx1, float xb, float ts,
                                                 float2* stage, int* s_scnt, uint64_t* buf, int* s_count,
                                                 const DetectSrc& S, int l) {
  const int lane = threadIdx.x & 31;
  unsigned flags = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) flags |= (v[j] > xb) ? (1u << j) : 0u;
  const int c = __popc(flags);
  const int tot = __reduce_add_sync(kFull, c);
  if (tot == 0) return;
  int cur = *s_scnt;
  if (cur + tot > kStagePerWarp) {
    flush_stage(stage, cur, ts, buf, s_count, S, l);
    __syncwarp();
    if (lane == 0) *s_scnt = 0;
    __syncwarp();
    cur = 0;
  }
  if (tot > kStagePerWarp) {
    // dense phase (pre-filter still loose): go through the stage in slices of one element position
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool p = (flags >> j) & 1u;
      const unsigned m = __ballot_sync(kFull, p);
      if (m == 0u) continue;
      if (p) stage[__popc(m & ((1u << lane) - 1u))] = make_float2(v[j], __uint_as_float((j < 4 ? idx0 : idx1 - 4) + j));
      __syncwarp();
      flush_stage(stage, __popc(m), ts, buf, s_count, S, l);
      __syncwarp();
    }
    return;
  }
  if (c == 1) {
    // the common case late in the scan: one candidate in this lane -> pick it with a select tree
    const int j = __ffs(flags) - 1;
    const float lo = (j & 2) ? ((j & 1) ? v[3] : v[2]) : ((j & 1) ? v[1] : v[0]);
    const float hi = (j & 2) ? ((j & 1) ? v[7] : v[6]) : ((j & 1) ? v[5] : v[4]);
    const int off = atomicAdd(s_scnt, 1);
    stage[off] = make_float2((j & 4) ? hi : lo, __uint_as_float(((j & 4) ? idx1 - 4 : idx0) + j));
  } else if (c > 1) {
    int off = atomicAdd(s_scnt, c);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if ((flags >> j) & 1u) stage[off++] = make_float2(v[j], __uint_as_float((j < 4 ? idx0 : idx1 - 4) + j));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kSelThreads, SEL_CTAS) detect_select_kernel(const SelectArgs A, const DetectLevels LV,
                                                                               const DetectSrc S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);  // kSelCap
  __shared__ unsigned hist[kBins];
  __shared__ int s_tmp[4];
  __shared__ int s_count;
  __shared__ int s_scnt[kSelThreads / 32];
  __shared__ float s_xb, s_ts;
  __shared__ bool s_last;
  __shared__ uint64_t s_red64[kSelThreads / 32];

  const int tid = threadIdx.x;
  grid_launch_dependents();   // (programmatic dependent launch: the NMS kernel behind may be scheduled early)
  grid_dependency_sync();
  const int n = blockIdx.y;
  int l = 0;
  while (l + 1 < LV.num_levels && (int)blockIdx.x >= LV.part_base[l + 1]) ++l;
  if (A.status != nullptr && A.status[n * LV.num_levels + l] != kStatusFallback) return;   // slab already done
  const int part = blockIdx.x - LV.part_base[l];
  const int k = LV.k[l];
  const int64_t hwa = LV.off[l + 1] - LV.off[l];
  const int64_t E = hwa * A.K;                         // elements in the slab
  const float* slab = S.cls[l] + (int64_t)n * S.cls_stride[l];
  const int64_t e0 = (int64_t)part * LV.part_len[l];
  const int64_t e1 = min(E, e0 + LV.part_len[l]);

  if (tid == 0) { s_count = 0; s_xb = A.xpre; s_ts = A.thr; }
  if (tid < kSelThreads / 32) s_scnt[tid] = 0;
  __syncthreads();

  const bool vec_ok = ((reinterpret_cast<uintptr_t>(slab) & 15) == 0);
  __shared__ float2 s_stage[kSelThreads / 32][kStagePerWarp];
  float2* stage = s_stage[tid >> 5];
  // checked loader for ragged ends / unaligned slabs
  auto load8_checked = [&](int64_t base, float* v) {
    const int64_t p0 = base + (int64_t)tid * 4;
    const int64_t p1 = p0 + kSelThreads * 4;
    if (vec_ok && p0 + 4 <= e1) {
      float4 t = ldg_stream4(slab + p0); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (p0 + j < e1) ? ldg_stream1(slab + p0 + j) : -INFINITY;
    }
    if (vec_ok && p1 + 4 <= e1) {
      float4 t = ldg_stream4(slab + p1); v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[4 + j] = (p1 + j < e1) ? ldg_stream1(slab + p1 + j) : -INFINITY;
    }
  };
  // one block iteration: pre-filter, stage, barrier, prune when the key buffer may overflow
  auto consume = [&](const float* v, int64_t base) {
    const float xb = s_xb, ts = s_ts;
    const float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
    if (__any_sync(kFull, m > xb)) {
      const uint32_t p0 = (uint32_t)(base + (int64_t)tid * 4);
      stage_candidates(v, p0, p0 + kSelThreads * 4, xb, ts, stage, &s_scnt[tid >> 5], buf, &s_count, S, l);
    }
    __syncthreads();
    if (s_count > kSelTrigger) {   // uniform: read after the barrier
      // in-stream prunes may keep up to 64 keys too many when that still leaves the buffer well under the trigger
      prune_topk<kSelThreads>(buf, &s_count, k, hist, s_tmp, (k + 64 <= kSelTrigger / 2 + 256) ? 64 : 0);
      const int now = s_count;
      if (now >= k) {
        // buffer now holds the k best so far: raise the bar to their minimum.  A candidate that is still waiting
        // in a warp's staging queue can have a LOWER index than keys already in the buffer (the queues of the
        // eight warps drain at different times), so an equal score must still be admitted (`score >= minimum`);
        // the exact (score, index) order is settled by the next prune.
        const uint64_t mn = block_min_u64<kSelThreads>(buf, now, s_red64);
        if (tid == 0) {
          const float t = nextafterf(key_score(mn), 0.f);  // `score > t`  <=>  `score >= minimum`
          s_ts = t;
          const float lg = logf(t / (1.f - t));
          s_xb = fmaxf(A.xpre, lg - 1e-4f * (1.f + fabsf(lg)));
        }
      }
      __syncthreads();
    }
  };
  // full iterations: unconditional 16-byte loads through a running pointer, two register sets in
  // ping-pong so the next block's loads are in flight while this one is consumed
  const int64_t n_full = vec_ok ? (e1 - e0) / kSelIter : 0;
  {
    const float* pa = slab + e0 + (int64_t)tid * 4;
    float va[8], vb[8];
    auto ld = [&](const float* p, float* v) {
      const float4 t0 = ldg_stream4(p), t1 = ldg_stream4(p + kSelThreads * 4);
      v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    };
    if (n_full > 0) ld(pa, va);
    for (int64_t it = 0; it < n_full; it += 2) {
      if (it + 1 < n_full) ld(pa + (it + 1) * kSelIter, vb);
      consume(va, e0 + it * kSelIter);
      if (it + 1 < n_full) {
        if (it + 2 < n_full) ld(pa + (it + 2) * kSelIter, va);
        consume(vb, e0 + (it + 1) * kSelIter);
      }
    }
  }
  for (int64_t base = e0 + n_full * kSelIter; base < e1; base += kSelIter) {   // ragged end / unaligned slab
    float v[8];
    load8_checked(base, v);
    consume(v, base);
  }
  flush_stage(stage, s_scnt[tid >> 5], s_ts, buf, &s_count, S, l);   // what is still waiting in the warp queues
  __syncthreads();
  prune_topk<kSelThreads>(buf, &s_count, k, hist, s_tmp);
  __syncthreads();

  // ---- publish this part's candidates
  const int cnt = s_count;
  const int64_t slot = (((int64_t)n * LV.num_levels + l) * LV.max_parts + part);
  uint64_t* gk = A.part_keys + slot * A.topk;
  for (int i = tid; i < cnt; i += kSelThreads) gk[i] = buf[i];
  __syncthreads();
  if (tid == 0) {
    A.part_count[slot] = cnt;
    __threadfence();
    s_last = (atomicAdd(&A.done[n * LV.num_levels + l], 1u) == (unsigned)LV.nparts[l] - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- last CTA of the slab: merge parts, exact top-k, sort, decode
  if (tid == 0) { s_count = 0; A.done[n * LV.num_levels + l] = 0u; }
  __syncthreads();
  const int nparts = LV.nparts[l];
  if (nparts > 1) {
    for (int p = 0; p < nparts; ++p) {
      const int64_t sl = (((int64_t)n * LV.num_levels + l) * LV.max_parts + p);
      const int c = __ldcg(&A.part_count[sl]);
      const uint64_t* src = A.part_keys + sl * A.topk;
      __shared__ int s_base;
      if (tid == 0) { s_base = s_count; s_count += c; }
      __syncthreads();
      for (int i = tid; i < c; i += kSelThreads) buf[s_base + i] = __ldcg(&src[i]);
      __syncthreads();
    }
    prune_topk<kSelThreads>(buf, &s_count, k, hist, s_tmp);
    __syncthreads();
  } else {
    if (tid == 0) s_count = cnt;
    __syncthreads();
  }
  const int fin = s_count;
  int m = 1;
  while (m < fin) m <<= 1;
  for (int i = fin + tid; i < m; i += kSelThreads) buf[i] = 0ull;
  __syncthreads();
  bitonic_desc<kSelThreads>(buf, m);
  const int64_t cbase = (int64_t)n * LV.num_levels * A.topk + (int64_t)l * A.topk;
  for (int t = tid; t < fin; t += kSelThreads) {
    const uint64_t key = buf[t];
    const uint32_t idx = key_index(key);
    const int64_t a = idx / (uint32_t)A.K;          // retinanet.py:498-499
    const int c = (int)(idx - (uint32_t)a * (uint32_t)A.K);
    const int64_t r = LV.off[l] + a;
    const float4 d = load_delta(S, l, n, (uint32_t)a);
    const float4 an = A.anchors[(int64_t)n * A.anchor_stride4 + r];
    A.cand_box[cbase + t] = decode_box3(d, an, A.wx, A.wy, A.ww, A.wh, A.clampv);
    A.cand_score[cbase + t] = key_score(key);
    A.cand_class[cbase + t] = c;
  }
  if (tid == 0) A.lvl_count[n * LV.num_levels + l] = fin;
}

constexpr size_t kSelSmem = (size_t)kSelCap * 8;

// ==========================================================================================================
// Sampled-bar path: detect_bar_kernel -> detect_scan_kernel -> detect_finalize_kernel
// ==========================================================================================================
constexpr int kBarThreads = 512;
constexpr int kBarKeys = 32768;            // sample values per CTA (16-bit keys: 64 KB of shared memory)
constexpr int kBarPiece = 32;              // floats per sampled piece (128 bytes: what one DRAM access brings in anyway)
constexpr int kBarPoolParts = 32;          // sampling CTAs per slab, at most
constexpr int kBarPoolKeys = 64;           // best keys each of them contributes to the slab's pool
constexpr float kBarOver = 3.f;            // aim for about kBarOver * k candidates per slab
constexpr int kBarRank = 48;               // rank wanted for the bar inside the sample (sets the sample size)
constexpr int kCandCap = 8192;             // candidate list entries per slab
constexpr int kScanThreads = 256;
constexpr int kScanIter = kScanThreads * 8;          // elements per block iteration (two float4 per thread)
constexpr int kScanIters = 16;
constexpr int kScanChunk = kScanIter * kScanIters;   // elements per CTA: 32768 (128 KB)
constexpr int kScanStage = 64;                       // per-warp staging entries
constexpr int kFinThreads = 512;
constexpr int kFinSort = 2048;                       // >= the largest supported top-k (kSelTrigger)

// order-preserving float <-> uint32 map (NaN is never produced by ord2f of a finite key)
__device__ __forceinline__ uint32_t f2ord(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// the slab's bar as the scan and the finalize kernels see it: the sampled value, floored at the logit pre-filter of
// the score threshold (at the floor the candidate list is simply "everything the threshold test can still pass")
__device__ __forceinline__ float slab_bar(const SelectArgs& A, int slab, bool& at_floor) {
  const float b = ord2f(~A.bar_inv[slab]);
  at_floor = !(b > A.xpre);
  return at_floor ? A.xpre : b;
}

// rank-th largest (rank >= 1, <= count) of `count` 16-bit keys in shared memory: two 8-bit radix passes with
// warp-aggregated histogram updates (sample keys share their exponent byte, plain shared atomics would serialise).
// Called by all NT threads; the result is valid in every thread.
template <int NT>
__device__ uint32_t kth_largest_u16(const uint16_t* keys, int count, int rank, unsigned* hist /*256*/, int* s_tmp /*2*/) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int rounds = (count + NT - 1) / NT;
  uint32_t prefix = 0;
  int need = rank;
  for (int pass = 0; pass < 2; ++pass) {
    for (int b = tid; b < 256; b += NT) hist[b] = 0u;
    __syncthreads();
    for (int rd = 0; rd < rounds; ++rd) {
      const int i = rd * NT + tid;
      unsigned bin = 0u;
      bool ok = false;
      if (i < count) {
        const uint32_t key = keys[i];
        ok = (pass == 0) || ((key >> 8) == prefix);
        bin = (pass == 0) ? (key >> 8) : (key & 255u);
      }
      warp_hist_add(hist, bin, ok);
    }
    __syncthreads();
    if (tid < 32) {
      unsigned sum = 0;
      for (int b = 0; b < 8; ++b) sum += hist[255 - 8 * lane - b];
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned before = inc - sum;
      if (before < (unsigned)need && inc >= (unsigned)need) {
        unsigned cum = before;
        for (int b = 0; b < 8; ++b) {
          const unsigned h = hist[255 - 8 * lane - b];
          if (cum < (unsigned)need && cum + h >= (unsigned)need) { s_tmp[0] = 255 - 8 * lane - b; s_tmp[1] = (int)cum; }
          cum += h;
        }
      }
    }
    __syncthreads();
    const int digit = s_tmp[0], above = s_tmp[1];
    __syncthreads();
    need -= above;
    if (pass == 0) prefix = (uint32_t)digit;
    else prefix = (prefix << 8) | (uint32_t)digit;
  }
  return prefix;
}

// rank-th largest for SMALL ranks (rank <= NT/2) without a pass over a histogram: the rank-th largest of the NT/2
// thread-pair maxima is a lower bound t0 of the answer (those maxima are distinct elements), the keys >= t0 -- about
// `rank` of them when the large keys are spread over the threads, which the strided ownership arranges -- are
// collected and ranked exactly by counting.  A plateau of equal keys that overflows the list falls back to the
// histogram passes.  Called by all NT threads; result valid in every thread.
constexpr int kSparseList = 1024;
template <int NT>
__device__ uint32_t kth_largest_sparse(const uint16_t* keys, int count, int rank, uint16_t* s_max /*NT/2*/,
                                       uint16_t* s_list /*kSparseList*/, int* s_n, unsigned* hist, int* s_tmp) {
  const int tid = threadIdx.x;
  if (rank > NT / 2) return kth_largest_u16<NT>(keys, count, rank, hist, s_tmp);
  uint32_t mx = 0u;
  for (int i = tid; i < count; i += NT) mx = max(mx, (uint32_t)keys[i]);
  mx = max(mx, __shfl_xor_sync(kFull, mx, 1));
  if ((tid & 1) == 0) s_max[tid >> 1] = (uint16_t)mx;
  if (tid == 0) *s_n = 0;
  __syncthreads();
  if (tid < NT / 2) {
    const uint32_t mine = s_max[tid];
    int above = 0;
    for (int s = 0; s < NT / 2; ++s) {
      const uint32_t v = s_max[s];
      above += (v > mine || (v == mine && s < tid)) ? 1 : 0;
    }
    if (above == rank - 1) s_tmp[0] = (int)mine;
  }
  __syncthreads();
  const uint32_t t0 = (uint32_t)s_tmp[0];
  for (int i = tid; i < count; i += NT) {
    const uint32_t key = keys[i];
    if (key >= t0) {
      const int o = atomicAdd(s_n, 1);
      if (o < kSparseList) s_list[o] = (uint16_t)key;
    }
  }
  __syncthreads();
  const int n = *s_n;
  if (n > kSparseList) return kth_largest_u16<NT>(keys, count, rank, hist, s_tmp);   // uniform branch
  for (int i = tid; i < n; i += NT) {
    const uint32_t v = s_list[i];
    int gt = 0, eq = 0;
    for (int j = 0; j < n; ++j) {
      const uint32_t u = s_list[j];
      gt += (u > v) ? 1 : 0;
      eq += (u == v) ? 1 : 0;
    }
    if (gt < rank && gt + eq >= rank) s_tmp[1] = (int)v;   // every writer writes the same value
  }
  __syncthreads();
  const uint32_t res = (uint32_t)s_tmp[1];
  __syncthreads();
  return res;
}

struct BarSmem {
  unsigned hist[256];
  uint16_t tmax[kBarThreads / 2];
  uint16_t list[kSparseList];
  int tmp[2];
  int cnt, n;
  bool last;
};

__global__ void __launch_bounds__(kBarThreads, 2) detect_bar_kernel(const SelectArgs A, const DetectLevels LV,
                                                                 const DetectSrc S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint16_t* skey = reinterpret_cast<uint16_t*>(smem_raw);   // kBarKeys
  __shared__ BarSmem sm;
  const int tid = threadIdx.x;
  grid_launch_dependents();
  const int p = blockIdx.x, l = blockIdx.y, n = blockIdx.z;
  const int P = LV.bar_parts[l];
  if (p >= P) return;
  const int slab_id = n * LV.num_levels + l;
  const int64_t E = (LV.off[l + 1] - LV.off[l]) * A.K;
  if (E <= 0) return;
  const float* slab = S.cls[l] + (int64_t)n * S.cls_stride[l];
  const int64_t pieces_all = (E + kBarPiece - 1) / kBarPiece;
  const bool exact = (LV.bar_rank[l] == 0);    // the sample is the whole slab: the bar is its k-th largest value
  const int npieces = LV.bar_pieces[l];
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(slab) & 15) == 0);
  // sampled piece g of the npieces*P sits at floor(g * pieces_all / (npieces*P)): a systematic sample of the slab
  const double step = exact ? 1.0 : (double)pieces_all / ((double)npieces * (double)P);

  // thread-global sample slot q = 4 floats; 8 consecutive slots (lanes) make one 128-byte piece
  constexpr int kSlots = kBarPiece / 4;
  const int nslots = npieces * kSlots;
  constexpr int kBatch = 8;   // 8 x 16 bytes in flight per thread
#pragma unroll 1
  for (int q0 = tid; q0 < nslots; q0 += kBatch * kBarThreads) {
    float4 t[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int q = q0 + i * kBarThreads;
      if (q < nslots) {
        const int j = q / kSlots, sub = q - j * kSlots;
        int64_t piece = exact ? (int64_t)j : (int64_t)((double)((int64_t)j * P + p) * step);
        if (piece >= pieces_all) piece = pieces_all - 1;
        const int64_t e = piece * kBarPiece + sub * 4;
        const float* src = slab + e;
        if (vec_ok && e + 4 <= E) {
          t[i] = ldg_stream4(src);
        } else {   // unaligned slab, or its ragged end
          t[i].x = (e + 0 < E) ? ldg_stream1(src + 0) : -INFINITY;
          t[i].y = (e + 1 < E) ? ldg_stream1(src + 1) : -INFINITY;
          t[i].z = (e + 2 < E) ? ldg_stream1(src + 2) : -INFINITY;
          t[i].w = (e + 3 < E) ? ldg_stream1(src + 3) : -INFINITY;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int q = q0 + i * kBarThreads;
      if (q < nslots) {
        const float v[4] = {t[i].x, t[i].y, t[i].z, t[i].w};
        uint32_t k[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) k[c] = (v[c] == v[c]) ? (f2ord(v[c]) >> 16) : 0u;   // NaN ranks last
        *reinterpret_cast<uint2*>(skey + (size_t)q * 4) = make_uint2(k[0] | (k[1] << 16), k[2] | (k[3] << 16));
      }
    }
  }
  __syncthreads();
  const int count = npieces * kBarPiece;
  uint32_t bar16;
  if (P == 1) {
    int rank = exact ? LV.k[l] : LV.bar_rank[l];
    if (rank > count) rank = count;
    bar16 = kth_largest_sparse<kBarThreads>(skey, count, rank, sm.tmax, sm.list, &sm.n, sm.hist, sm.tmp);
  } else {
    // pooled sample: contribute this CTA's kBarPoolKeys best keys; the slab's last sampling CTA ranks the pool
    const uint32_t t16 = kth_largest_sparse<kBarThreads>(skey, count, kBarPoolKeys, sm.tmax, sm.list, &sm.n, sm.hist,
                                                         sm.tmp);
    uint16_t* pool = A.pool + ((size_t)slab_id * kBarPoolParts + p) * kBarPoolKeys;
    if (tid == 0) sm.cnt = 0;
    __syncthreads();
    for (int i = tid; i < count; i += kBarThreads) {
      const uint32_t key = skey[i];
      if (key > t16) {
        const int o = atomicAdd(&sm.cnt, 1);   // fewer than kBarPoolKeys keys are strictly above the 64th largest
        pool[o] = (uint16_t)key;
      }
    }
    __syncthreads();
    for (int i = sm.cnt + tid; i < kBarPoolKeys; i += kBarThreads) pool[i] = (uint16_t)t16;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      sm.last = (atomicAdd(&A.pool_done[slab_id], 1u) == (unsigned)P - 1u);
    }
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    const uint16_t* all = A.pool + (size_t)slab_id * kBarPoolParts * kBarPoolKeys;
    const int tot = P * kBarPoolKeys;
    for (int i = tid; i < tot; i += kBarThreads) skey[i] = __ldcg(&all[i]);
    __syncthreads();
    int rank = LV.bar_rank[l];
    if (rank > kBarPoolKeys) rank = kBarPoolKeys;
    bar16 = kth_largest_sparse<kBarThreads>(skey, tot, rank, sm.tmax, sm.list, &sm.n, sm.hist, sm.tmp);
    if (tid == 0) A.pool_done[slab_id] = 0u;
  }
  // lower edge of the selected 16-bit bucket (a slightly lower bar only admits a few more candidates)
  if (tid == 0) A.bar_inv[slab_id] = ~(bar16 << 16);
}

// ---- scan ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scan_flush(uint64_t* stage, int cur, unsigned* gcount, uint64_t* gcand, int cap) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(gcount, (unsigned)cur);
  base = __shfl_sync(kFull, base, 0);
  for (int i = lane; i < cur; i += 32)
    if (base + (unsigned)i < (unsigned)cap) gcand[base + i] = stage[i];
  __syncwarp();
}

// all 32 lanes; v[0..3] are slab elements e0.., v[4..7] are e1..; `ok` masks elements outside the slab
__device__ __forceinline__ void scan_hits(const float* v, uint32_t e0, uint32_t e1, unsigned ok, float bar,
                                          uint64_t* stage, int& cur, unsigned* gcount, uint64_t* gcand, int cap) {
  const int lane = threadIdx.x & 31;
  unsigned flags = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) flags |= (v[j] >= bar) ? (1u << j) : 0u;
  flags &= ok;
  const int c = __popc(flags);
  const unsigned bm = __ballot_sync(kFull, c != 0);
  if (bm == 0u) return;
  const int tot = __reduce_add_sync(kFull, c);
  int pos;
  if (tot == __popc(bm)) {
    pos = __popc(bm & ((1u << lane) - 1u));          // at most one hit per lane (the usual case)
  } else {
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    pos = inc - c;
  }
  if (tot > kScanStage) {
    // a dense burst (the bar sits inside a plateau of equal values): straight to the list
    if (cur > 0) { scan_flush(stage, cur, gcount, gcand, cap); cur = 0; }
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(gcount, (unsigned)tot);
    base = __shfl_sync(kFull, base, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((flags >> j) & 1u) {
        const unsigned o = base + (unsigned)pos++;
        if (o < (unsigned)cap)
          gcand[o] = ((uint64_t)__float_as_uint(v[j]) << 32) | (uint64_t)((j < 4 ? e0 : e1 - 4u) + (uint32_t)j);
      }
    }
    return;
  }
  if (cur + tot > kScanStage) { scan_flush(stage, cur, gcount, gcand, cap); cur = 0; }
  pos += cur;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if ((flags >> j) & 1u)
      stage[pos++] = ((uint64_t)__float_as_uint(v[j]) << 32) | (uint64_t)((j < 4 ? e0 : e1 - 4u) + (uint32_t)j);
  cur += tot;
}

__global__ void __launch_bounds__(kScanThreads, 4) detect_scan_kernel(const SelectArgs A, const DetectLevels LV,
                                                                      const DetectSrc S) {
  __shared__ uint64_t s_stage[kScanThreads / 32][kScanStage];
  const int tid = threadIdx.x;
  grid_launch_dependents();
  grid_dependency_sync();     // the bars come from detect_bar_kernel
  const int n = blockIdx.y;
  int l = 0;
  while (l + 1 < LV.num_levels && (int)blockIdx.x >= LV.scan_base[l + 1]) ++l;
  const int chunk = blockIdx.x - LV.scan_base[l];
  const int slab_id = n * LV.num_levels + l;
  const int64_t E = (LV.off[l + 1] - LV.off[l]) * A.K;
  const float* slab = S.cls[l] + (int64_t)n * S.cls_stride[l];
  // virtual index v = element + mis, so that every v % 4 == 0 is a 16-byte aligned address
  const int mis = (int)((reinterpret_cast<uintptr_t>(slab) >> 2) & 3);
  const float* al = slab - mis;
  const int64_t vend = E + mis;
  const int64_t v0 = (int64_t)chunk * kScanChunk;
  if (v0 >= vend) return;
  const int64_t v1 = min(vend, v0 + (int64_t)kScanChunk);
  bool at_floor;
  const float bar = slab_bar(A, slab_id, at_floor);
  uint64_t* stage = s_stage[tid >> 5];
  unsigned* gcount = A.cand_count + slab_id;
  uint64_t* gcand = A.cand_list + (size_t)slab_id * A.cand_cap;
  int cur = 0;

  // full block iterations: every element of the 2048-wide block is inside the slab
  const int64_t f0 = (v0 >= mis) ? v0 : v0 + kScanIter;           // first full block start (v0 = 0 with mis > 0 is ragged)
  const int64_t nfull = (v1 - f0 >= kScanIter) ? (v1 - f0) / kScanIter : 0;
  auto ld = [&](int64_t vb, float* v) {
    const float* p = al + vb + (int64_t)tid * 4;
    const float4 t0 = ldg_stream4(p), t1 = ldg_stream4(p + kScanThreads * 4);
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
  };
  auto consume = [&](const float* v, int64_t vb) {
    const float m = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
    if (__any_sync(kFull, m >= bar)) {
      const uint32_t e0 = (uint32_t)(vb + (int64_t)tid * 4 - mis);
      scan_hits(v, e0, e0 + kScanThreads * 4, 0xffu, bar, stage, cur, gcount, gcand, A.cand_cap);
    }
  };
  auto ragged = [&](int64_t vb) {   // block partly outside [mis, vend): scalar loads, masked
    float v[8];
    unsigned ok = 0u;
    const int64_t q0 = vb + (int64_t)tid * 4, q1 = q0 + kScanThreads * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in0 = (q0 + j >= mis) && (q0 + j < v1), in1 = (q1 + j >= mis) && (q1 + j < v1);
      v[j] = in0 ? ldg_stream1(al + q0 + j) : -INFINITY;
      v[4 + j] = in1 ? ldg_stream1(al + q1 + j) : -INFINITY;
      ok |= (in0 ? (1u << j) : 0u) | (in1 ? (1u << (4 + j)) : 0u);
    }
    if (__any_sync(kFull, ok != 0u)) {
      const uint32_t e0 = (uint32_t)(q0 - mis);
      scan_hits(v, e0, e0 + kScanThreads * 4, ok, bar, stage, cur, gcount, gcand, A.cand_cap);
    }
  };
  if (f0 != v0) ragged(v0);
  {
    float va[8], vb[8];
    if (nfull > 0) ld(f0, va);
    for (int64_t it = 0; it < nfull; it += 2) {
      if (it + 1 < nfull) ld(f0 + (it + 1) * kScanIter, vb);
      consume(va, f0 + it * kScanIter);
      if (it + 1 < nfull) {
        if (it + 2 < nfull) ld(f0 + (it + 2) * kScanIter, va);
        consume(vb, f0 + (it + 1) * kScanIter);
      }
    }
  }
  for (int64_t vb = f0 + nfull * kScanIter; vb < v1; vb += kScanIter) ragged(vb);
  if (cur > 0) scan_flush(stage, cur, gcount, gcand, A.cand_cap);
}

// ---- finalize --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) detect_finalize_kernel(const SelectArgs A, const DetectLevels LV,
                                                                      const DetectSrc S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);            // kCandCap
  uint64_t* srt = buf + kCandCap;                                   // kFinSort
  __shared__ unsigned hist[kBins];
  __shared__ int s_tmp[4];
  __shared__ int s_count, s_count2;
  __shared__ uint64_t s_red64[kFinThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  grid_launch_dependents();
  grid_dependency_sync();     // the candidate lists come from detect_scan_kernel
  const int l = blockIdx.x, n = blockIdx.y;
  const int slab_id = n * LV.num_levels + l;
  const int64_t E = (LV.off[l + 1] - LV.off[l]) * A.K;
  const int k = LV.k[l];
  if (E <= 0 || k <= 0) {
    if (tid == 0) { A.lvl_count[slab_id] = 0; A.status[slab_id] = kStatusOk; }
    return;
  }
  const unsigned raw = A.cand_count[slab_id];
  if (raw > (unsigned)A.cand_cap) {          // more hits than the list holds: the streaming kernel redoes the slab
    if (tid == 0) A.status[slab_id] = kStatusFallback;
    return;
  }
  bool at_floor;
  const float bar = slab_bar(A, slab_id, at_floor);
  if (tid == 0) { s_count = 0; s_count2 = 0; }
  __syncthreads();
  // exact scores of the candidates; `score > threshold` (retinanet.py:494) applied here
  const uint64_t* gcand = A.cand_list + (size_t)slab_id * A.cand_cap;
  const int cnt = (int)raw;
  for (int i0 = 0; i0 < cnt; i0 += kFinThreads) {
    const int i = i0 + tid;
    bool p = false;
    uint64_t key = 0ull;
    if (i < cnt) {
      const uint64_t e = gcand[i];
      const float sc = sigmoid_score(__uint_as_float((uint32_t)(e >> 32)));
      p = sc > A.thr;
      key = make_key(sc, ref_index(S, l, (uint32_t)e));
    }
    const unsigned m = __ballot_sync(kFull, p);
    if (m != 0u) {
      int base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(&s_count, __popc(m));
      base = __shfl_sync(kFull, base, leader);
      if (p) buf[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const int valid = s_count;
  if (!at_floor && valid < k) {              // the sampled bar was too high
    if (tid == 0) A.status[slab_id] = kStatusFallback;
    return;
  }
  uint64_t T = 0ull;
  if (valid > k) T = select_kth<kFinThreads>(buf, valid, k, hist, s_tmp, 0);
  for (int i0 = 0; i0 < valid; i0 += kFinThreads) {
    const int i = i0 + tid;
    const bool p = (i < valid) && (buf[i] >= T);
    const unsigned m = __ballot_sync(kFull, p);
    if (m != 0u) {
      int base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(&s_count2, __popc(m));
      base = __shfl_sync(kFull, base, leader);
      if (p) srt[base + __popc(m & ((1u << lane) - 1u))] = buf[i];
    }
  }
  __syncthreads();
  const int fin = s_count2;   // == min(valid, k): keys are distinct
  if (!at_floor) {
    // Proof obligation: nothing outside the candidate list can rank among the k best.  A non-candidate has
    // logit < bar, so its computed score is at most sigmoid(bar) up to the rounding of expf / the division
    // (a few ulp; 4e-6 relative is generous).  Ties at the k-th score therefore cannot involve a non-candidate.
    const uint64_t mn = block_min_u64<kFinThreads>(srt, fin, s_red64);
    const float ub = sigmoid_score(bar) * 1.000004f;
    if (!(key_score(mn) > ub)) {
      if (tid == 0) A.status[slab_id] = kStatusFallback;
      return;
    }
  }
  int m = 1;
  while (m < fin) m <<= 1;
  for (int i = fin + tid; i < m; i += kFinThreads) srt[i] = ~0ull;
  __syncthreads();
  bitonic_asc<kFinThreads>(srt, m);   // two keys per thread in registers, shuffles below stride 64; padding sorts last
  const int64_t cbase = (int64_t)n * LV.num_levels * A.topk + (int64_t)l * A.topk;
  for (int t = tid; t < fin; t += kFinThreads) {
    const uint64_t key = srt[fin - 1 - t];           // descending: score, then index ascending
    const uint32_t idx = key_index(key);
    const uint32_t a = idx / (uint32_t)A.K;          // retinanet.py:498-499
    const int c = (int)(idx - a * (uint32_t)A.K);
    const int64_t r = LV.off[l] + a;
    const float4 d = load_delta(S, l, n, a);
    const float4 an = A.anchors[(int64_t)n * A.anchor_stride4 + r];
    A.cand_box[cbase + t] = decode_box3(d, an, A.wx, A.wy, A.ww, A.wh, A.clampv);
    A.cand_score[cbase + t] = key_score(key);
    A.cand_class[cbase + t] = c;
  }
  if (tid == 0) { A.lvl_count[slab_id] = fin; A.status[slab_id] = kStatusOk; }
}

constexpr size_t kBarSmem = (size_t)kBarKeys * 2;
constexpr size_t kFinSmem = (size_t)(kCandCap + kFinSort) * 8;

static DetectLevels plan_levels(const int64_t* off, int num_levels, int K, int topk, int N) {
  DetectLevels lv;
  lv.num_levels = num_levels;
  int base = 0, maxp = 1, sbase = 0;
  const int cap_parts = kSelCap / (topk > 0 ? topk : 1);
  for (int l = 0; l < num_levels; ++l) {
    lv.off[l] = off[l];
    const int64_t hwa = off[l + 1] - off[l];
    const int64_t E = hwa * K;
    lv.k[l] = (int)(hwa < topk ? hwa : topk);
    int parts = (int)ceil_div(E > 0 ? E : 1, (int64_t)64 * kSelIter);  // ~256K elements per CTA
    // a longer stream per CTA tightens its running threshold (pass rate ~ k/n_seen) and needs fewer prunes:
    // only split a slab as far as one wave of resident CTAs (4 per SM) can take
    int want = (SEL_CTAS * 148) / (N * num_levels);   // 4 resident CTAs per SM: keep every stream in the first wave
    if (want < 1) want = 1;
    if (parts > want) parts = want;
    if (parts > 16) parts = 16;
    if (parts > cap_parts) parts = cap_parts;
    if (parts < 1) parts = 1;
    int64_t plen = ceil_div(ceil_div(E > 0 ? E : 1, parts), kSelIter) * kSelIter;
    parts = (int)ceil_div(E > 0 ? E : 1, plen);
    lv.nparts[l] = parts;
    lv.part_len[l] = plen;
    lv.part_base[l] = base;
    base += parts;
    if (parts > maxp) maxp = parts;
    // sampled-bar path: scan chunks (+3: a slab that does not start on a 16-byte boundary is scanned from the
    // boundary below it) and the sampling plan
    lv.scan_base[l] = sbase;
    sbase += (int)ceil_div(E > 0 ? E + 3 : 0, (int64_t)kScanChunk);
    {
      // sample size: enough for the bar to sit at rank ~kBarRank of the sample (the number of slab elements above the
      // rank-r sample value has a relative spread of 1/sqrt(r)), at most kBarKeys per CTA; slabs too large for one
      // CTA's sample at that rank pool the best keys of several sampling CTAs
      constexpr int kPerCta = kBarKeys / kBarPiece;
      const double want_hits = (double)kBarOver * (lv.k[l] > 0 ? lv.k[l] : 1);
      const int64_t pieces_all = ceil_div(E > 0 ? E : 1, (int64_t)kBarPiece);
      const int64_t need = (int64_t)ceil((double)kBarRank * (double)(E > 0 ? E : 1) / (want_hits * kBarPiece));
      int P = 1;
      int64_t pieces;
      if (need >= pieces_all && pieces_all <= kPerCta) {
        pieces = pieces_all;                 // the sample is the whole slab: the bar is its k-th largest value
        lv.bar_rank[l] = 0;
      } else {
        pieces = need < pieces_all ? need : pieces_all;
        if (pieces > kPerCta) {
          P = (int)ceil_div(pieces, (int64_t)kPerCta);
          if (P > kBarPoolParts) P = kBarPoolParts;
          while (P & (P - 1)) ++P;
          pieces = ceil_div(pieces, (int64_t)P);
          if (pieces > kPerCta) pieces = kPerCta;
        }
        int rank = (int)ceil(want_hits * (double)pieces * kBarPiece * P / (double)E);
        if (rank < 4) rank = 4;              // (tiny top-k on a big slab: a rank-1 bar would be too volatile)
        if (rank > kBarPoolKeys && P > 1) rank = kBarPoolKeys;
        lv.bar_rank[l] = rank;
      }
      lv.bar_parts[l] = P;
      lv.bar_pieces[l] = (int)pieces;
    }
  }
  lv.off[num_levels] = off[num_levels];
  lv.scan_base[num_levels] = sbase;
  for (int l = num_levels; l < kMaxLevels; ++l) {
    lv.nparts[l] = 0; lv.part_base[l] = base; lv.k[l] = 0; lv.part_len[l] = 0;
    lv.scan_base[l + 1] = sbase; lv.bar_parts[l] = 0; lv.bar_rank[l] = 0; lv.bar_pieces[l] = 0;
  }
  lv.total_parts = base;
  lv.max_parts = maxp;
  return lv;
}

struct DetectWs {
  size_t off_done, off_pcount, off_pkeys, off_lvl, off_cbox, off_cscore, off_ccls, off_nms, total;
  size_t off_ccount, off_pooldone, off_zero_end, off_bar, off_status, off_pool, off_clist;
};
static DetectWs detect_ws_layout(int N, int num_levels, int topk, int max_parts, int max_det = 0) {
  DetectWs w;
  size_t o = 0;
  const size_t slabs = (size_t)N * num_levels;
  // zero-initialised by ONE memset: done, lvl_count, cand_count, pool_done and the head of the NMS workspace (its
  // completion counters and survivor bitmaps), which is why that workspace sits here and not at the end
  w.off_done = o;     o += align_up(sizeof(unsigned) * slabs, 16);
  w.off_lvl = o;      o += align_up(sizeof(int) * slabs, 16);
  w.off_ccount = o;   o += align_up(sizeof(unsigned) * slabs, 16);
  w.off_pooldone = o; o += align_up(sizeof(unsigned) * slabs, 16);
  {
    const NmsWs nw = nms_ws_layout(N, 8, max_det > 0 ? max_det : 1024);
    w.off_nms = o;
    w.off_zero_end = o + nw.off_cnt;   // (nw.off_done == 0: the zeroed part of the NMS workspace comes first)
    o += align_up(nw.total, 16);
  }
  w.off_bar = o;      o += align_up(sizeof(unsigned) * slabs, 16);
  w.off_status = o;   o += align_up(sizeof(int) * slabs, 16);
  w.off_pcount = o; o += align_up(sizeof(int) * slabs * max_parts, 16);
  w.off_pkeys = o;  o += align_up(sizeof(uint64_t) * slabs * max_parts * topk, 16);
  w.off_cbox = o;   o += align_up(sizeof(float4) * slabs * topk, 16);
  w.off_cscore = o; o += align_up(sizeof(float) * slabs * topk, 16);
  w.off_ccls = o;   o += align_up(sizeof(int64_t) * slabs * topk, 16);
  w.off_pool = o;   o += align_up(sizeof(uint16_t) * slabs * kBarPoolParts * kBarPoolKeys, 16);
  w.off_clist = o;  o += align_up(sizeof(uint64_t) * slabs * kCandCap, 16);
  w.total = o;
  return w;
}

static int detect_max_parts(int topk) {
  int maxp = kSelCap / topk;
  if (maxp > 16) maxp = 16;
  if (maxp < 1) maxp = 1;
  return maxp;
}

// FSG_DETECT_LEGACY=1: run only the streaming top-k kernel on every slab (the round-1 select stage; A/B timing)
static bool detect_legacy_mode() {
  static const int v = [] {
    const char* e = getenv("FSG_DETECT_LEGACY");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v != 0;
}

struct DetectCall {
  int N, K, num_levels, topk, max_det;
  int64_t R;
  const int64_t* h_level_offsets;
  const float* anchors;
  int64_t anchor_image_stride;
  float score_threshold;
  double nms_threshold;
  const float* h_box_weights;
  float scale_clamp;
  float* out_boxes; float* out_scores; int64_t* out_classes; int32_t* out_count;
  float* cand_boxes; float* cand_scores; int64_t* cand_classes; int32_t* cand_count;
  int64_t* keep_idx;
  const float* postprocess;
  void* workspace; size_t workspace_bytes;
};

static int detect_run(const DetectCall& c, const DetectSrc& src, cudaStream_t s) {
  const int N = c.N, K = c.K, num_levels = c.num_levels, topk = c.topk, max_det = c.max_det;
  if (c.postprocess && ((uintptr_t)c.postprocess & 15)) return FSG_ERR_INVALID_ARG;
  if (!c.anchors || !c.out_boxes || !c.out_scores || !c.out_classes || !c.out_count || !c.h_box_weights)
    return FSG_ERR_INVALID_ARG;
  if (topk <= 0 || max_det <= 0) return FSG_ERR_INVALID_ARG;
  if (max_det > 1024) return FSG_ERR_UNSUPPORTED;
  if (c.anchor_image_stride % 4 != 0) return FSG_ERR_INVALID_ARG;
  if (topk > kSelTrigger || (int64_t)num_levels * topk > kNmsCap || K > 65535 || N > 65535)
    return FSG_ERR_UNSUPPORTED;
  for (int l = 0; l < num_levels; ++l) {
    const int64_t hwa = c.h_level_offsets[l + 1] - c.h_level_offsets[l];
    if (hwa < 0 || hwa * K >= ((int64_t)1 << 32) - 8) return FSG_ERR_UNSUPPORTED;
  }
  if ((c.cand_boxes || c.cand_scores || c.cand_classes) && !(c.cand_boxes && c.cand_scores && c.cand_classes))
    return FSG_ERR_INVALID_ARG;
  DetectLevels lv = plan_levels(c.h_level_offsets, num_levels, K, topk, N);
  const int maxp = detect_max_parts(topk);
  lv.max_parts = maxp;
  const DetectWs w = detect_ws_layout(N, num_levels, topk, maxp);
  if (!c.workspace || c.workspace_bytes < w.total || ((uintptr_t)c.workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)c.workspace;
  FSG_CUDA_TRY(cudaMemsetAsync(ws + w.off_done, 0, w.off_zero_end - w.off_done, s));

  SelectArgs sa;
  sa.anchors = (const float4*)c.anchors;
  sa.anchor_stride4 = c.anchor_image_stride / 4; sa.R = c.R; sa.K = K; sa.topk = topk;
  sa.thr = c.score_threshold;
  {
    // logit of the score threshold, minus a safety margin (the exact `score > thr` test follows)
    const double t = (double)c.score_threshold;
    double lg = (t <= 0.0) ? -INFINITY : ((t >= 1.0) ? INFINITY : log(t / (1.0 - t)));
    sa.xpre = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
    if (t <= 0.0) sa.xpre = -INFINITY;
  }
  sa.wx = c.h_box_weights[0]; sa.wy = c.h_box_weights[1]; sa.ww = c.h_box_weights[2]; sa.wh = c.h_box_weights[3];
  sa.clampv = c.scale_clamp;
  sa.part_keys = (uint64_t*)(ws + w.off_pkeys); sa.part_count = (int*)(ws + w.off_pcount);
  sa.done = (unsigned*)(ws + w.off_done);
  sa.cand_box = (float4*)(ws + w.off_cbox); sa.cand_score = (float*)(ws + w.off_cscore);
  sa.cand_class = (int64_t*)(ws + w.off_ccls); sa.lvl_count = (int*)(ws + w.off_lvl);
  sa.bar_inv = (unsigned*)(ws + w.off_bar); sa.cand_count = (unsigned*)(ws + w.off_ccount);
  sa.cand_list = (uint64_t*)(ws + w.off_clist); sa.pool = (uint16_t*)(ws + w.off_pool);
  sa.pool_done = (unsigned*)(ws + w.off_pooldone); sa.status = (int*)(ws + w.off_status);
  sa.cand_cap = kCandCap;

  const bool legacy = detect_legacy_mode();
  // all memsets first, so that the kernels form one chain under programmatic dependent launch (each kernel is
  // scheduled while its predecessor drains and waits on the device before it reads what that one wrote)
  const int nms_split = nms_split_for(N);
  const NmsWs nms_w = nms_ws_layout(N, nms_split, max_det);   // (its zeroed head is part of the memset above)
  if (!legacy) {
    int pmax = 1;
    for (int l = 0; l < num_levels; ++l) pmax = lv.bar_parts[l] > pmax ? lv.bar_parts[l] : pmax;
    FSG_CUDA_TRY(cudaFuncSetAttribute(detect_bar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBarSmem));
    detect_bar_kernel<<<dim3((unsigned)pmax, (unsigned)num_levels, (unsigned)N), kBarThreads, kBarSmem, s>>>(sa, lv, src);
    FSG_LAUNCH_CHECK();
    if (lv.scan_base[num_levels] > 0) {
      launch_pdl(detect_scan_kernel, dim3((unsigned)lv.scan_base[num_levels], (unsigned)N), dim3(kScanThreads), 0, s,
                 true, sa, lv, src);
      FSG_LAUNCH_CHECK();
    }
    FSG_CUDA_TRY(cudaFuncSetAttribute(detect_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmem));
    launch_pdl(detect_finalize_kernel, dim3((unsigned)num_levels, (unsigned)N), dim3(kFinThreads), kFinSmem, s, true,
               sa, lv, src);
    FSG_LAUNCH_CHECK();
  } else {
    sa.status = nullptr;
  }
  FSG_CUDA_TRY(cudaFuncSetAttribute(detect_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmem));
  launch_pdl(detect_select_kernel, dim3((unsigned)lv.total_parts, (unsigned)N), dim3(kSelThreads), kSelSmem, s, !legacy,
             sa, lv, src);
  FSG_LAUNCH_CHECK();

  NmsArgs a = {};
  a.boxes = sa.cand_box; a.scores = sa.cand_score; a.classes = sa.cand_class;
  a.slots_per_image = (int64_t)num_levels * topk; a.lvl_count = sa.lvl_count; a.L = num_levels; a.topk = topk;
  a.fixed_count = 0; a.thr = threshold_floor(c.nms_threshold); a.max_out = max_det;
  {
    char* nws = ws + w.off_nms;
    a.split = nms_split; a.part_cap = max_det;
    a.part_keys = (uint64_t*)(nws + nms_w.off_keys); a.part_cnt = (int*)(nws + nms_w.off_cnt);
    a.done = (unsigned*)(nws + nms_w.off_done);
    a.sorted_runs = 1;   // every level's candidates leave the select stage sorted by (score, index)
    a.alive = (unsigned*)(nws + nms_w.off_alive); a.rank2cand = (uint16_t*)(nws + nms_w.off_r2c);
  }
  a.keep = c.keep_idx; a.keep_stride = max_det; a.num_keep = c.out_count;
  a.out_boxes = (float4*)c.out_boxes; a.out_scores = c.out_scores; a.out_classes = c.out_classes;
  a.post = (const float4*)c.postprocess;
  a.exp_boxes = (float4*)c.cand_boxes; a.exp_scores = c.cand_scores; a.exp_classes = c.cand_classes;
  a.exp_count = c.cand_count;
  return launch_nms_image(a, N, s, true);
}

}  // namespace fsg

using namespace fsg;

extern "C" size_t fsg_detect_workspace_bytes(int N, int64_t R, int K, int num_levels, int topk) {
  if (N <= 0 || num_levels <= 0 || num_levels > kMaxLevels || topk <= 0 || K <= 0) return 0;
  (void)R;
  return detect_ws_layout(N, num_levels, topk, detect_max_parts(topk)).total;
}

extern "C" size_t fsg_detect_status_offset(int N, int num_levels, int topk) {
  if (N <= 0 || num_levels <= 0 || num_levels > kMaxLevels || topk <= 0) return 0;
  return detect_ws_layout(N, num_levels, topk, detect_max_parts(topk)).off_status;
}

extern "C" int fsg_detect(const float* logits, const float* deltas, const float* anchors,
                          int64_t anchor_image_stride, int N, int64_t R, int K, const int64_t* h_level_offsets,
                          int num_levels, float score_threshold, int topk, double nms_threshold, int max_det,
                          const float* h_box_weights, float scale_clamp, float* out_boxes, float* out_scores,
                          int64_t* out_classes, int32_t* out_count, float* cand_boxes, float* cand_scores,
                          int64_t* cand_classes, int32_t* cand_count, int64_t* keep_idx, const float* postprocess,
                          void* workspace, size_t workspace_bytes, fsg_stream_t stream) {
  if (N <= 0 || R <= 0 || K <= 0 || !h_level_offsets || num_levels <= 0 || num_levels > kMaxLevels)
    return FSG_ERR_INVALID_ARG;
  if (!logits || !deltas) return FSG_ERR_INVALID_ARG;
  if (h_level_offsets[0] != 0 || h_level_offsets[num_levels] != R) return FSG_ERR_INVALID_ARG;
  if (((uintptr_t)deltas & 15)) return FSG_ERR_INVALID_ARG;
  DetectSrc src = {};
  src.A = 1; src.K = K; src.native = 0;
  for (int l = 0; l < num_levels; ++l) {
    src.cls[l] = logits + h_level_offsets[l] * K;
    src.reg[l] = deltas + h_level_offsets[l] * 4;
    src.cls_stride[l] = R * K;
    src.reg_stride[l] = R * 4;
    src.HW[l] = 0;
  }
  DetectCall c = {N, K, num_levels, topk, max_det, R, h_level_offsets, anchors, anchor_image_stride, score_threshold,
                  nms_threshold, h_box_weights, scale_clamp, out_boxes, out_scores, out_classes, out_count, cand_boxes,
                  cand_scores, cand_classes, cand_count, keep_idx, postprocess, workspace, workspace_bytes};
  return detect_run(c, src, (cudaStream_t)stream);
}

extern "C" int fsg_detect_levels(const fsg_detect_level* h_levels, int num_levels, int A, int K, const float* anchors,
                                 int64_t anchor_image_stride, int N, int64_t R, float score_threshold, int topk,
                                 double nms_threshold, int max_det, const float* h_box_weights, float scale_clamp,
                                 float* out_boxes, float* out_scores, int64_t* out_classes, int32_t* out_count,
                                 float* cand_boxes, float* cand_scores, int64_t* cand_classes, int32_t* cand_count,
                                 int64_t* keep_idx, const float* postprocess, void* workspace,
                                 size_t workspace_bytes, fsg_stream_t stream) {
  if (N <= 0 || R <= 0 || K <= 0 || A <= 0 || !h_levels || num_levels <= 0 || num_levels > kMaxLevels)
    return FSG_ERR_INVALID_ARG;
  DetectSrc src = {};
  src.A = A; src.K = K; src.native = 1;
  int64_t offs[kMaxLevels + 1];
  offs[0] = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int64_t hw = (int64_t)h_levels[l].H * h_levels[l].W;
    if (h_levels[l].H < 0 || h_levels[l].W < 0 || hw > (1 << 30)) return FSG_ERR_INVALID_ARG;
    if (hw > 0 && (!h_levels[l].logits || !h_levels[l].deltas)) return FSG_ERR_INVALID_ARG;
    src.cls[l] = h_levels[l].logits;
    src.reg[l] = h_levels[l].deltas;
    src.cls_stride[l] = (int64_t)A * K * hw;
    src.reg_stride[l] = (int64_t)A * 4 * hw;
    src.HW[l] = (int)hw;
    offs[l + 1] = offs[l] + hw * A;
  }
  if (offs[num_levels] != R) return FSG_ERR_INVALID_ARG;
  DetectCall c = {N, K, num_levels, topk, max_det, R, offs, anchors, anchor_image_stride, score_threshold,
                  nms_threshold, h_box_weights, scale_clamp, out_boxes, out_scores, out_classes, out_count, cand_boxes,
                  cand_scores, cand_classes, cand_count, keep_idx, postprocess, workspace, workspace_bytes};
  return detect_run(c, src, (cudaStream_t)stream);
}
