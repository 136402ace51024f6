// K3: score scan + per-level top-k + threshold + box decode + per-class NMS + top detections.
//
// Reference semantics (paths relative to the reference tree):
//   detectron2/modeling/meta_arch/retinanet.py:460-520  RetinaNet.inference_single_image
//   detectron2/modeling/box_regression.py:69-107        Box2BoxTransform.apply_deltas
//   detectron2/layers/nms.py:6,9-26                     nms / batched_nms (torchvision greedy NMS,
//                                                        per-class un-offset form)
//
// Where the reference sorts every level's H*W*A*K scores to keep 1000, the scan kernel streams the
// logits once (coalesced 16-byte loads), keeps a running exact top-k candidate buffer in shared memory
// per CTA (radix-select pruning raises a logit pre-filter as it goes), and the last CTA of each
// (image, level) slab merges, sorts and decodes.  One CTA per image then runs the per-class greedy
// NMS entirely in shared memory (bitonic sort on composite keys, warp-per-class suppression) and
// emits the final detections -- no host round trip, no n^2 mask in global memory.
#include <math.h>

#include "common.cuh"
#include "nms_large.cuh"

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
constexpr int kMaxLevels = 8;
constexpr int kNmsThreads = 1024;
constexpr int kNmsBigSeg = 512;                     // class segments above this size are suppressed by the whole CTA
constexpr int kNmsCap = 8192;                       // candidates per image the NMS kernel holds

struct DetectLevels {
  int64_t off[kMaxLevels + 1];  // anchor offsets of the levels
  int nparts[kMaxLevels];       // CTAs per (image, level) slab
  int part_base[kMaxLevels];    // first blockIdx.x of the level
  int k[kMaxLevels];            // min(topk, HWA_l)
  int64_t part_len[kMaxLevels]; // elements per part (multiple of kSelIter)
  int num_levels;
  int total_parts;
  int max_parts;                // slot stride (parts) per level in the scratch arrays
};

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
      unsigned bin = 0x80000000u | (unsigned)lane;   // unique sentinel: lanes without a key match nobody
      if (i < count) {
        const uint64_t key = buf[i];
        const bool match = (top == 64) || ((key >> top) == (prefix >> top));
        if (match) bin = (unsigned)(key >> shift) & ((1u << width) - 1u);
      }
      const unsigned peers = __match_any_sync(kFull, bin);
      if (!(bin & 0x80000000u) && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
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
template <int NT>
__device__ void bitonic_desc(uint64_t* a, int m) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += NT) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t x = a[lo], y = a[hi];
        if (desc ? (x < y) : (x > y)) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

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
  const float* logits;
  const float4* deltas;
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
};

// Evaluate the warp's staged raw candidates as dense batches: sigmoid (expf + IEEE divide), the exact
// `score > ts` test and the append to the CTA's key buffer.  Called by all 32 lanes.
__device__ __forceinline__ void flush_stage(const float2* stage, int n, float ts, uint64_t* buf, int* s_count) {
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
      if (p) buf[base + __popc(m & ((1u << lane) - 1u))] = make_key(sc, id);
    }
  }
}

// Collect the lanes' logits that pass the (cheap, conservative) logit pre-filter into the warp's staging
// queue; the expensive part runs later on full warps (flush_stage).  v[0..3] sit at idx0.., v[4..7] at idx1..
__device__ __forceinline__ void stage_candidates(const float* v, uint32_t idx0, uint32_t idx1, float xb, float ts,
                                                 float2* stage, int* s_scnt, uint64_t* buf, int* s_count) {
  const int lane = threadIdx.x & 31;
  unsigned flags = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) flags |= (v[j] > xb) ? (1u << j) : 0u;
  const int c = __popc(flags);
  const int tot = __reduce_add_sync(kFull, c);
  if (tot == 0) return;
  int cur = *s_scnt;
  if (cur + tot > kStagePerWarp) {
    flush_stage(stage, cur, ts, buf, s_count);
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
      flush_stage(stage, __popc(m), ts, buf, s_count);
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

__global__ void __launch_bounds__(kSelThreads, SEL_CTAS) detect_select_kernel(const SelectArgs A, const DetectLevels LV) {
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
  const int n = blockIdx.y;
  int l = 0;
  while (l + 1 < LV.num_levels && (int)blockIdx.x >= LV.part_base[l + 1]) ++l;
  const int part = blockIdx.x - LV.part_base[l];
  const int k = LV.k[l];
  const int64_t hwa = LV.off[l + 1] - LV.off[l];
  const int64_t E = hwa * A.K;                         // elements in the slab
  const float* slab = A.logits + ((int64_t)n * A.R + LV.off[l]) * A.K;
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
      stage_candidates(v, p0, p0 + kSelThreads * 4, xb, ts, stage, &s_scnt[tid >> 5], buf, &s_count);
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
  flush_stage(stage, s_scnt[tid >> 5], s_ts, buf, &s_count);   // what is still waiting in the warp queues
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
    const float4 d = A.deltas[(int64_t)n * A.R + r];
    const float4 an = A.anchors[(int64_t)n * A.anchor_stride4 + r];
    A.cand_box[cbase + t] = decode_box3(d, an, A.wx, A.wy, A.ww, A.wh, A.clampv);
    A.cand_score[cbase + t] = key_score(key);
    A.cand_class[cbase + t] = c;
  }
  if (tid == 0) A.lvl_count[n * LV.num_levels + l] = fin;
}

// ------------------------------------------------------------------------------------------
// NMS: one CTA per image (or per stand-alone call)
// ------------------------------------------------------------------------------------------
struct NmsArgs {
  const float4* boxes;      // per image: slots_per_image entries
  const float* scores;
  const int64_t* classes;   // may be NULL (single class)
  int64_t slots_per_image;
  const int* lvl_count;     // (N, L) or NULL -> fixed_count
  int L;
  int topk;                 // slot stride per level
  int fixed_count;
  float thr;                // largest float <= the double threshold (strict > compare, see fsg_nms)
  int max_out;              // truncate to this many (DETECTIONS_PER_IMAGE); <= 0: all
  int split;                // CTAs per image; CTA c owns the classes with class % split == c
  int part_cap;             // survivors each CTA hands to the merge (max_out, or the candidate count)
  uint64_t* part_keys;      // (N, split, part_cap) scratch
  int* part_cnt;            // (N, split)
  unsigned* done;           // (N)
  // outputs
  int64_t* keep;            // (N, keep_stride) candidate indices in concatenation order
  int64_t keep_stride;
  int32_t* num_keep;        // (N)
  float4* out_boxes;        // (N, max_out) or NULL
  float* out_scores;
  int64_t* out_classes;
  const float4* post;       // (N) [scale_x, scale_y, clip_w, clip_h] or NULL: detector_postprocess fused in
  // optional compact export of the candidates
  float4* exp_boxes;        // (N, L*topk)
  float* exp_scores;
  int64_t* exp_classes;
  int32_t* exp_count;
};

// Boxes.scale (boxes.py:205-210: fp32 * fp32(scale)) then Boxes.clip (boxes.py:122-136: clamp(min=0, max=size));
// pp = [scale_x, scale_y, clip_w, clip_h]
__device__ __forceinline__ float4 postprocess_box(float4 b, float4 pp) {
  b.x = fminf(fmaxf(__fmul_rn(b.x, pp.x), 0.f), pp.z);
  b.y = fminf(fmaxf(__fmul_rn(b.y, pp.y), 0.f), pp.w);
  b.z = fminf(fmaxf(__fmul_rn(b.z, pp.x), 0.f), pp.z);
  b.w = fminf(fmaxf(__fmul_rn(b.w, pp.y), 0.f), pp.w);
  return b;
}

// ascending bitonic sort of m (power of two) keys in shared memory, block-wide
template <int NT>
__device__ void bitonic_asc_smem(uint64_t* a, int m) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += NT) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const uint64_t x = a[lo], y = a[hi];
        if (asc ? (x > y) : (x < y)) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

// The same network with two keys per thread held in registers (64 <= m <= 2*NT): compare-exchanges with a partner
// up to 32 elements away run on warp shuffles, only the strides >= 64 go through shared memory and a barrier --
// 20 barriers instead of 66 for 2048 keys.
template <int NT>
__device__ void bitonic_asc(uint64_t* a, int m) {
  if (m < 64 || m > 2 * NT) {
    bitonic_asc_smem<NT>(a, m);
    return;
  }
  const int t = threadIdx.x;
  const bool act = t < (m >> 1);
  const int i0 = 2 * t;
  uint64_t v0 = act ? a[i0] : 0ull, v1 = act ? a[i0 + 1] : 0ull;
  for (int size = 2; size <= m; size <<= 1) {
    const bool asc = ((i0 & size) == 0);
    int stride = size >> 1;
    if (stride >= 64) {
      if (act) { a[i0] = v0; a[i0 + 1] = v1; }
      __syncthreads();
      for (; stride >= 64; stride >>= 1) {
        if (act) {
          const int lo = ((t / stride) * (stride << 1)) + (t % stride);
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const uint64_t x = a[lo], y = a[hi];
          if (up ? (x > y) : (x < y)) { a[lo] = y; a[hi] = x; }
        }
        __syncthreads();
      }
      if (act) { v0 = a[i0]; v1 = a[i0 + 1]; }
    }
    for (; stride >= 2; stride >>= 1) {
      const int pl = stride >> 1;   // partner thread = t ^ (stride / 2): same warp for stride <= 32
      const uint64_t p0 = __shfl_xor_sync(kFull, (unsigned long long)v0, pl);
      const uint64_t p1 = __shfl_xor_sync(kFull, (unsigned long long)v1, pl);
      const bool keep_min = (((i0 & stride) == 0) == asc);
      v0 = keep_min ? min(v0, p0) : max(v0, p0);
      v1 = keep_min ? min(v1, p1) : max(v1, p1);
    }
    if ((v0 > v1) == asc) { const uint64_t tmp = v0; v0 = v1; v1 = tmp; }
  }
  if (act) { a[i0] = v0; a[i0 + 1] = v1; }
  __syncthreads();
}

// Per-class NMS is independent across classes, so an image is split over `split` CTAs by class id; each
// sorts and suppresses only its own candidates (4x fewer keys per bitonic network at split = 4) and hands
// its best survivors to the last CTA of the image, which merges them by score.
__global__ void __launch_bounds__(kNmsThreads) nms_image_kernel(const NmsArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                      // kNmsCap * 8
  float4* sbox = reinterpret_cast<float4*>(smem_raw + (size_t)kNmsCap * 8);     // kNmsCap * 16
  uint16_t* seg = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kNmsCap * 24); // kNmsCap * 2
  unsigned char* dead = smem_raw + (size_t)kNmsCap * 26;                        // kNmsCap
  __shared__ int s_pref[kMaxLevels + 1];
  __shared__ int s_warp[kNmsThreads / 32];
  __shared__ int s_nseg, s_next, s_nkeep, s_mine, s_nbatch;
  __shared__ int s_batch[32];
  __shared__ bool s_last;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int part = blockIdx.x;
  const int n = blockIdx.y;
  const int S = A.split;
  if (tid == 0) {
    s_pref[0] = 0;
    if (A.lvl_count) {
      for (int l = 0; l < A.L; ++l) s_pref[l + 1] = s_pref[l] + A.lvl_count[n * A.L + l];
    } else {
      s_pref[1] = A.fixed_count;
    }
    s_nseg = 0; s_next = 0; s_nkeep = 0; s_mine = 0;
  }
  __syncthreads();
  const int L = A.lvl_count ? A.L : 1;
  const int nc = s_pref[L];
  const float4* gbox = A.boxes + (int64_t)n * A.slots_per_image;
  const float* gscore = A.scores + (int64_t)n * A.slots_per_image;
  const int64_t* gcls = A.classes ? A.classes + (int64_t)n * A.slots_per_image : nullptr;

  // slot of concatenation index i
  auto slot_of = [&](int i) -> int {
    int l = 0;
    while (l + 1 < L && i >= s_pref[l + 1]) ++l;
    return l * A.topk + (i - s_pref[l]);
  };

  // ---- 1. composite keys of this CTA's classes: class (18 bits) | inverted score (32) | concat index (14)
  //         ascending => class, score descending, index ascending
  for (int i0 = 0; i0 < nc; i0 += kNmsThreads) {
    const int i = i0 + tid;
    bool mine = false;
    uint64_t key = 0;
    if (i < nc) {
      const int s = slot_of(i);
      const int64_t craw = gcls ? gcls[s] : 0;
      const uint64_t c = (uint64_t)(craw & 0x3ffff);
      mine = ((int)(c % (uint64_t)S) == part);
      if (mine) {
        const uint32_t sb = __float_as_uint(gscore[s]);
        // order-preserving map for any float (negative scores can reach the stand-alone nms)
        const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
        key = (c << 46) | ((uint64_t)(0xffffffffu - ord) << 14) | (uint64_t)i;
      }
      if (A.exp_boxes && part == 0) {
        const int64_t eo = (int64_t)n * A.L * A.topk + i;
        A.exp_boxes[eo] = gbox[s];
        A.exp_scores[eo] = gscore[s];
        A.exp_classes[eo] = craw;
      }
    }
    const unsigned bm = __ballot_sync(kFull, mine);
    int base = 0;
    if (bm != 0u) {
      const int leader = __ffs(bm) - 1;
      if (lane == leader) base = atomicAdd(&s_mine, __popc(bm));
      base = __shfl_sync(kFull, base, leader);
      if (mine) keys[base + __popc(bm & ((1u << lane) - 1u))] = key;
    }
  }
  if (tid == 0 && part == 0 && A.exp_count) A.exp_count[n] = nc;
  __syncthreads();
  const int mc = s_mine;
  int m = 1;
  while (m < mc) m <<= 1;
  for (int i = mc + tid; i < m; i += kNmsThreads) keys[i] = ~0ull;
  __syncthreads();
  bitonic_asc<kNmsThreads>(keys, m);

  // ---- 2. boxes in sorted order, segment starts
  for (int i = tid; i < mc; i += kNmsThreads) {
    sbox[i] = gbox[slot_of((int)(keys[i] & 0x3fff))];
    dead[i] = 0;
  }
  __syncthreads();
  for (int i0 = 0; i0 < mc; i0 += kNmsThreads) {
    const int i = i0 + tid;
    const bool start = (i < mc) && (i == 0 || (keys[i] >> 46) != (keys[i - 1] >> 46));
    const unsigned bm = __ballot_sync(kFull, start);
    if (lane == 0) s_warp[wid] = __popc(bm);
    __syncthreads();
    int before = s_nseg;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    if (start) seg[before + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)i;
    int tot = 0;
    if (tid == 0)
      for (int w = 0; w < kNmsThreads / 32; ++w) tot += s_warp[w];
    __syncthreads();
    if (tid == 0) s_nseg += tot;
    __syncthreads();
  }
  const int nseg = s_nseg;

  // ---- 3a. big class segments (RPN levels: thousands of boxes in one class): the whole CTA works on one segment.
  //      Batches of 32 boxes in score order: warp 0 runs the greedy pass inside the batch, then every thread
  //      tests the boxes behind the batch against the batch's survivors.  Same result as the sequential greedy
  //      pass (a box is suppressed iff an earlier KEPT box overlaps it), two barriers per 32 boxes.
  for (int s = 0; s < nseg; ++s) {
    const int b = seg[s];
    const int e = (s + 1 < nseg) ? (int)seg[s + 1] : mc;
    if (e - b <= kNmsBigSeg) continue;   // uniform
    for (int i0 = b; i0 < e; i0 += 32) {
      const int i1 = min(i0 + 32, e);
      if (wid == 0) {
        for (int i = i0; i < i1; ++i) {
          if (dead[i]) continue;   // warp-uniform
          const float4 bi = sbox[i];
          const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
          const int j = i + 1 + lane;
          if (j < i1 && !dead[j] && nms_suppresses(bi, ai, sbox[j], A.thr)) dead[j] = 1;
          __syncwarp();
        }
        const bool alive = (i0 + lane < i1) && !dead[i0 + lane];
        const unsigned bm = __ballot_sync(kFull, alive);
        if (alive) s_batch[__popc(bm & ((1u << lane) - 1u))] = i0 + lane;
        if (lane == 0) s_nbatch = __popc(bm);
      }
      __syncthreads();
      const int nk = s_nbatch;
      if (nk > 0) {
        for (int j = i1 + tid; j < e; j += kNmsThreads) {
          if (dead[j]) continue;
          const float4 bj = sbox[j];
          for (int q = 0; q < nk; ++q) {
            const float4 bi = sbox[s_batch[q]];
            const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
            if (nms_suppresses(bi, ai, bj, A.thr)) { dead[j] = 1; break; }
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- 3b. greedy NMS, one warp per (small) class segment (torchvision nms_kernel semantics)
  for (;;) {
    int s = 0;
    if (lane == 0) s = atomicAdd(&s_next, 1);
    s = __shfl_sync(kFull, s, 0);
    if (s >= nseg) break;
    const int b = seg[s];
    const int e = (s + 1 < nseg) ? (int)seg[s + 1] : mc;
    if (e - b > kNmsBigSeg) continue;    // done above
    for (int i = b; i < e; ++i) {
      if (dead[i]) continue;   // warp-uniform (shared memory, synchronised below)
      const float4 bi = sbox[i];
      const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
      for (int j = i + 1 + lane; j < e; j += 32) {
        if (dead[j]) continue;
        if (nms_suppresses(bi, ai, sbox[j], A.thr)) dead[j] = 1;
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- 4. this CTA's survivors by score descending (ties: lower concat index first)
  uint64_t* k2 = reinterpret_cast<uint64_t*>(sbox);  // the sorted boxes are no longer needed
  //      (compacted first: only the survivors are sorted, not the whole padded candidate list)
  for (int i0 = 0; i0 < mc; i0 += kNmsThreads) {
    const int i = i0 + tid;
    const bool alive = (i < mc) && !dead[i];
    const unsigned bm = __ballot_sync(kFull, alive);
    if (bm != 0u) {
      int base = 0;
      const int leader = __ffs(bm) - 1;
      if (lane == leader) base = atomicAdd(&s_nkeep, __popc(bm));
      base = __shfl_sync(kFull, base, leader);
      if (alive) k2[base + __popc(bm & ((1u << lane) - 1u))] = keys[i] & ((1ull << 46) - 1ull);  // class field dropped
    }
  }
  __syncthreads();
  {
    const int nk0 = s_nkeep;
    int m4 = 1;
    while (m4 < nk0) m4 <<= 1;
    for (int i = nk0 + tid; i < m4; i += kNmsThreads) k2[i] = ~0ull;
    __syncthreads();
    bitonic_asc<kNmsThreads>(k2, m4);
  }
  int mine_keep = s_nkeep;
  if (mine_keep > A.part_cap) mine_keep = A.part_cap;
  uint64_t* pk = A.part_keys + ((int64_t)n * S + part) * A.part_cap;
  for (int t = tid; t < mine_keep; t += kNmsThreads) pk[t] = k2[t];
  __syncthreads();
  if (tid == 0) {
    A.part_cnt[n * S + part] = mine_keep;
    __threadfence();
    s_last = (atomicAdd(&A.done[n], 1u) == (unsigned)S - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- 5. last CTA of the image: merge the parts' survivors by score and emit
  if (tid == 0) { A.done[n] = 0u; s_mine = 0; }
  __syncthreads();
  for (int p = 0; p < S; ++p) {
    const int c = __ldcg(&A.part_cnt[n * S + p]);
    const uint64_t* src = A.part_keys + ((int64_t)n * S + p) * A.part_cap;
    const int base = s_mine;
    for (int t = tid; t < c; t += kNmsThreads) keys[base + t] = __ldcg(&src[t]);
    __syncthreads();
    if (tid == 0) s_mine = base + c;
    __syncthreads();
  }
  const int tot = s_mine;
  int m2 = 1;
  while (m2 < tot) m2 <<= 1;
  for (int i = tot + tid; i < m2; i += kNmsThreads) keys[i] = ~0ull;
  __syncthreads();
  if (S > 1) bitonic_asc<kNmsThreads>(keys, m2);
  int nk = tot;
  if (A.max_out > 0 && nk > A.max_out) nk = A.max_out;
  if (A.post && A.out_boxes) {
    // detector_postprocess (modeling/postprocessing.py:8-52) on the final detections: Boxes.scale, Boxes.clip,
    // drop boxes that became empty (Boxes.nonempty), stable compaction.  max_out <= kNmsThreads: one row per thread.
    const float4 pp = A.post[n];
    const int t = tid;
    bool ok = false;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int64_t cl = 0;
    int ci = -1;
    if (t < nk) {
      ci = (int)(keys[t] & 0x3fff);
      const int s = slot_of(ci);
      b = postprocess_box(gbox[s], pp);
      sc = gscore[s];
      cl = gcls ? gcls[s] : 0;
      ok = (__fsub_rn(b.z, b.x) > 0.f) && (__fsub_rn(b.w, b.y) > 0.f);
    }
    const unsigned bm = __ballot_sync(kFull, ok);
    if (lane == 0) s_warp[wid] = __popc(bm);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      const int c = s_warp[w];
      if (w < wid) before += c;
      total += c;
    }
    const int pos = before + __popc(bm & ((1u << lane) - 1u));
    const int64_t ob = (int64_t)n * A.max_out;
    if (ok) {
      A.out_boxes[ob + pos] = b;
      A.out_scores[ob + pos] = sc;
      A.out_classes[ob + pos] = cl;
      if (A.keep) A.keep[(int64_t)n * A.keep_stride + pos] = ci;
    }
    if (t >= total && t < A.max_out) {
      A.out_boxes[ob + t] = make_float4(0.f, 0.f, 0.f, 0.f);
      A.out_scores[ob + t] = 0.f;
      A.out_classes[ob + t] = 0;
      if (A.keep && t < A.keep_stride) A.keep[(int64_t)n * A.keep_stride + t] = -1;
    }
    if (tid == 0 && A.num_keep) A.num_keep[n] = total;
    return;
  }
  if (tid == 0 && A.num_keep) A.num_keep[n] = nk;
  const int out_rows = (A.max_out > 0) ? A.max_out : nk;
  for (int t = tid; t < out_rows; t += kNmsThreads) {
    if (t < nk) {
      const int ci = (int)(keys[t] & 0x3fff);
      const int s = slot_of(ci);
      if (A.keep) A.keep[(int64_t)n * A.keep_stride + t] = ci;
      if (A.out_boxes) {
        A.out_boxes[(int64_t)n * A.max_out + t] = gbox[s];
        A.out_scores[(int64_t)n * A.max_out + t] = gscore[s];
        if (A.out_classes) A.out_classes[(int64_t)n * A.max_out + t] = gcls ? gcls[s] : 0;
      }
    } else {
      if (A.keep && t < A.keep_stride) A.keep[(int64_t)n * A.keep_stride + t] = -1;
      if (A.out_boxes) {
        A.out_boxes[(int64_t)n * A.max_out + t] = make_float4(0.f, 0.f, 0.f, 0.f);
        A.out_scores[(int64_t)n * A.max_out + t] = 0.f;
        if (A.out_classes) A.out_classes[(int64_t)n * A.max_out + t] = 0;
      }
    }
  }
}

// stand-alone detector_postprocess on any (n,4) box list: scaled + clipped boxes and a keep flag per box
__global__ void __launch_bounds__(256) postprocess_boxes_kernel(const float4* __restrict__ boxes, int64_t n, float4 pp,
                                                                float4* __restrict__ out, uint8_t* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 b = postprocess_box(boxes[i], pp);
  out[i] = b;
  keep[i] = ((__fsub_rn(b.z, b.x) > 0.f) && (__fsub_rn(b.w, b.y) > 0.f)) ? 1 : 0;
}

static int nms_split_for(int N) {
  int s = 1;
  while (s * 2 * N <= 148 && s < 8) s <<= 1;   // fill the 148 SMs: one CTA per SM, up to 8 per image
  return s;
}
struct NmsWs {
  size_t off_done, off_cnt, off_keys, total;
};
static NmsWs nms_ws_layout(int N, int split, int part_cap) {
  NmsWs w;
  size_t o = 0;
  w.off_done = o; o += align_up(sizeof(unsigned) * (size_t)N, 16);
  w.off_cnt = o;  o += align_up(sizeof(int) * (size_t)N * split, 16);
  w.off_keys = o; o += align_up(sizeof(uint64_t) * (size_t)N * split * part_cap, 16);
  w.total = o;
  return w;
}

static float threshold_floor(double thr) {
  // fp32 IoU `ovr > (double)thr`  <=>  `ovr > f` with f the largest float <= thr
  float f = (float)thr;
  if ((double)f > thr) f = nextafterf(f, -INFINITY);
  return f;
}

constexpr size_t kNmsSmem = (size_t)kNmsCap * 27;
constexpr size_t kSelSmem = (size_t)kSelCap * 8;

static DetectLevels plan_levels(const int64_t* off, int num_levels, int K, int topk, int N) {
  DetectLevels lv;
  lv.num_levels = num_levels;
  int base = 0, maxp = 1;
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
  }
  lv.off[num_levels] = off[num_levels];
  for (int l = num_levels; l < kMaxLevels; ++l) { lv.nparts[l] = 0; lv.part_base[l] = base; lv.k[l] = 0; lv.part_len[l] = 0; }
  lv.total_parts = base;
  lv.max_parts = maxp;
  return lv;
}

struct DetectWs {
  size_t off_done, off_pcount, off_pkeys, off_lvl, off_cbox, off_cscore, off_ccls, off_nms, total;
};
static DetectWs detect_ws_layout(int N, int num_levels, int topk, int max_parts, int max_det = 0) {
  DetectWs w;
  size_t o = 0;
  const size_t slabs = (size_t)N * num_levels;
  w.off_done = o;   o += align_up(sizeof(unsigned) * slabs, 16);
  w.off_lvl = o;    o += align_up(sizeof(int) * slabs, 16);
  w.off_pcount = o; o += align_up(sizeof(int) * slabs * max_parts, 16);
  w.off_pkeys = o;  o += align_up(sizeof(uint64_t) * slabs * max_parts * topk, 16);
  w.off_cbox = o;   o += align_up(sizeof(float4) * slabs * topk, 16);
  w.off_cscore = o; o += align_up(sizeof(float) * slabs * topk, 16);
  w.off_ccls = o;   o += align_up(sizeof(int64_t) * slabs * topk, 16);
  w.off_nms = o;    o += nms_ws_layout(N, 8, max_det > 0 ? max_det : 1024).total;
  w.total = o;
  return w;
}


// ------------------------------------------------------------------------------------------
// RPN proposal selection (proposal_generator/rpn_outputs.py:52-151, steps 1-2 and the clip / min-size filter
// of step 3): one CTA per (level, image) row.  Exact top-k of the raw objectness logits by a radix select on
// 64-bit keys (inverted ordered score | index: unique, so "score descending, lower index first" is a total
// order), bitonic sort of the k winners in shared memory, gather of their proposal boxes, Boxes.clip to the
// image, Boxes.nonempty(threshold = min_box_side_len), stable compaction into the (N, L*topk) candidate slots
// that nms_image_kernel consumes (class id = level id, so its per-class NMS is the per-level batched_nms).
// ------------------------------------------------------------------------------------------
constexpr int kRpnThreads = 1024;
constexpr int kRpnBins = 2048;          // 11-bit digits
constexpr int kRpnMaxK = 8192;

struct RpnArgs {
  const float* logits[kMaxLevels];      // level l: (N, hwa[l])
  const float4* props[kMaxLevels];      // level l: (N, hwa[l], 4)
  int hwa[kMaxLevels];
  int k[kMaxLevels];                    // min(pre_nms_topk, hwa)
  int topk;                             // slot stride per level
  int num_levels;
  const float* image_sizes;             // device (N, 2): height, width
  float min_size;
  float4* cand_box;                     // (N, L*topk)
  float* cand_score;
  int64_t* cand_class;
  int* lvl_count;                       // (N, L)
};

__device__ __forceinline__ uint64_t rpn_key(float v, uint32_t idx) {
  const uint32_t sb = __float_as_uint(v);
  const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);
  return ((uint64_t)(0xffffffffu - ord) << 32) | (uint64_t)idx;   // ascending = score desc, index asc
}

__global__ void __launch_bounds__(kRpnThreads) rpn_select_kernel(const RpnArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);     // next_pow2(k) entries
  __shared__ unsigned hist[kRpnBins];
  __shared__ int s_warp[kRpnThreads / 32];
  __shared__ uint64_t s_prefix;
  __shared__ int s_remaining, s_done, s_count, s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int l = blockIdx.x, n = blockIdx.y;
  const int len = A.hwa[l], k = A.k[l];
  const float* row = A.logits[l] + (int64_t)n * len;
  if (k <= 0) {
    if (tid == 0) A.lvl_count[n * A.num_levels + l] = 0;
    return;
  }
  // ---- 1. radix select: the k smallest 64-bit keys are those whose top `bits` bits are <= prefix
  int bits = 0;           // number of leading bits fixed so far
  uint64_t prefix = 0;    // their value
  if (k < len) {
    if (tid == 0) { s_remaining = k; s_done = 0; }
    for (int shift = 64 - 11; ; shift -= 11) {
      const int sh = shift < 0 ? 0 : shift;
      const int width = shift < 0 ? 11 + shift : 11;   // the last digit is 64 - 5*11 = 9 bits
      for (int b = tid; b < kRpnBins; b += kRpnThreads) hist[b] = 0u;
      __syncthreads();
      for (int i0 = 0; i0 < len; i0 += kRpnThreads) {
        const int i = i0 + tid;
        bool in = false;
        unsigned digit = 0;
        if (i < len) {
          const uint64_t key = rpn_key(row[i], (uint32_t)i);
          in = (bits == 0) || ((key >> (64 - bits)) == prefix);
          digit = (unsigned)((key >> sh) & ((1u << width) - 1u));
        }
        // warp-aggregated histogram update (objectness logits crowd a few exponent bins)
        const unsigned act = __ballot_sync(kFull, in);
        if (in) {
          const unsigned peers = __match_any_sync(act, digit);
          if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
        }
      }
      __syncthreads();
      // ascending scan over the bins: thread t owns bins 2t, 2t+1
      const unsigned h0 = hist[2 * tid], h1 = hist[2 * tid + 1];
      int c = (int)(h0 + h1), incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[wid] = incl;
      __syncthreads();
      int before = 0;
      for (int w = 0; w < wid; ++w) before += s_warp[w];
      incl += before;
      const int excl = incl - c;
      const int rem = s_remaining;
      __syncthreads();
      if (excl < rem && rem <= incl) {   // exactly one thread
        int bin = 2 * tid, below = excl, cnt = (int)h0;
        if (rem > excl + (int)h0) { bin = 2 * tid + 1; below = excl + (int)h0; cnt = (int)h1; }
        s_prefix = (prefix << width) | (uint64_t)bin;
        s_remaining = rem - below;
        s_done = (cnt == rem - below) ? 1 : 0;   // the whole bin is taken: no need to look at lower bits
      }
      __syncthreads();
      prefix = s_prefix;
      bits += width;
      if (s_done || bits >= 64) break;
    }
  }
  // ---- 2. collect the k winners (unordered), sort them
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int i0 = 0; i0 < len; i0 += kRpnThreads) {
    const int i = i0 + tid;
    bool take = false;
    uint64_t key = 0;
    if (i < len) {
      key = rpn_key(row[i], (uint32_t)i);
      take = (bits == 0) || ((key >> (64 - bits)) <= prefix);
    }
    const unsigned bm = __ballot_sync(kFull, take);
    if (bm) {
      int base = 0;
      const int leader = __ffs(bm) - 1;
      if (lane == leader) base = atomicAdd(&s_count, __popc(bm));
      base = __shfl_sync(kFull, base, leader);
      if (take) keys[base + __popc(bm & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const int got = s_count;   // == k
  int m = 1;
  while (m < got) m <<= 1;
  for (int i = got + tid; i < m; i += kRpnThreads) keys[i] = ~0ull;
  __syncthreads();
  bitonic_asc<kRpnThreads>(keys, m);

  // ---- 3. gather, clip, min-size filter, stable compaction into the candidate slots
  const float ih = A.image_sizes[2 * n], iw = A.image_sizes[2 * n + 1];
  const float4* prow = A.props[l] + (int64_t)n * len;
  const int64_t slot0 = ((int64_t)n * A.num_levels + l) * A.topk;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < got; t0 += kRpnThreads) {
    const int t = t0 + tid;
    bool ok = false;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t idx = 0;
    if (t < got) {
      idx = (uint32_t)keys[t];
      b = prow[idx];
      b.x = fminf(fmaxf(b.x, 0.f), iw); b.y = fminf(fmaxf(b.y, 0.f), ih);   // Boxes.clip (boxes.py:122-136)
      b.z = fminf(fmaxf(b.z, 0.f), iw); b.w = fminf(fmaxf(b.w, 0.f), ih);
      ok = (__fsub_rn(b.z, b.x) > A.min_size) && (__fsub_rn(b.w, b.y) > A.min_size);   // nonempty (:138-151)
    }
    const unsigned bm = __ballot_sync(kFull, ok);
    if (lane == 0) s_warp[wid] = __popc(bm);
    __syncthreads();
    int before = s_base, total = 0;
    for (int w = 0; w < kRpnThreads / 32; ++w) {
      const int c = s_warp[w];
      if (w < wid) before += c;
      total += c;
    }
    if (ok) {
      const int64_t o = slot0 + before + __popc(bm & ((1u << lane) - 1u));
      A.cand_box[o] = b;
      A.cand_score[o] = row[idx];
      A.cand_class[o] = l;
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  if (tid == 0) A.lvl_count[n * A.num_levels + l] = s_base;
}

struct RpnWs {
  size_t off_lvl, off_cbox, off_cscore, off_ccls, off_nms, total;
  int split;
};
static RpnWs rpn_ws_layout(int N, int num_levels, int topk, int post, int split) {
  RpnWs w;
  size_t o = 0;
  const size_t slabs = (size_t)N * num_levels;
  w.off_lvl = o;    o += align_up(sizeof(int) * slabs, 16);
  w.off_cbox = o;   o += align_up(sizeof(float4) * slabs * topk, 16);
  w.off_cscore = o; o += align_up(sizeof(float) * slabs * topk, 16);
  w.off_ccls = o;   o += align_up(sizeof(int64_t) * slabs * topk, 16);
  w.off_nms = o;    o += nms_ws_layout(N, split, post).total;
  w.total = o;
  w.split = split;
  return w;
}
// CTAs per image for the per-level NMS: enough to fill the SMs, and enough that no CTA's levels (l % split)
// hold more than kNmsCap candidates; the merge of split * post survivors must fit the same buffer.
static int rpn_split_for(int N, const int* k, int num_levels, int post) {
  int smax = 1;
  while (smax * 2 <= 8 && (int64_t)smax * 2 * post <= kNmsCap) smax <<= 1;
  int smin = 0;
  for (int s = 1; s <= 8; s <<= 1) {
    int worst = 0;
    for (int p = 0; p < s; ++p) {
      int sum = 0;
      for (int l = p; l < num_levels; l += s) sum += k[l];
      if (sum > worst) worst = sum;
    }
    if (worst <= kNmsCap) { smin = s; break; }
  }
  if (smin == 0 || smin > smax) return 0;
  int s = nms_split_for(N);
  if (s < smin) s = smin;
  if (s > smax) s = smax;
  return s;
}

// ------------------------------------------------------------------------------------------
// fast_rcnn_inference_single_image, candidate stage (roi_heads/fast_rcnn.py:76-105): drop the background
// column, Boxes.clip to the image, keep (r, k) with score > thresh in row-major order (= torch.nonzero order,
// which fixes the tie-break of the NMS that follows).  One CTA walks the R x K score matrix in order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) score_filter_kernel(const float4* __restrict__ boxes, int C,
                                                            const float* __restrict__ scores, int64_t R, int K,
                                                            float img_h, float img_w, float thresh,
                                                            float4* __restrict__ out_boxes,
                                                            float* __restrict__ out_scores,
                                                            int64_t* __restrict__ out_classes,
                                                            int64_t* __restrict__ out_rows,
                                                            int32_t* __restrict__ out_count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int64_t E = R * K;
  for (int64_t e0 = 0; e0 < E; e0 += 1024) {
    const int64_t e = e0 + tid;
    bool ok = false;
    float sc = 0.f;
    int64_t r = 0;
    int k = 0;
    if (e < E) {
      r = e / K;
      k = (int)(e - r * K);
      sc = scores[r * (K + 1) + k];
      ok = sc > thresh;
    }
    const unsigned bm = __ballot_sync(kFull, ok);
    if (lane == 0) s_warp[wid] = __popc(bm);
    __syncthreads();
    int before = s_base, total = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = s_warp[w];
      if (w < wid) before += c;
      total += c;
    }
    if (ok) {
      const int o = before + __popc(bm & ((1u << lane) - 1u));
      float4 b = boxes[r * C + (C == 1 ? 0 : k)];
      b.x = fminf(fmaxf(b.x, 0.f), img_w); b.y = fminf(fmaxf(b.y, 0.f), img_h);
      b.z = fminf(fmaxf(b.z, 0.f), img_w); b.w = fminf(fmaxf(b.w, 0.f), img_h);
      out_boxes[o] = b;
      out_scores[o] = sc;
      out_classes[o] = k;
      out_rows[o] = r;
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  if (tid == 0) *out_count = s_base;
}

}  // namespace fsg

using namespace fsg;

extern "C" size_t fsg_nms_workspace_bytes(int64_t n) {
  if (n <= 0) return 16;
  if (n > kNmsCap) return n > kNmsLargeMax ? 0 : nms_large_ws_layout(n).total;
  return nms_ws_layout(1, nms_split_for(1), (int)(n < kNmsCap ? n : kNmsCap)).total;
}

extern "C" int fsg_nms(const float* boxes, const float* scores, const int64_t* class_ids, int64_t n,
                       double iou_threshold, int64_t* keep, int32_t* num_keep, void* workspace,
                       size_t workspace_bytes, fsg_stream_t stream) {
  if (n < 0 || !num_keep) return FSG_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    FSG_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), s));
    return FSG_OK;
  }
  if (!boxes || !scores || !keep) return FSG_ERR_INVALID_ARG;
  if (n > kNmsCap)   // more boxes than one CTA's shared memory holds: rank / bit-matrix / sweep kernels
    return nms_large(boxes, scores, class_ids, n, threshold_floor(iou_threshold), keep, num_keep, workspace,
                     workspace_bytes, s);
  const int split = class_ids ? nms_split_for(1) : 1;
  const NmsWs w = nms_ws_layout(1, split, (int)n);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  FSG_CUDA_TRY(cudaMemsetAsync(ws + w.off_done, 0, w.off_cnt - w.off_done, s));
  NmsArgs a = {};
  a.boxes = (const float4*)boxes; a.scores = scores; a.classes = class_ids;
  a.slots_per_image = n; a.lvl_count = nullptr; a.L = 1; a.topk = (int)n; a.fixed_count = (int)n;
  a.thr = threshold_floor(iou_threshold); a.max_out = 0;
  a.split = split; a.part_cap = (int)n;
  a.part_keys = (uint64_t*)(ws + w.off_keys); a.part_cnt = (int*)(ws + w.off_cnt); a.done = (unsigned*)(ws + w.off_done);
  a.keep = keep; a.keep_stride = n; a.num_keep = num_keep;
  FSG_CUDA_TRY(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmem));
  nms_image_kernel<<<dim3((unsigned)split, 1), kNmsThreads, kNmsSmem, s>>>(a);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" size_t fsg_detect_workspace_bytes(int N, int64_t R, int K, int num_levels, int topk) {
  if (N <= 0 || num_levels <= 0 || num_levels > kMaxLevels || topk <= 0 || K <= 0) return 0;
  (void)R;
  int maxp = kSelCap / topk;
  if (maxp > 16) maxp = 16;
  if (maxp < 1) maxp = 1;
  return detect_ws_layout(N, num_levels, topk, maxp).total;
}

extern "C" int fsg_detect(const float* logits, const float* deltas, const float* anchors,
                          int64_t anchor_image_stride, int N, int64_t R, int K, const int64_t* h_level_offsets,
                          int num_levels, float score_threshold, int topk, double nms_threshold, int max_det,
                          const float* h_box_weights, float scale_clamp, float* out_boxes, float* out_scores,
                          int64_t* out_classes, int32_t* out_count, float* cand_boxes, float* cand_scores,
                          int64_t* cand_classes, int32_t* cand_count, int64_t* keep_idx, const float* postprocess,
                          void* workspace, size_t workspace_bytes, fsg_stream_t stream) {
  if (postprocess && ((uintptr_t)postprocess & 15)) return FSG_ERR_INVALID_ARG;
  if (N <= 0 || R <= 0 || K <= 0 || !h_level_offsets || num_levels <= 0 || num_levels > kMaxLevels)
    return FSG_ERR_INVALID_ARG;
  if (!logits || !deltas || !anchors || !out_boxes || !out_scores || !out_classes || !out_count || !h_box_weights)
    return FSG_ERR_INVALID_ARG;
  if (topk <= 0 || max_det <= 0) return FSG_ERR_INVALID_ARG;
  if (max_det > 1024) return FSG_ERR_UNSUPPORTED;
  if (anchor_image_stride % 4 != 0) return FSG_ERR_INVALID_ARG;
  if (h_level_offsets[0] != 0 || h_level_offsets[num_levels] != R) return FSG_ERR_INVALID_ARG;
  if (topk > kSelTrigger || (int64_t)num_levels * topk > kNmsCap || K > 65535 || N > 65535)
    return FSG_ERR_UNSUPPORTED;
  for (int l = 0; l < num_levels; ++l) {
    const int64_t hwa = h_level_offsets[l + 1] - h_level_offsets[l];
    if (hwa < 0 || hwa * K >= ((int64_t)1 << 32)) return FSG_ERR_UNSUPPORTED;
  }
  if ((cand_boxes || cand_scores || cand_classes) && !(cand_boxes && cand_scores && cand_classes))
    return FSG_ERR_INVALID_ARG;
  const DetectLevels lv = plan_levels(h_level_offsets, num_levels, K, topk, N);
  int maxp = kSelCap / topk;
  if (maxp > 16) maxp = 16;
  if (maxp < 1) maxp = 1;
  DetectLevels lv2 = lv;
  lv2.max_parts = maxp;
  const DetectWs w = detect_ws_layout(N, num_levels, topk, maxp);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  FSG_CUDA_TRY(cudaMemsetAsync(ws + w.off_done, 0, w.off_pcount - w.off_done, s));  // done + lvl_count

  SelectArgs sa;
  sa.logits = logits; sa.deltas = (const float4*)deltas; sa.anchors = (const float4*)anchors;
  sa.anchor_stride4 = anchor_image_stride / 4; sa.R = R; sa.K = K; sa.topk = topk;
  sa.thr = score_threshold;
  {
    // logit of the score threshold, minus a safety margin (the exact `score > thr` test follows)
    const double t = (double)score_threshold;
    double lg = (t <= 0.0) ? -INFINITY : ((t >= 1.0) ? INFINITY : log(t / (1.0 - t)));
    sa.xpre = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
    if (t <= 0.0) sa.xpre = -INFINITY;
  }
  sa.wx = h_box_weights[0]; sa.wy = h_box_weights[1]; sa.ww = h_box_weights[2]; sa.wh = h_box_weights[3];
  sa.clampv = scale_clamp;
  sa.part_keys = (uint64_t*)(ws + w.off_pkeys); sa.part_count = (int*)(ws + w.off_pcount);
  sa.done = (unsigned*)(ws + w.off_done);
  sa.cand_box = (float4*)(ws + w.off_cbox); sa.cand_score = (float*)(ws + w.off_cscore);
  sa.cand_class = (int64_t*)(ws + w.off_ccls); sa.lvl_count = (int*)(ws + w.off_lvl);
  FSG_CUDA_TRY(cudaFuncSetAttribute(detect_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelSmem));
  dim3 grid((unsigned)lv2.total_parts, (unsigned)N);
  detect_select_kernel<<<grid, kSelThreads, kSelSmem, s>>>(sa, lv2);
  FSG_LAUNCH_CHECK();

  NmsArgs a = {};
  a.boxes = sa.cand_box; a.scores = sa.cand_score; a.classes = sa.cand_class;
  a.slots_per_image = (int64_t)num_levels * topk; a.lvl_count = sa.lvl_count; a.L = num_levels; a.topk = topk;
  a.fixed_count = 0; a.thr = threshold_floor(nms_threshold); a.max_out = max_det;
  {
    const int split = nms_split_for(N);
    const NmsWs nw = nms_ws_layout(N, split, max_det);
    char* nws = ws + w.off_nms;
    FSG_CUDA_TRY(cudaMemsetAsync(nws + nw.off_done, 0, nw.off_cnt - nw.off_done, s));
    a.split = split; a.part_cap = max_det;
    a.part_keys = (uint64_t*)(nws + nw.off_keys); a.part_cnt = (int*)(nws + nw.off_cnt);
    a.done = (unsigned*)(nws + nw.off_done);
  }
  a.keep = keep_idx; a.keep_stride = max_det; a.num_keep = out_count;
  a.out_boxes = (float4*)out_boxes; a.out_scores = out_scores; a.out_classes = out_classes;
  a.post = (const float4*)postprocess;
  a.exp_boxes = (float4*)cand_boxes; a.exp_scores = cand_scores; a.exp_classes = cand_classes;
  a.exp_count = cand_count;
  FSG_CUDA_TRY(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmem));
  nms_image_kernel<<<dim3((unsigned)a.split, (unsigned)N), kNmsThreads, kNmsSmem, s>>>(a);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" int fsg_postprocess_boxes(const float* boxes, int64_t n, float scale_x, float scale_y, float clip_w,
                                     float clip_h, float* out_boxes, uint8_t* keep, fsg_stream_t stream) {
  if (n < 0) return FSG_ERR_INVALID_ARG;
  if (n == 0) return FSG_OK;
  if (!boxes || !out_boxes || !keep) return FSG_ERR_INVALID_ARG;
  if (((uintptr_t)boxes | (uintptr_t)out_boxes) & 15) return FSG_ERR_INVALID_ARG;
  postprocess_boxes_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)boxes, n, make_float4(scale_x, scale_y, clip_w, clip_h), (float4*)out_boxes, keep);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

static int rpn_plan(const int64_t* h_level_sizes, int num_levels, int pre_nms_topk, int post_nms_topk, int N,
                    int* k, int* topk_out, int* split_out) {
  if (!h_level_sizes || num_levels <= 0 || num_levels > kMaxLevels || pre_nms_topk <= 0 || post_nms_topk <= 0 || N <= 0)
    return FSG_ERR_INVALID_ARG;
  int topk = 0;
  int64_t total = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (h_level_sizes[l] < 0 || h_level_sizes[l] >= ((int64_t)1 << 31)) return FSG_ERR_UNSUPPORTED;
    k[l] = (int)(h_level_sizes[l] < pre_nms_topk ? h_level_sizes[l] : pre_nms_topk);
    if (k[l] > topk) topk = k[l];
    total += k[l];
  }
  if (topk > kRpnMaxK || total >= (1 << 14) || post_nms_topk > kNmsCap || N > 65535) return FSG_ERR_UNSUPPORTED;
  if (topk < 1) topk = 1;
  const int split = rpn_split_for(N, k, num_levels, post_nms_topk);
  if (split == 0) return FSG_ERR_UNSUPPORTED;
  *topk_out = topk;
  *split_out = split;
  return FSG_OK;
}

extern "C" size_t fsg_rpn_proposals_workspace_bytes(int N, const int64_t* h_level_sizes, int num_levels,
                                                    int pre_nms_topk, int post_nms_topk) {
  int k[kMaxLevels], topk, split;
  if (rpn_plan(h_level_sizes, num_levels, pre_nms_topk, post_nms_topk, N, k, &topk, &split) != FSG_OK) return 0;
  return rpn_ws_layout(N, num_levels, topk, post_nms_topk, split).total;
}

extern "C" int fsg_rpn_proposals(const float* const* h_level_proposals, const float* const* h_level_logits,
                                 const int64_t* h_level_sizes, int num_levels, int N, const float* image_sizes,
                                 int pre_nms_topk, int post_nms_topk, double nms_threshold, float min_box_side_len,
                                 float* out_boxes, float* out_logits, int64_t* out_levels, int32_t* out_count,
                                 void* workspace, size_t workspace_bytes, fsg_stream_t stream) {
  int k[kMaxLevels], topk, split;
  const int st = rpn_plan(h_level_sizes, num_levels, pre_nms_topk, post_nms_topk, N, k, &topk, &split);
  if (st != FSG_OK) return st;
  if (!h_level_proposals || !h_level_logits || !image_sizes || !out_boxes || !out_logits || !out_count)
    return FSG_ERR_INVALID_ARG;
  if (((uintptr_t)out_boxes) & 15) return FSG_ERR_INVALID_ARG;
  const RpnWs w = rpn_ws_layout(N, num_levels, topk, post_nms_topk, split);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  cudaStream_t s = (cudaStream_t)stream;

  RpnArgs ra = {};
  for (int l = 0; l < num_levels; ++l) {
    if (h_level_sizes[l] > 0 && (!h_level_proposals[l] || !h_level_logits[l])) return FSG_ERR_INVALID_ARG;
    if (((uintptr_t)h_level_proposals[l]) & 15) return FSG_ERR_INVALID_ARG;
    ra.logits[l] = h_level_logits[l]; ra.props[l] = (const float4*)h_level_proposals[l];
    ra.hwa[l] = (int)h_level_sizes[l]; ra.k[l] = k[l];
  }
  ra.topk = topk; ra.num_levels = num_levels; ra.image_sizes = image_sizes; ra.min_size = min_box_side_len;
  ra.cand_box = (float4*)(ws + w.off_cbox); ra.cand_score = (float*)(ws + w.off_cscore);
  ra.cand_class = (int64_t*)(ws + w.off_ccls); ra.lvl_count = (int*)(ws + w.off_lvl);
  int m = 1;
  while (m < topk) m <<= 1;
  const size_t smem = sizeof(uint64_t) * (size_t)m;
  FSG_CUDA_TRY(cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rpn_select_kernel<<<dim3((unsigned)num_levels, (unsigned)N), kRpnThreads, smem, s>>>(ra);
  FSG_LAUNCH_CHECK();

  NmsArgs a = {};
  a.boxes = ra.cand_box; a.scores = ra.cand_score; a.classes = ra.cand_class;
  a.slots_per_image = (int64_t)num_levels * topk; a.lvl_count = ra.lvl_count; a.L = num_levels; a.topk = topk;
  a.fixed_count = 0; a.thr = threshold_floor(nms_threshold); a.max_out = post_nms_topk;
  const NmsWs nw = nms_ws_layout(N, split, post_nms_topk);
  char* nws = ws + w.off_nms;
  FSG_CUDA_TRY(cudaMemsetAsync(nws + nw.off_done, 0, nw.off_cnt - nw.off_done, s));
  a.split = split; a.part_cap = post_nms_topk;
  a.part_keys = (uint64_t*)(nws + nw.off_keys); a.part_cnt = (int*)(nws + nw.off_cnt);
  a.done = (unsigned*)(nws + nw.off_done);
  a.keep = nullptr; a.keep_stride = post_nms_topk; a.num_keep = out_count;
  a.out_boxes = (float4*)out_boxes; a.out_scores = out_logits; a.out_classes = out_levels;
  FSG_CUDA_TRY(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmem));
  nms_image_kernel<<<dim3((unsigned)split, (unsigned)N), kNmsThreads, kNmsSmem, s>>>(a);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

extern "C" int fsg_score_filter(const float* boxes, int num_bbox_reg_classes, const float* scores, int64_t R, int K,
                                float image_height, float image_width, float score_thresh, float* out_boxes,
                                float* out_scores, int64_t* out_classes, int64_t* out_rows, int32_t* out_count,
                                fsg_stream_t stream) {
  if (R < 0 || K <= 0 || !out_count) return FSG_ERR_INVALID_ARG;
  if (num_bbox_reg_classes != 1 && num_bbox_reg_classes != K) return FSG_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (R == 0) {
    FSG_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
    return FSG_OK;
  }
  if (!boxes || !scores || !out_boxes || !out_scores || !out_classes || !out_rows) return FSG_ERR_INVALID_ARG;
  if (((uintptr_t)boxes | (uintptr_t)out_boxes) & 15) return FSG_ERR_INVALID_ARG;
  if (R * K >= ((int64_t)1 << 31)) return FSG_ERR_UNSUPPORTED;
  score_filter_kernel<<<1, 1024, 0, s>>>((const float4*)boxes, num_bbox_reg_classes, scores, R, K, image_height,
                                         image_width, score_thresh, (float4*)out_boxes, out_scores, out_classes,
                                         out_rows, out_count);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
