// Per-image / per-call NMS in shared memory (torchvision greedy semantics, per-class un-offset form) with the optional
// detector_postprocess epilogue; nms / batched_nms entry points (n <= 8192 here, larger n in nms_large.cu).
//
// Reference: detectron2/layers/nms.py:6,9-26; detectron2/modeling/postprocessing.py:8-52.
#include <stdlib.h>

#include "nms_kernel.cuh"
#include "nms_large.cuh"
#include "sort_utils.cuh"

namespace fsg {

__device__ __forceinline__ float4 postprocess_box(float4 b, float4 pp) {
  b.x = fminf(fmaxf(__fmul_rn(b.x, pp.x), 0.f), pp.z);
  b.y = fminf(fmaxf(__fmul_rn(b.y, pp.y), 0.f), pp.w);
  b.z = fminf(fmaxf(__fmul_rn(b.z, pp.x), 0.f), pp.z);
  b.w = fminf(fmaxf(__fmul_rn(b.w, pp.y), 0.f), pp.w);
  return b;
}

// Final detections of image n: the nk best survivors in score order, `ci_of(t)` = concatenation index of the t-th.
// Optional detector_postprocess (modeling/postprocessing.py:8-52): Boxes.scale, Boxes.clip, drop boxes that became
// empty (Boxes.nonempty), stable compaction.  Called by all kNmsThreads threads of the image's last CTA.
template <typename CiOf>
__device__ __forceinline__ void nms_emit(const NmsArgs& A, int n, int nk, bool bad, int L, const int* s_pref,
                                         int* s_warp, CiOf ci_of) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float4* gbox = A.boxes + (int64_t)n * A.slots_per_image;
  const float* gscore = A.scores + (int64_t)n * A.slots_per_image;
  const int64_t* gcls = A.classes ? A.classes + (int64_t)n * A.slots_per_image : nullptr;
  auto slot_of = [&](int i) -> int {
    int l = 0;
    while (l + 1 < L && i >= s_pref[l + 1]) ++l;
    return l * A.topk + (i - s_pref[l]);
  };
  if (A.post && A.out_boxes) {
    // max_out <= kNmsThreads: one row per thread
    const float4 pp = A.post[n];
    const int t = tid;
    bool ok = false;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int64_t cl = 0;
    int ci = -1;
    if (t < nk) {
      ci = ci_of(t);
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
    if (tid == 0 && A.num_keep) A.num_keep[n] = bad ? -1 : total;
    return;
  }
  if (tid == 0 && A.num_keep) A.num_keep[n] = bad ? -1 : nk;   // -1: a class id outside [0, 2^18)
  const int out_rows = (A.max_out > 0) ? A.max_out : nk;
  for (int t = tid; t < out_rows; t += kNmsThreads) {
    if (t < nk) {
      const int ci = ci_of(t);
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

// Greedy suppression over the class segments of `mc` boxes sorted by (class, score descending) in shared memory
// (torchvision nms_kernel semantics).  Big segments: the whole CTA, batches of 32 boxes in score order -- warp 0 runs
// the greedy pass inside the batch, then every thread tests the boxes behind the batch against the batch's survivors
// (same result as the sequential pass, two barriers per 32 boxes).  Small segments: one warp each.
struct NmsSuppressShared {
  int next, nbatch;
  int batch[32];
};
__device__ __forceinline__ void nms_suppress(const float4* sbox, unsigned char* dead, const uint16_t* seg, int nseg,
                                             int mc, float thr, NmsSuppressShared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
          if (j < i1 && !dead[j] && nms_suppresses(bi, ai, sbox[j], thr)) dead[j] = 1;
          __syncwarp();
        }
        const bool alive = (i0 + lane < i1) && !dead[i0 + lane];
        const unsigned bm = __ballot_sync(kFull, alive);
        if (alive) sh.batch[__popc(bm & ((1u << lane) - 1u))] = i0 + lane;
        if (lane == 0) sh.nbatch = __popc(bm);
      }
      __syncthreads();
      const int nk = sh.nbatch;
      if (nk > 0) {
        for (int j = i1 + tid; j < e; j += kNmsThreads) {
          if (dead[j]) continue;
          const float4 bj = sbox[j];
          for (int q = 0; q < nk; ++q) {
            const float4 bi = sbox[sh.batch[q]];
            const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
            if (nms_suppresses(bi, ai, bj, thr)) { dead[j] = 1; break; }
          }
        }
      }
      __syncthreads();
    }
  }
  for (;;) {
    int s = 0;
    if (lane == 0) s = atomicAdd(&sh.next, 1);
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
        if (nms_suppresses(bi, ai, sbox[j], thr)) dead[j] = 1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_image_kernel(const NmsArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);                      // kNmsCap * 8
  float4* sbox = reinterpret_cast<float4*>(smem_raw + (size_t)kNmsCap * 8);     // kNmsCap * 16
  uint16_t* seg = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kNmsCap * 24); // kNmsCap * 2
  unsigned char* dead = smem_raw + (size_t)kNmsCap * 26;                        // kNmsCap
  __shared__ int s_pref[kMaxLevels + 1];
  __shared__ int s_warp[kNmsThreads / 32];
  __shared__ int s_nseg, s_next, s_nkeep, s_mine, s_nbatch, s_bad;
  __shared__ int s_batch[32];
  __shared__ bool s_last;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int part = blockIdx.x;
  const int n = blockIdx.y;
  const int S = A.split;
  grid_dependency_sync();   // (no-op unless launched under programmatic dependent launch behind the select stage)
  if (tid == 0) {
    s_pref[0] = 0;
    if (A.lvl_count) {
      for (int l = 0; l < A.L; ++l) s_pref[l + 1] = s_pref[l] + A.lvl_count[n * A.L + l];
    } else {
      s_pref[1] = A.fixed_count;
    }
    s_nseg = 0; s_next = 0; s_nkeep = 0; s_mine = 0; s_bad = 0;
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
      if (craw < 0 || craw > 0x3ffff) s_bad = 1;   // the composite key holds 18 bits of class id (every CTA sees every
      const uint64_t c = (uint64_t)(craw & 0x3ffff);  // candidate, so the image's last CTA knows too): reported below
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
  nms_emit(A, n, nk, s_bad != 0, L, s_pref, s_warp, [&](int t) { return (int)(keys[t] & 0x3fff); });
}

// ------------------------------------------------------------------------------------------
// The same per-class NMS for candidates that arrive as L runs (the FPN levels), each already sorted by
// (score descending, position ascending) -- what the select stage of fsg_detect and the RPN select emit.
// The order by score of the concatenation is then a merge, not a sort: a candidate's rank is its position in its
// own run plus, for every other run, the number of entries that go before it (two binary searches per run; equal
// scores: the lower level first, i.e. the lower concatenation index -- the reference's stable sort).  With the
// merged rank in hand
//   * the (class, score) order needs only a 32-bit key (class << 13 | rank): half the shuffle traffic of the
//     64-bit composite key;
//   * the survivors need no second sort and the split CTAs no merge: each sets the bits of its survivors' ranks in
//     a per-image bitmap, and the image's last CTA turns the bitmap into positions with one prefix popcount.
// ------------------------------------------------------------------------------------------
constexpr size_t kRunsSmem = (size_t)kNmsCap * (4 + 2 + 4 + 16);   // scores|seg+dead, rank->cand, keys, boxes

__global__ void __launch_bounds__(kNmsThreads, 1) nms_runs_kernel(const NmsArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sbox = reinterpret_cast<float4*>(smem_raw);                                   // kNmsCap * 16
  uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kNmsCap * 16);        // kNmsCap * 4
  float* s_score = reinterpret_cast<float*>(smem_raw + (size_t)kNmsCap * 20);           // kNmsCap * 4 (phase A/B)
  uint16_t* seg = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kNmsCap * 20);         //   reused: kNmsCap * 2
  unsigned char* dead = smem_raw + (size_t)kNmsCap * 22;                                //   reused: kNmsCap
  uint16_t* cand_of = reinterpret_cast<uint16_t*>(smem_raw + (size_t)kNmsCap * 24);     // kNmsCap * 2
  __shared__ int s_pref[kMaxLevels + 1];
  __shared__ int s_warp[kNmsThreads / 32];
  __shared__ int s_nseg, s_mine, s_bad;
  __shared__ NmsSuppressShared s_sup;
  __shared__ bool s_last;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int part = blockIdx.x;
  const int n = blockIdx.y;
  const int S = A.split;
  grid_dependency_sync();   // (no-op unless launched under programmatic dependent launch behind the select stage)
  if (tid == 0) {
    s_pref[0] = 0;
    for (int l = 0; l < A.L; ++l) s_pref[l + 1] = s_pref[l] + A.lvl_count[n * A.L + l];
    s_nseg = 0; s_mine = 0; s_bad = 0; s_sup.next = 0;
  }
  __syncthreads();
  const int L = A.L;
  const int nc = s_pref[L];
  const float4* gbox = A.boxes + (int64_t)n * A.slots_per_image;
  const float* gscore = A.scores + (int64_t)n * A.slots_per_image;
  const int64_t* gcls = A.classes + (int64_t)n * A.slots_per_image;
  auto slot_of = [&](int i) -> int {
    int l = 0;
    while (l + 1 < L && i >= s_pref[l + 1]) ++l;
    return l * A.topk + (i - s_pref[l]);
  };

  // ---- A. every candidate's score (all runs are needed for the merge ranks), the optional candidate export
  for (int i = tid; i < nc; i += kNmsThreads) {
    const int s = slot_of(i);
    s_score[i] = gscore[s];
    if (A.exp_boxes && part == 0) {
      const int64_t eo = (int64_t)n * A.L * A.topk + i;
      A.exp_boxes[eo] = gbox[s];
      A.exp_scores[eo] = gscore[s];
      A.exp_classes[eo] = gcls[s];
    }
  }
  if (tid == 0 && part == 0 && A.exp_count) A.exp_count[n] = nc;
  __syncthreads();

  // ---- B. merged rank of this CTA's candidates; key = class << 13 | rank
  uint16_t* g_r2c = A.rank2cand + (int64_t)n * kNmsCap;
  for (int i0 = 0; i0 < nc; i0 += kNmsThreads) {
    const int i = i0 + tid;
    bool mine = false;
    uint32_t key = 0u;
    if (i < nc) {
      int l0 = 0;
      while (l0 + 1 < L && i >= s_pref[l0 + 1]) ++l0;
      const int64_t craw = gcls[l0 * A.topk + (i - s_pref[l0])];
      if (craw < 0 || craw > 0x3ffff) s_bad = 1;
      const uint32_t c = (uint32_t)(craw & 0x3ffff);
      mine = ((int)(c % (uint32_t)S) == part);
      if (mine) {
        const float sc = s_score[i];
        int rank = i - s_pref[l0];
        for (int l = 0; l < L; ++l) {
          if (l == l0) continue;
          // entries of run l that go before candidate i: score greater, or equal and the run is an earlier level
          const float* run = s_score + s_pref[l];
          int lo = 0, hi = s_pref[l + 1] - s_pref[l];
          if (l < l0) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (run[mid] >= sc) lo = mid + 1; else hi = mid; } }
          else        { while (lo < hi) { const int mid = (lo + hi) >> 1; if (run[mid] > sc) lo = mid + 1; else hi = mid; } }
          rank += lo;
        }
        key = (c << 13) | (uint32_t)rank;
        cand_of[rank] = (uint16_t)i;
        g_r2c[rank] = (uint16_t)i;
      }
    }
    const unsigned bm = __ballot_sync(kFull, mine);
    if (bm != 0u) {
      int base = 0;
      const int leader = __ffs(bm) - 1;
      if (lane == leader) base = atomicAdd(&s_mine, __popc(bm));
      base = __shfl_sync(kFull, base, leader);
      if (mine) keys[base + __popc(bm & ((1u << lane) - 1u))] = key;
    }
  }
  __syncthreads();
  const int mc = s_mine;
  int m = 1;
  while (m < mc) m <<= 1;
  for (int i = mc + tid; i < m; i += kNmsThreads) keys[i] = ~0u;
  __syncthreads();
  bitonic_asc_u32<kNmsThreads>(keys, m);

  // ---- C. boxes in (class, score) order, segment starts (s_score is dead from here on: seg / dead reuse it)
  for (int i = tid; i < mc; i += kNmsThreads) {
    sbox[i] = gbox[slot_of((int)cand_of[keys[i] & 0x1fff])];
    dead[i] = 0;
  }
  __syncthreads();
  for (int i0 = 0; i0 < mc; i0 += kNmsThreads) {
    const int i = i0 + tid;
    const bool start = (i < mc) && (i == 0 || (keys[i] >> 13) != (keys[i - 1] >> 13));
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

  // ---- D. greedy suppression per class segment
  nms_suppress(sbox, dead, seg, s_nseg, mc, A.thr, s_sup);

  // ---- E. survivors -> bits of the image's rank bitmap
  unsigned* g_alive = A.alive + (int64_t)n * (kNmsCap / 32);
  for (int i = tid; i < mc; i += kNmsThreads)
    if (!dead[i]) {
      const uint32_t r = keys[i] & 0x1fff;
      atomicOr(&g_alive[r >> 5], 1u << (r & 31));
    }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&A.done[n], 1u) == (unsigned)S - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- F. last CTA of the image: ranks of all survivors in order = prefix popcount over the bitmap
  if (tid == 0) A.done[n] = 0u;
  constexpr int kWords = kNmsCap / 32;          // 256
  uint16_t* sel = cand_of;                      // the first max_out survivors' concatenation indices, in score order
  unsigned word = 0u;
  if (tid < kWords) word = __ldcg(&g_alive[tid]);
  int cnt = __popc(word), incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < kWords / 32; ++w) {
    if (w < wid) before += s_warp[w];
    total += s_warp[w];
  }
  const int limit = (A.max_out > 0 && A.max_out < total) ? A.max_out : total;
  __syncthreads();   // cand_of (this CTA's own ranks) is overwritten below; everybody is past phase C
  if (tid < kWords) {
    int pos = before + incl - cnt;
    while (word != 0u && pos < limit) {
      const int bit = __ffs(word) - 1;
      word &= word - 1u;
      sel[pos++] = __ldcg(&g_r2c[tid * 32 + bit]);
    }
  }
  __syncthreads();
  nms_emit(A, n, limit, s_bad != 0, L, s_pref, s_warp, [&](int t) { return (int)sel[t]; });
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

// FSG_NMS_SORT=1: always the sorting kernel (A/B timing of the merge-rank kernel)
static bool nms_force_sort() {
  static const int v = [] {
    const char* e = getenv("FSG_NMS_SORT");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  return v != 0;
}

int launch_nms_image(const NmsArgs& a, int N, cudaStream_t s, bool pdl) {
  // the merge-rank kernel holds every run's scores of an image in one CTA and packs ranks into 13 bits: all
  // L * topk slots must fit kNmsCap (config 4: 5 x 1000; the FPN RPN's 5 x 2000 does not and keeps the sort kernel)
  if (a.sorted_runs && a.lvl_count && a.classes && a.alive && a.rank2cand && (int64_t)a.L * a.topk <= kNmsCap &&
      !nms_force_sort()) {
    FSG_CUDA_TRY(cudaFuncSetAttribute(nms_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRunsSmem));
    launch_pdl(nms_runs_kernel, dim3((unsigned)a.split, (unsigned)N), dim3(kNmsThreads), kRunsSmem, s, pdl, a);
    FSG_LAUNCH_CHECK();
    return FSG_OK;
  }
  FSG_CUDA_TRY(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmem));
  launch_pdl(nms_image_kernel, dim3((unsigned)a.split, (unsigned)N), dim3(kNmsThreads), kNmsSmem, s, pdl, a);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
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
  return launch_nms_image(a, 1, s);
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
