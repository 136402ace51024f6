// Two-stage callers of the NMS kernel: batched RPN proposal selection and the candidate
// stage of fast_rcnn_inference_single_image.
//
// Reference: detectron2/modeling/proposal_generator/rpn_outputs.py:52-151; roi_heads/fast_rcnn.py:76-118.
#include "nms_kernel.cuh"
#include "nms_large.cuh"
#include "sort_utils.cuh"

namespace fsg {

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
constexpr int kRpnMaxK = 16384;        // pre_nms_topk per level (128 KB of keys); C4 models train with 12000

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
  int fill_tail;                        // large path: unused slots of a level get score -inf and an empty box
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
        warp_hist_add(hist, digit, in);
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
  if (A.fill_tail) {
    // the general-n NMS behind this kernel takes a fixed number of boxes: pad with boxes that sort last (score -inf)
    // and never suppress or get suppressed (zero area: IoU 0 or NaN)
    const int cnt = s_base;
    for (int j = cnt + tid; j < A.topk; j += kRpnThreads) {
      A.cand_box[slot0 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
      A.cand_score[slot0 + j] = -INFINITY;
      A.cand_class[slot0 + j] = l;
    }
  }
}

// large path epilogue: keep[] of the general-n NMS (score order; the padding sorts last) -> the post_nms_topk best
// real survivors of image n
__global__ void __launch_bounds__(256) rpn_gather_kernel(const float4* __restrict__ cand_box,
                                                         const float* __restrict__ cand_score,
                                                         const int64_t* __restrict__ cand_class,
                                                         const int* __restrict__ lvl_count, int L, int64_t slots,
                                                         const int64_t* __restrict__ keep,
                                                         const int32_t* __restrict__ num_keep, int post,
                                                         float4* __restrict__ out_boxes, float* __restrict__ out_logits,
                                                         int64_t* __restrict__ out_levels, int32_t* __restrict__ out_count) {
  const int n = blockIdx.x;
  int real = 0;
  for (int l = 0; l < L; ++l) real += lvl_count[n * L + l];
  int kept = num_keep[n] - (int)(slots - real);   // every padded slot is "kept" and sits behind the real ones
  if (kept < 0) kept = 0;
  if (kept > post) kept = post;
  for (int t = threadIdx.x; t < post; t += 256) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int64_t lv = 0;
    if (t < kept) {
      const int64_t i = keep[(int64_t)n * slots + t];
      b = cand_box[(int64_t)n * slots + i];
      sc = cand_score[(int64_t)n * slots + i];
      lv = cand_class[(int64_t)n * slots + i];
    }
    out_boxes[(int64_t)n * post + t] = b;
    out_logits[(int64_t)n * post + t] = sc;
    if (out_levels) out_levels[(int64_t)n * post + t] = lv;
  }
  if (threadIdx.x == 0) out_count[n] = kept;
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

// -> FSG_OK and *large_out = 0: everything fits the shared-memory NMS kernel (<= 8192 candidates per split CTA);
//    *large_out = 1: more candidates per image (e.g. RPN.PRE_NMS_TOPK_TRAIN = 12000 of the C4 models,
//    config/defaults.py:219): same select kernel, then the general-n NMS (nms_large.cu) per image.
static int rpn_plan(const int64_t* h_level_sizes, int num_levels, int pre_nms_topk, int post_nms_topk, int N,
                    int* k, int* topk_out, int* split_out, int* large_out) {
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
  if (topk > kRpnMaxK || N > 65535) return FSG_ERR_UNSUPPORTED;
  if (topk < 1) topk = 1;
  *topk_out = topk;
  *large_out = 0;
  int split = 0;
  if (topk <= kNmsCap && total < (1 << 14) && post_nms_topk <= kNmsCap)
    split = rpn_split_for(N, k, num_levels, post_nms_topk);
  if (split == 0) {
    if ((int64_t)num_levels * topk > kNmsLargeMax) return FSG_ERR_UNSUPPORTED;
    *large_out = 1;
    split = 1;
  }
  *split_out = split;
  return FSG_OK;
}

struct RpnLargeWs {
  size_t off_keep, off_numkeep, off_nms, total;
};
static RpnLargeWs rpn_large_ws_layout(int N, int64_t slots, size_t base) {
  RpnLargeWs w;
  size_t o = base;
  w.off_keep = o;    o += align_up(sizeof(int64_t) * (size_t)N * (size_t)slots, 16);
  w.off_numkeep = o; o += align_up(sizeof(int32_t) * (size_t)N, 16);
  w.off_nms = o;     o += nms_large_ws_layout(slots).total;
  w.total = o;
  return w;
}

extern "C" size_t fsg_rpn_proposals_workspace_bytes(int N, const int64_t* h_level_sizes, int num_levels,
                                                    int pre_nms_topk, int post_nms_topk) {
  int k[kMaxLevels], topk, split, large;
  if (rpn_plan(h_level_sizes, num_levels, pre_nms_topk, post_nms_topk, N, k, &topk, &split, &large) != FSG_OK) return 0;
  const RpnWs w = rpn_ws_layout(N, num_levels, topk, post_nms_topk, split);
  if (!large) return w.total;
  return rpn_large_ws_layout(N, (int64_t)num_levels * topk, w.off_nms).total;
}

extern "C" int fsg_rpn_proposals(const float* const* h_level_proposals, const float* const* h_level_logits,
                                 const int64_t* h_level_sizes, int num_levels, int N, const float* image_sizes,
                                 int pre_nms_topk, int post_nms_topk, double nms_threshold, float min_box_side_len,
                                 float* out_boxes, float* out_logits, int64_t* out_levels, int32_t* out_count,
                                 void* workspace, size_t workspace_bytes, fsg_stream_t stream) {
  int k[kMaxLevels], topk, split, large;
  const int st = rpn_plan(h_level_sizes, num_levels, pre_nms_topk, post_nms_topk, N, k, &topk, &split, &large);
  if (st != FSG_OK) return st;
  if (!h_level_proposals || !h_level_logits || !image_sizes || !out_boxes || !out_logits || !out_count)
    return FSG_ERR_INVALID_ARG;
  if (((uintptr_t)out_boxes) & 15) return FSG_ERR_INVALID_ARG;
  const RpnWs w = rpn_ws_layout(N, num_levels, topk, post_nms_topk, split);
  const int64_t slots = (int64_t)num_levels * topk;
  const RpnLargeWs lw = rpn_large_ws_layout(N, slots, w.off_nms);
  const size_t need = large ? lw.total : w.total;
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
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
  ra.fill_tail = large;
  int m = 1;
  while (m < topk) m <<= 1;
  const size_t smem = sizeof(uint64_t) * (size_t)m;
  FSG_CUDA_TRY(cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rpn_select_kernel<<<dim3((unsigned)num_levels, (unsigned)N), kRpnThreads, smem, s>>>(ra);
  FSG_LAUNCH_CHECK();

  if (large) {
    // general-n NMS per image on its L*topk slots (class id = level => per-level batched_nms), one workspace reused
    // in stream order; then the gather of the post_nms_topk best real survivors
    int64_t* keep = (int64_t*)(ws + lw.off_keep);
    int32_t* numkeep = (int32_t*)(ws + lw.off_numkeep);
    const float thr = threshold_floor(nms_threshold);
    for (int n = 0; n < N; ++n) {
      const int rc = nms_large((const float*)(ra.cand_box + (int64_t)n * slots), ra.cand_score + (int64_t)n * slots,
                               ra.cand_class + (int64_t)n * slots, slots, thr, keep + (int64_t)n * slots, numkeep + n,
                               ws + lw.off_nms, lw.total - lw.off_nms, s);
      if (rc != FSG_OK) return rc;
    }
    rpn_gather_kernel<<<(unsigned)N, 256, 0, s>>>(ra.cand_box, ra.cand_score, ra.cand_class, ra.lvl_count, num_levels,
                                                  slots, keep, numkeep, post_nms_topk, (float4*)out_boxes, out_logits,
                                                  out_levels, out_count);
    FSG_LAUNCH_CHECK();
    return FSG_OK;
  }

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
  a.sorted_runs = 1;   // each level's proposals leave the select kernel sorted by (logit, index); the min-size filter
                       // is a stable compaction
  a.alive = (unsigned*)(nws + nw.off_alive); a.rank2cand = (uint16_t*)(nws + nw.off_r2c);
  a.keep = nullptr; a.keep_stride = post_nms_topk; a.num_keep = out_count;
  a.out_boxes = (float4*)out_boxes; a.out_scores = out_logits; a.out_classes = out_levels;
  return launch_nms_image(a, N, s);
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
