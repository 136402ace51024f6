// nms / batched_nms for more boxes than one CTA's shared memory holds (n > 8192): the general-n form of
// torchvision.ops.nms semantics behind detectron2/layers/nms.py:6,9-26 (callers with large n: RPN
// find_top_rpn_proposals, rpn_outputs.py:137; fast_rcnn_inference_single_image, fast_rcnn.py:109).
//
// Three launches, everything stays on the device (torchvision's CUDA op copies the n^2/64 mask to the host and
// sweeps it there):
//   1. nms_rank_kernel   stable score-descending order by counting (rank[i] = #keys smaller than key[i]; the
//                        key carries the index, so keys are unique) and a gather of boxes/classes into that order;
//   2. nms_mask_kernel   upper-triangular suppression bit matrix, 64 x 64 tile per CTA (same fp32 IoU op order
//                        and strict `>` as the small-n kernel; class equality folded in for batched_nms);
//   3. nms_sweep_kernel  one CTA walks the rows in super-blocks of 1024: diagonal blocks resolved from shared
//                        memory on warp shuffles, kept rows OR-ed into the `removed` bitset (see the kernel).
#include <math.h>

#include "common.cuh"
#include "nms_large.cuh"

namespace fsg {

constexpr int kRankThreads = 256;
constexpr int kRankTile = 2048;
constexpr int kSweepThreads = 1024;

__device__ __forceinline__ uint64_t order_key(float score, uint32_t idx) {
  const uint32_t sb = __float_as_uint(score);
  const uint32_t ord = (sb & 0x80000000u) ? ~sb : (sb | 0x80000000u);  // order-preserving for any float
  return ((uint64_t)(0xffffffffu - ord) << 32) | (uint64_t)idx;        // ascending = score desc, index asc
}

__global__ void __launch_bounds__(kRankThreads) nms_rank_kernel(const float4* __restrict__ boxes,
                                                                const float* __restrict__ scores,
                                                                const int64_t* __restrict__ classes, int n,
                                                                float4* __restrict__ sbox,
                                                                int64_t* __restrict__ scls,
                                                                int* __restrict__ sidx) {
  __shared__ uint64_t tile[kRankTile];
  const int i = blockIdx.x * kRankThreads + threadIdx.x;
  const uint64_t mine = (i < n) ? order_key(scores[i], (uint32_t)i) : ~0ull;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += kRankTile) {
    const int len = min(kRankTile, n - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += kRankThreads) tile[t] = order_key(scores[j0 + t], (uint32_t)(j0 + t));
    __syncthreads();
    int c = 0;
#pragma unroll 8
    for (int t = 0; t < len; ++t) c += (tile[t] < mine) ? 1 : 0;
    rank += c;
  }
  if (i < n) {
    sbox[rank] = boxes[i];
    if (classes) scls[rank] = classes[i];
    sidx[rank] = i;
  }
}

__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ sbox,
                                                      const int64_t* __restrict__ scls, int n, int nb, float thr,
                                                      uint64_t* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  __shared__ float4 cbox[64];
  __shared__ int64_t ccls[64];
  const int t = threadIdx.x;
  const int cj = cb * 64 + t;
  if (cj < n) {
    cbox[t] = sbox[cj];
    ccls[t] = scls ? scls[cj] : 0;
  }
  __syncthreads();
  const int i = rb * 64 + t;
  if (i >= n) return;
  const float4 bi = sbox[i];
  const int64_t ci = scls ? scls[i] : 0;
  const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
  const int ncol = min(64, n - cb * 64);
  uint64_t word = 0;
  const int start = (cb == rb) ? t + 1 : 0;
  for (int j = start; j < ncol; ++j) {
    if (ccls[j] == ci && nms_suppresses(bi, ai, cbox[j], thr)) word |= (1ull << j);
  }
  mask[(int64_t)i * nb + cb] = word;
}

// Sweep: one CTA walks the sorted boxes in super-blocks of kSB 64-row blocks (1024 rows).
//   phase 0  the super-block's own kSB x kSB words of the bit matrix are loaded into shared memory (one row per
//            thread, all loads independent);
//   phase 1  warp 0 resolves the kSB diagonal blocks in order: the 64 diagonal words sit in registers and the
//            serial keep / suppress chain runs on shuffles, then the kept rows of the block are OR-ed into the
//            `removed` words of the later blocks of the super-block (shared memory only);
//   phase 2  all threads OR the kept rows of the super-block into the `removed` words behind it, reading the bit
//            matrix with batches of independent loads (rows split over thread groups when few words remain).
constexpr int kSB = 16;
constexpr int kTileStride = kSB + 1;   // words per tile row (+1: conflict-free column reads)

__device__ __forceinline__ uint64_t warp_or64(uint64_t v) {
  const unsigned lo = __reduce_or_sync(kFull, (unsigned)v);
  const unsigned hi = __reduce_or_sync(kFull, (unsigned)(v >> 32));
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const uint64_t* __restrict__ mask,
                                                                  const int* __restrict__ sidx, int n, int nb,
                                                                  int64_t* __restrict__ keep,
                                                                  int32_t* __restrict__ num_keep) {
  extern __shared__ __align__(16) uint64_t smem64[];
  uint64_t* removed = smem64;                                  // nb words
  uint64_t* tile = smem64 + nb;                                // 1024 * kTileStride words
  int* rowlist = reinterpret_cast<int*>(tile + 1024 * kTileStride);   // 1024 kept rows of the super-block
  __shared__ int s_count, s_nrows;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int w = tid; w < nb; w += kSweepThreads) removed[w] = 0ull;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int sb0 = 0; sb0 < nb; sb0 += kSB) {
    const int nblk = min(kSB, nb - sb0);
    const int row0 = sb0 * 64;
    const int rows = min(nblk * 64, n - row0);
    // ---- phase 0
    if (tid < rows) {
      const uint64_t* src = mask + (int64_t)(row0 + tid) * nb + sb0;
      uint64_t v[kSB];
#pragma unroll
      for (int j = 0; j < kSB; ++j) v[j] = (j < nblk && j >= (tid >> 6)) ? src[j] : 0ull;   // upper triangle only
#pragma unroll
      for (int j = 0; j < kSB; ++j) tile[tid * kTileStride + j] = v[j];
    }
    if (tid == 0) s_nrows = 0;
    __syncthreads();
    // ---- phase 1
    if (wid == 0) {
      int cnt = s_count, nrows = 0;
      for (int b = 0; b < nblk; ++b) {
        const int r_lo = b * 64 + lane, r_hi = r_lo + 32;
        const uint64_t d_lo = (r_lo < rows) ? tile[r_lo * kTileStride + b] : 0ull;
        const uint64_t d_hi = (r_hi < rows) ? tile[r_hi * kTileStride + b] : 0ull;
        uint64_t rem = removed[sb0 + b];
        const int rb_rows = min(64, rows - b * 64);
        if (rb_rows < 64) rem |= ~0ull << rb_rows;             // rows past the end never survive
        uint64_t kept = 0ull;
#pragma unroll 8
        for (int t = 0; t < 64; ++t) {
          const uint64_t dt = __shfl_sync(kFull, (t < 32) ? d_lo : d_hi, t & 31);
          if (!((rem >> t) & 1ull)) {
            kept |= 1ull << t;
            rem |= dt;
          }
        }
        const bool k_lo = (kept >> lane) & 1ull, k_hi = (kept >> (lane + 32)) & 1ull;
        for (int j = b + 1; j < nblk; ++j) {
          uint64_t v = (k_lo ? tile[r_lo * kTileStride + j] : 0ull) | (k_hi ? tile[r_hi * kTileStride + j] : 0ull);
          v = warp_or64(v);
          if (lane == 0) removed[sb0 + j] |= v;
        }
        __syncwarp();
        const int p_lo = __popcll(kept & ((1ull << lane) - 1ull));
        const int p_hi = __popcll(kept & ((1ull << (lane + 32)) - 1ull));
        if (k_lo) { keep[cnt + p_lo] = (int64_t)sidx[row0 + r_lo]; rowlist[nrows + p_lo] = row0 + r_lo; }
        if (k_hi) { keep[cnt + p_hi] = (int64_t)sidx[row0 + r_hi]; rowlist[nrows + p_hi] = row0 + r_hi; }
        const int c = __popcll(kept);
        cnt += c;
        nrows += c;
      }
      if (lane == 0) { s_count = cnt; s_nrows = nrows; }
    }
    __syncthreads();
    // ---- phase 2
    const int w0 = sb0 + nblk;
    const int W = nb - w0;
    const int K = s_nrows;
    if (W > 0 && K > 0) {
      int span = 1;                      // threads along the word axis (power of two)
      while (span < W && span < kSweepThreads) span <<= 1;
      const int groups = kSweepThreads / span;   // thread groups along the row axis
      const int g = tid / span, wi = tid - g * span;
      for (int w = w0 + wi; w < nb; w += span) {
        uint64_t acc = 0ull;
        int q = g;
        for (; q + 7 * groups < K; q += 8 * groups) {
          uint64_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = mask[(int64_t)rowlist[q + u * groups] * nb + w];
#pragma unroll
          for (int u = 0; u < 8; ++u) acc |= v[u];
        }
        for (; q < K; q += groups) acc |= mask[(int64_t)rowlist[q] * nb + w];
        if (acc) atomicOr(reinterpret_cast<unsigned long long*>(&removed[w]), (unsigned long long)acc);
      }
    }
    __syncthreads();
  }
  if (tid == 0) *num_keep = s_count;
}

NmsLargeWs nms_large_ws_layout(int64_t n) {
  NmsLargeWs w;
  const int64_t nb = ceil_div(n, 64);
  size_t o = 0;
  w.off_box = o;  o += align_up(sizeof(float4) * (size_t)n, 16);
  w.off_cls = o;  o += align_up(sizeof(int64_t) * (size_t)n, 16);
  w.off_idx = o;  o += align_up(sizeof(int) * (size_t)n, 16);
  w.off_mask = o; o += align_up(sizeof(uint64_t) * (size_t)n * (size_t)nb, 16);
  w.total = o;
  return w;
}

int nms_large(const float* boxes, const float* scores, const int64_t* class_ids, int64_t n, float thr,
              int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (n > kNmsLargeMax) return FSG_ERR_UNSUPPORTED;
  const NmsLargeWs w = nms_large_ws_layout(n);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  const int nb = (int)ceil_div(n, 64);
  float4* sbox = (float4*)(ws + w.off_box);
  int64_t* scls = class_ids ? (int64_t*)(ws + w.off_cls) : nullptr;
  int* sidx = (int*)(ws + w.off_idx);
  uint64_t* mask = (uint64_t*)(ws + w.off_mask);
  nms_rank_kernel<<<(unsigned)ceil_div(n, kRankThreads), kRankThreads, 0, s>>>((const float4*)boxes, scores,
                                                                               class_ids, (int)n, sbox, scls, sidx);
  FSG_LAUNCH_CHECK();
  nms_mask_kernel<<<dim3((unsigned)nb, (unsigned)nb), 64, 0, s>>>(sbox, scls, (int)n, nb, thr, mask);
  FSG_LAUNCH_CHECK();
  const size_t sweep_smem = sizeof(uint64_t) * ((size_t)nb + 1024 * kTileStride) + sizeof(int) * 1024;
  FSG_CUDA_TRY(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
  nms_sweep_kernel<<<1, kSweepThreads, sweep_smem, s>>>(mask, sidx, (int)n, nb, keep, num_keep);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

}  // namespace fsg
