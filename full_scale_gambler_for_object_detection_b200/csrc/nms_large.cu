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
//   3. nms_sweep_kernel  one CTA walks the 64-row blocks in order: resolves the diagonal word serially, then
//                        ORs the kept rows into the `removed` bitset held in shared memory.
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
    const float4 bj = cbox[j];
    const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    const float w = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
    if (ovr > thr && ccls[j] == ci) word |= (1ull << j);
  }
  mask[(int64_t)i * nb + cb] = word;
}

__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const uint64_t* __restrict__ mask,
                                                                  const int* __restrict__ sidx, int n, int nb,
                                                                  int64_t* __restrict__ keep,
                                                                  int32_t* __restrict__ num_keep) {
  extern __shared__ uint64_t removed[];  // nb words
  __shared__ uint64_t s_diag[64];
  __shared__ uint64_t s_kept;
  __shared__ int s_count;
  const int tid = threadIdx.x;
  for (int w = tid; w < nb; w += kSweepThreads) removed[w] = 0ull;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int rb = 0; rb < nb; ++rb) {
    const int rows = min(64, n - rb * 64);
    if (tid < 64) s_diag[tid] = (tid < rows) ? mask[(int64_t)(rb * 64 + tid) * nb + rb] : 0ull;
    __syncthreads();
    if (tid == 0) {
      uint64_t rem = removed[rb], kept = 0ull;
      for (int t = 0; t < rows; ++t) {
        if (!((rem >> t) & 1ull)) {
          kept |= (1ull << t);
          rem |= s_diag[t];
        }
      }
      s_kept = kept;
    }
    __syncthreads();
    const uint64_t kept = s_kept;
    const int base = s_count;
    // kept rows -> output (score-descending order = sorted order) and into the bitset of later blocks
    if (tid < 64 && ((kept >> tid) & 1ull)) {
      const int pos = base + __popcll(kept & ((1ull << tid) - 1ull));
      keep[pos] = (int64_t)sidx[rb * 64 + tid];
    }
    if (kept != 0ull) {
      for (int w = rb + 1 + tid; w < nb; w += kSweepThreads) {
        uint64_t acc = 0ull;
        uint64_t k = kept;
        while (k) {
          const int t = __ffsll((long long)k) - 1;
          k &= k - 1ull;
          acc |= mask[(int64_t)(rb * 64 + t) * nb + w];
        }
        removed[w] |= acc;
      }
    }
    __syncthreads();
    if (tid == 0) s_count = base + __popcll(kept);
    // the next iteration's first barrier orders this write before any read of s_count
  }
  __syncthreads();
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
  nms_sweep_kernel<<<1, kSweepThreads, sizeof(uint64_t) * (size_t)nb, s>>>(mask, sidx, (int)n, nb, keep, num_keep);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}

}  // namespace fsg
