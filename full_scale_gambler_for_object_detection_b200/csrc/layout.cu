// Head-layout adapter: (N, A*K, H, W) conv output  <->  flattened (N, sum_l H*W*A, K) rows.
//
// Reference: permute_to_N_HWA_K / permute_all_cls_and_box_to_N_HWA_K_and_concat
// (detectron2/modeling/meta_arch/retinanet.py:24-54) and the fork's twins
// (ImbalanceDetection/imbalancedetection/gambler_heads.py:34-101).  With C = A*K channels the
// permutation (N,A,K,H,W) -> (N,H,W,A,K) is a plain (C x HW) -> (HW x C) matrix transpose per
// image; anchor index (h*W+w)*A+a and channel a*K+k fall out of that directly.  Tiled through
// shared memory (32x33 padding) so both the read and the write are coalesced.
#include "common.cuh"

namespace fsg {

constexpr int kTile = 32;

// to_nchw == 0: nchw (N,C,HW) -> flat[n*flat_stride + flat_off + hw*C + c]
// to_nchw == 1: the inverse (used for gradients)
__global__ void __launch_bounds__(256) permute_level_kernel(float* __restrict__ nchw, float* __restrict__ flat,
                                                            int C, int64_t HW, int64_t flat_stride,
                                                            int64_t flat_off, int to_nchw) {
  __shared__ float tile[kTile][kTile + 1];
  const int n = blockIdx.z;
  const int64_t hw0 = (int64_t)blockIdx.x * kTile;
  const int c0 = blockIdx.y * kTile;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float* src_img = nchw + (int64_t)n * C * HW;
  float* dst_img = flat + (int64_t)n * flat_stride + flat_off;
  if (!to_nchw) {
#pragma unroll
    for (int k = 0; k < kTile; k += 8) {
      const int c = c0 + ty + k;
      const int64_t hw = hw0 + tx;
      if (c < C && hw < HW) tile[ty + k][tx] = src_img[(int64_t)c * HW + hw];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile; k += 8) {
      const int64_t hw = hw0 + ty + k;
      const int c = c0 + tx;
      if (c < C && hw < HW) dst_img[hw * C + c] = tile[tx][ty + k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < kTile; k += 8) {
      const int64_t hw = hw0 + ty + k;
      const int c = c0 + tx;
      if (c < C && hw < HW) tile[ty + k][tx] = dst_img[hw * C + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile; k += 8) {
      const int c = c0 + ty + k;
      const int64_t hw = hw0 + tx;
      if (c < C && hw < HW) src_img[(int64_t)c * HW + hw] = tile[tx][ty + k];
    }
  }
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_permute_level(float* nchw, float* flat, int N, int C, int64_t HW, int64_t flat_image_stride,
                                 int64_t flat_offset, int to_nchw, fsg_stream_t stream) {
  if (N < 0 || C <= 0 || HW < 0 || flat_offset < 0) return FSG_ERR_INVALID_ARG;
  if (N == 0 || HW == 0) return FSG_OK;
  if (!nchw || !flat) return FSG_ERR_INVALID_ARG;
  if (N > 65535 || ceil_div(C, kTile) > 65535) return FSG_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(HW, kTile), (unsigned)ceil_div(C, kTile), (unsigned)N);
  permute_level_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nchw, flat, C, HW, flat_image_stride, flat_offset,
                                                               to_nchw);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
