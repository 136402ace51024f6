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

// (N,R)-sized per-anchor maps between the flattened anchor order and the per-level (N, A, H, W) maps the
// gambler produces / consumes (betting maps in, NAKHW_loss and d/d bets out; gambler_heads.py:91-101,291-318).
// A few MB in total, so one launch covers all levels and up to three tensors.
struct SmallMaps {
  float* lvl[3][FSG_MAX_LEVELS];   // per tensor, per level: (N, A, H, W)
  float* flat[3];                  // per tensor: (N, R)
  int64_t off[FSG_MAX_LEVELS + 1];
  int HW[FSG_MAX_LEVELS];
  int tile_base[FSG_MAX_LEVELS + 1];   // tiles of 256 hw positions per level
  int num_levels, A, ntensors, to_levels;
  int64_t R;
};

__global__ void __launch_bounds__(256) anchor_maps_kernel(const SmallMaps M) {
  // one CTA: 256 consecutive hw positions of one level and image, all A anchor slots, staged through shared
  // memory so that both the per-plane side and the flat side are accessed with consecutive addresses
  __shared__ float s_t[FSG_MAX_CELL_ANCHORS][257];
  const int n = blockIdx.y, tid = threadIdx.x;
  int l = 0;
  while (l + 1 < M.num_levels && (int)blockIdx.x >= M.tile_base[l + 1]) ++l;
  const int64_t hw0 = (int64_t)((int)blockIdx.x - M.tile_base[l]) * 256;
  const int HW = M.HW[l], A = M.A;
  const int cnt = (int)min((int64_t)256, HW - hw0);          // hw positions in this tile
  const int64_t fbase = (int64_t)n * M.R + M.off[l] + hw0 * A;
  for (int t = 0; t < M.ntensors; ++t) {
    float* lv = M.lvl[t][l] + (int64_t)n * A * HW + hw0;
    float* fl = M.flat[t] + fbase;
    if (!M.to_levels) {
      if (tid < cnt)
        for (int a = 0; a < A; ++a) s_t[a][tid] = lv[(int64_t)a * HW + tid];
      __syncthreads();
      for (int i = tid; i < cnt * A; i += 256) fl[i] = s_t[i % A][i / A];
    } else {
      for (int i = tid; i < cnt * A; i += 256) s_t[i % A][i / A] = fl[i];
      __syncthreads();
      if (tid < cnt)
        for (int a = 0; a < A; ++a) lv[(int64_t)a * HW + tid] = s_t[a][tid];
    }
    __syncthreads();
  }
}

// DefaultAnchorGenerator.grid_anchors (detectron2/modeling/anchor_generator.py:41-50,121-129): anchor
// (level, y, x, a) = (x*stride, y*stride, x*stride, y*stride) + cell_anchor[a] in fp32, flattened in the order
// (level, y, x, a) = the r index every other kernel uses.  x*stride is an exact fp32 integer (arange with an integer
// step), so one rounded addition reproduces the reference bit for bit.
struct AnchorGrid {
  int64_t off[FSG_MAX_LEVELS + 1];
  int W[FSG_MAX_LEVELS], stride[FSG_MAX_LEVELS], A[FSG_MAX_LEVELS];
  float cell[FSG_MAX_LEVELS][FSG_MAX_CELL_ANCHORS][4];
  int num_levels;
};

__global__ void __launch_bounds__(256) grid_anchors_kernel(const AnchorGrid G, float4* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= G.off[G.num_levels]) return;
  int l = 0;
  while (l + 1 < G.num_levels && r >= G.off[l + 1]) ++l;
  const int64_t local = r - G.off[l];
  const int A = G.A[l];
  const int64_t cellidx = local / A;
  const int a = (int)(local - cellidx * A);
  const int y = (int)(cellidx / G.W[l]);
  const int x = (int)(cellidx - (int64_t)y * G.W[l]);
  const float sx = (float)(x * G.stride[l]), sy = (float)(y * G.stride[l]);
  const float* c = G.cell[l][a];
  out[r] = make_float4(__fadd_rn(sx, c[0]), __fadd_rn(sy, c[1]), __fadd_rn(sx, c[2]), __fadd_rn(sy, c[3]));
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_grid_anchors(const fsg_anchor_level* h_levels, int num_levels, float* anchors, int64_t R,
                                fsg_stream_t stream) {
  if (!h_levels || num_levels <= 0 || num_levels > FSG_MAX_LEVELS || R < 0) return FSG_ERR_INVALID_ARG;
  AnchorGrid g;
  int64_t off = 0;
  for (int l = 0; l < num_levels; ++l) {
    const fsg_anchor_level& L = h_levels[l];
    if (L.H < 0 || L.W < 0 || L.stride <= 0 || L.A <= 0 || L.A > FSG_MAX_CELL_ANCHORS) return FSG_ERR_INVALID_ARG;
    if ((int64_t)(L.H > L.W ? L.H : L.W) * L.stride >= (1 << 24)) return FSG_ERR_UNSUPPORTED;  // exact in fp32
    g.off[l] = off; g.W[l] = L.W > 0 ? L.W : 1; g.stride[l] = L.stride; g.A[l] = L.A;
    for (int a = 0; a < FSG_MAX_CELL_ANCHORS; ++a)
      for (int j = 0; j < 4; ++j) g.cell[l][a][j] = a < L.A ? L.cell[a][j] : 0.f;
    off += (int64_t)L.H * L.W * L.A;
  }
  for (int l = num_levels; l <= FSG_MAX_LEVELS; ++l) g.off[l] = off;
  g.num_levels = num_levels;
  if (off != R) return FSG_ERR_INVALID_ARG;
  if (R == 0) return FSG_OK;
  if (!anchors || ((uintptr_t)anchors & 15)) return FSG_ERR_INVALID_ARG;
  grid_anchors_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(g, (float4*)anchors);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}


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

extern "C" int fsg_anchor_maps(float* const* h_level_ptrs, float* const* h_flat_ptrs, int ntensors,
                               const int32_t* h_HW, int num_levels, int A, int N, int to_levels,
                               fsg_stream_t stream) {
  if (!h_level_ptrs || !h_flat_ptrs || !h_HW || ntensors <= 0 || ntensors > 3 || num_levels <= 0 ||
      num_levels > FSG_MAX_LEVELS || A <= 0 || A > FSG_MAX_CELL_ANCHORS || N <= 0 || N > 65535)
    return FSG_ERR_INVALID_ARG;
  SmallMaps m;
  int64_t off = 0;
  int tiles = 0;
  for (int l = 0; l < FSG_MAX_LEVELS; ++l) {
    m.off[l] = off;
    m.tile_base[l] = tiles;
    if (l < num_levels && h_HW[l] > 0) tiles += (int)ceil_div(h_HW[l], 256);
    m.HW[l] = l < num_levels ? h_HW[l] : 0;
    if (l < num_levels) {
      if (h_HW[l] < 0) return FSG_ERR_INVALID_ARG;
      off += (int64_t)h_HW[l] * A;
    }
    for (int t = 0; t < 3; ++t) {
      m.lvl[t][l] = (t < ntensors && l < num_levels) ? h_level_ptrs[t * num_levels + l] : nullptr;
      if (t < ntensors && l < num_levels && h_HW[l] > 0 && !m.lvl[t][l]) return FSG_ERR_INVALID_ARG;
    }
  }
  m.off[FSG_MAX_LEVELS] = off;
  m.tile_base[FSG_MAX_LEVELS] = tiles;
  for (int t = 0; t < 3; ++t) {
    m.flat[t] = t < ntensors ? h_flat_ptrs[t] : nullptr;
    if (t < ntensors && !m.flat[t]) return FSG_ERR_INVALID_ARG;
  }
  m.num_levels = num_levels; m.A = A; m.ntensors = ntensors; m.to_levels = to_levels; m.R = off;
  if (off == 0) return FSG_OK;
  anchor_maps_kernel<<<dim3((unsigned)tiles, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(m);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
