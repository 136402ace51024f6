// Block-wide bitonic sorting networks over 64-bit keys in shared memory (used by the top-k select, the NMS kernel and
// the RPN proposal select).
#pragma once
#include "common.cuh"

namespace fsg {

template <int NT>
__device__ void bitonic_desc(uint64_t* a, int m) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += NT) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1))   /* stride is a power of two */;
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t x = a[lo], y = a[hi];
        if (desc ? (x < y) : (x > y)) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

template <int NT>
__device__ void bitonic_asc_smem(uint64_t* a, int m) {
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (m >> 1); t += NT) {
        const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1))   /* stride is a power of two */;
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
          const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1))   /* stride is a power of two */;
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

// 32-bit keys, ascending, block-wide; same scheme as bitonic_asc (two keys per thread in registers, shuffles below
// stride 64) for 64 <= m <= 2*NT, plain shared-memory network otherwise.
template <int NT>
__device__ void bitonic_asc_u32(uint32_t* a, int m) {
  if (m < 64 || m > 2 * NT) {
    for (int size = 2; size <= m; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (m >> 1); t += NT) {
          const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
          const int hi = lo + stride;
          const bool asc = ((lo & size) == 0);
          const uint32_t x = a[lo], y = a[hi];
          if (asc ? (x > y) : (x < y)) { a[lo] = y; a[hi] = x; }
        }
        __syncthreads();
      }
    }
    return;
  }
  const int t = threadIdx.x;
  const bool act = t < (m >> 1);
  const int i0 = 2 * t;
  uint32_t v0 = act ? a[i0] : 0u, v1 = act ? a[i0 + 1] : 0u;
  for (int size = 2; size <= m; size <<= 1) {
    const bool asc = ((i0 & size) == 0);
    int stride = size >> 1;
    if (stride >= 64) {
      if (act) { a[i0] = v0; a[i0 + 1] = v1; }
      __syncthreads();
      for (; stride >= 64; stride >>= 1) {
        if (act) {
          const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const uint32_t x = a[lo], y = a[hi];
          if (up ? (x > y) : (x < y)) { a[lo] = y; a[hi] = x; }
        }
        __syncthreads();
      }
      if (act) { v0 = a[i0]; v1 = a[i0 + 1]; }
    }
    for (; stride >= 2; stride >>= 1) {
      const int pl = stride >> 1;
      const uint32_t p0 = __shfl_xor_sync(kFull, v0, pl);
      const uint32_t p1 = __shfl_xor_sync(kFull, v1, pl);
      const bool keep_min = (((i0 & stride) == 0) == asc);
      v0 = keep_min ? min(v0, p0) : max(v0, p0);
      v1 = keep_min ? min(v1, p1) : max(v1, p1);
    }
    if ((v0 > v1) == asc) { const uint32_t tmp = v0; v0 = v1; v1 = tmp; }
  }
  if (act) { a[i0] = v0; a[i0 + 1] = v1; }
  __syncthreads();
}

// Per-class NMS is independent across classes, so an image is split over `split` CTAs by class id; each
// sorts and suppresses only its own candidates (4x fewer keys per bitonic network at split = 4) and hands
// its best survivors to the last CTA of the image, which merges them by score.
}  // namespace fsg
