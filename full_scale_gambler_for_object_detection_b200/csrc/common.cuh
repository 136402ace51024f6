// Shared device helpers for the sm_100a dense-detection kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fsg_dense.h"

#define FSG_LAUNCH_CHECK()                                         \
  do {                                                             \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return FSG_ERR_CUDA + (int)e__;        \
  } while (0)

#define FSG_CUDA_TRY(expr)                                         \
  do {                                                             \
    cudaError_t e__ = (expr);                                      \
    if (e__ != cudaSuccess) return FSG_ERR_CUDA + (int)e__;        \
  } while (0)

namespace fsg {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- streaming global access (read-once / write-once data: keep it out of L1) -------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream2(float* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk; SASS: UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- NMS pair test ------------------------------------------------------------------------
// `fl(inter / uni) > thr` (torchvision's test) without the IEEE division whenever the answer is clear: a correctly
// rounded quotient q satisfies |q - inter/uni| <= 2^-24 * inter/uni, so inter outside thr*uni*(1 +- 2^-20) decides it;
// only pairs inside that band (and the degenerate uni <= 0 / NaN cases) take the exact division.
__device__ __forceinline__ bool nms_suppresses(float4 bi, float ai, float4 bj, float thr) {
  const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  const float w = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
  if (uni > 0.f && thr >= 0.f) {
    const float c = thr * uni;
    if (inter < c * 0.999999f) return false;
    if (inter > c * 1.000001f && c > 1e-30f) return true;
  }
  return __fdiv_rn(inter, uni) > thr;
}

// ---- programmatic dependent launch: device side --------------------------------------------------
// No-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void grid_dependency_sync() {
#if __CUDA_ARCH__ >= 900
  cudaGridDependencySynchronize();
#endif
}
__device__ __forceinline__ void grid_launch_dependents() {
#if __CUDA_ARCH__ >= 900
  cudaTriggerProgrammaticLaunchCompletion();
#endif
}

// Pointers to what the PRECEDING kernel of a programmatic-dependent-launch chain produced go through this right after
// grid_dependency_sync().  nvcc treats loads through `const __restrict__` pointers as invariant for the kernel's
// lifetime (ld.global.nc) and is free to hoist them above griddepcontrol.wait -- seen in SASS as an LDG of pass A's
// best IoUs in front of ACQBULK, i.e. read while pass A was still running.  The empty volatile asm makes the pointer
// value opaque until the wait has executed (volatile asm statements keep their order); profiles/pdl_audit.py lists
// the global loads every kernel of the library issues before its wait.
template <typename T>
__device__ __forceinline__ T* produced_by_dependency(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// L2 prefetch of one 128-byte line (fire and forget: no register, no scoreboard entry)
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- launch with or without programmatic dependent launch -------------------------------------
// pdl: the kernel may be scheduled while the kernel in front of it in the stream is still draining; it must call
// cudaGridDependencySynchronize() before touching anything that kernel wrote (captured into CUDA graphs as a
// programmatic dependency edge).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- shared-memory histogram update of one warp round ---------------------------------------
// Radix-select digits of score keys are either almost all equal inside a warp (the leading digits: keys share sign
// and exponent) or almost all different (the trailing digits).  Up to two groups of equal bins are added with one
// aggregated atomic each; whatever is left goes through plain shared-memory atomics, where distinct bins do not
// conflict.  (match.any resolves one distinct value per step -- 32 steps for a warp of distinct bins.)
__device__ __forceinline__ void warp_hist_add(unsigned* hist, unsigned bin, bool ok) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const unsigned act = __ballot_sync(kFull, ok);
    if (act == 0u) return;
    const int leader = __ffs(act) - 1;
    const unsigned b0 = __shfl_sync(kFull, bin, leader);
    const unsigned same = __ballot_sync(kFull, ok && bin == b0);
    if (lane == leader) atomicAdd(&hist[b0], (unsigned)__popc(same));
    ok = ok && bin != b0;
  }
  if (ok) atomicAdd(&hist[bin], 1u);
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace fsg
