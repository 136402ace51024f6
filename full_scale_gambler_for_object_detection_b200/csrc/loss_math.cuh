// Per-element arithmetic shared by the K2 loss kernels (dense_loss.cu: (N,R,K) layout; dense_loss_levels.cu:
// the head's native (N, A*K, H, W) layout) and the deterministic final reduction of their tile partials.
#pragma once
#include "common.cuh"

namespace fsg {

constexpr int kLossBlock = 256;
constexpr int kPartialStride = 8;  // floats per tile partial: cls, reg, wl, l, maxl

// ---- vector load/store by width -------------------------------------------------------------
template <int V> struct Vec;
template <> struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = ldg_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { stg_stream4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct Vec<2> {
  float v[2];
  __device__ __forceinline__ void load(const float* p) { float2 t = ldg_stream2(p); v[0] = t.x; v[1] = t.y; }
  __device__ __forceinline__ void store(float* p) const { stg_stream2(p, make_float2(v[0], v[1])); }
};
template <> struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = ldg_stream1(p); }
  __device__ __forceinline__ void store(float* p) const { stg_stream1(p, v[0]); }
};

// log1p(e) for e in [0,1]: e * P7(e), max relative error 3.3e-7 (fit in DESIGN.md)
__device__ __forceinline__ float log1p_unit(float e) {
  float p = -8.539245470e-03f;
  p = fmaf(p, e, 4.408976170e-02f);
  p = fmaf(p, e, -1.076818928e-01f);
  p = fmaf(p, e, 1.774525379e-01f);
  p = fmaf(p, e, -2.449546718e-01f);
  p = fmaf(p, e, 3.327548051e-01f);
  p = fmaf(p, e, -4.999740544e-01f);
  p = fmaf(p, e, 9.999998057e-01f);
  return p * e;
}

// MUFU.EX2 / MUFU.RCP without the libdevice range fix-ups (inputs are bounded: argument <= 0, 1+e in [1,2])
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared sub-expressions of sigmoid / BCE for one logit (t = 0 side):
//   p = sigmoid(x), omp = 1 - p (computed as sigmoid(-x), no cancellation), ce0 = softplus(x)
__device__ __forceinline__ void sigmoid_parts(float x, float& p, float& omp, float& ce0) {
  const float e = ex2_approx(fabsf(x) * -1.4426950408889634f);  // exp(-|x|) in (0,1]
  const float r = rcp_approx(1.f + e);
  const float er = e * r;
  const float l1p = log1p_unit(e);
  const bool pos = x >= 0.f;
  p = pos ? r : er;
  omp = pos ? er : r;
  ce0 = fmaxf(x, 0.f) + l1p;
}

// log1p(e) for e in [0,1] with a degree-6 polynomial: max relative error 1.4e-6 (one FMA less than log1p_unit; used
// on the all-negatives fast path only, where it stays an order of magnitude inside the 1e-5 budget)
__device__ __forceinline__ float log1p_unit6(float e) {
  float p = 1.414097077e-02f;
  p = fmaf(p, e, -6.640165794e-02f);
  p = fmaf(p, e, 1.492237392e-01f);
  p = fmaf(p, e, -2.350385509e-01f);
  p = fmaf(p, e, 3.310944840e-01f);
  p = fmaf(p, e, -4.998696571e-01f);
  p = fmaf(p, e, 9.999987316e-01f);
  return p * e;
}

// one element with target t = 0, gamma = 2:  loss/(1-alpha) and dloss/dx/(1-alpha)
__device__ __forceinline__ void focal_neg_g2(float x, float& loss, float& grad) {
  const float e = ex2_approx(fabsf(x) * -1.4426950408889634f);  // exp(-|x|) in (0,1]
  const float r = rcp_approx(1.f + e);
  const float p = (x >= 0.f) ? r : e * r;
  const float ce = fmaxf(x, 0.f) + log1p_unit6(e);
  const float p2 = p * p;
  loss = p2 * ce;
  // 1 - p straight from p: its absolute error (6e-8) only enters through 2(1-p)ce next to p, i.e. at most 1e-6
  // relative for |x| <= 10 -- no second select needed on this path
  grad = p2 * fmaf(2.f * (1.f - p), ce, p);
}

// general element (any t, any gamma, either mode); unscaled by alpha_t
__device__ __forceinline__ void cls_elem_general(float x, bool t, float gamma, float& focal, float& fgrad,
                                                 float& bce, float& bgrad) {
  float p, omp, ce0;
  sigmoid_parts(x, p, omp, ce0);
  const float ce = t ? ce0 - x : ce0;   // BCE-with-logits
  const float q = t ? omp : p;          // 1 - p_t
  const float pt = t ? p : omp;
  const float qg = (gamma == 2.f) ? q * q : ((gamma == 0.f) ? 1.f : powf(q, gamma));
  focal = qg * ce;
  const float inner = fmaf(gamma * pt, ce, q);
  fgrad = t ? -qg * inner : qg * inner;
  bce = ce;
  bgrad = t ? -omp : p;
}

__device__ __forceinline__ float4 encode_deltas_loss(float4 s, float4 t, float wx, float wy, float ww, float wh) {
  float sw = __fsub_rn(s.z, s.x), sh = __fsub_rn(s.w, s.y);
  float sx = __fadd_rn(s.x, __fmul_rn(0.5f, sw)), sy = __fadd_rn(s.y, __fmul_rn(0.5f, sh));
  float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
  float tx = __fadd_rn(t.x, __fmul_rn(0.5f, tw)), ty = __fadd_rn(t.y, __fmul_rn(0.5f, th));
  float4 d;
  d.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tx, sx)), sw);
  d.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(ty, sy)), sh);
  d.z = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
  d.w = __fmul_rn(wh, logf(__fdiv_rn(th, sh)));
  return d;
}

__device__ __forceinline__ void smooth_l1_elem(float pd, float gd, float beta, float& loss, float& grad) {
  const float d = pd - gd;
  const float n = fabsf(d);
  if (beta < 1e-5f) {
    loss = n;
    grad = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  } else if (n < beta) {
    loss = 0.5f * n * n / beta;
    grad = d / beta;
  } else {
    loss = n - 0.5f * beta;
    grad = (d > 0.f) ? 1.f : -1.f;
  }
}

// ---- peer exchange polled by the consumer ------------------------------------------------------------
// Mailbox of a rank (1 KiB): 2 (epoch parity) x 8 (sender) slots of {double v0, double v1, u64 flag, u64 fast word}
// in the first 512 bytes; bytes 512..519 are scratch of the owning rank: epoch << 32 | global num_foreground, published
// by the first CTA that has collected every peer's fast word.
//   fast word  = epoch << 32 | num_foreground (one atomic 8-byte store, posted by K1's fold kernel): all the loss
//                kernel waits for before it can scale a gradient
//   full record {v0 = num_foreground, v1 = S_batch, flag = epoch}: posted by CTA (0,0) of the loss kernel, read by its
//                last CTA for the step's statistics (stats[0..1])
struct PeerPoll {
  const double* mailbox;              // this rank's own mailbox
  const unsigned long long* epoch;    // written by K1's last CTA: the epoch of this step
  int* error;
  double* stats_out;                  // stats[0..1] receive the global sums (written once, by the last CTA)
  long long timeout_cycles;
  double* peer_mailbox[8];            // every rank's mailbox as mapped here (for the full record)
  int world, rank;
};
inline PeerPoll make_peer_poll(const fsg_peer_ctx* h, double* stats) {
  PeerPoll p = {};
  p.world = 1;
  if (h && h->world > 1) {
    p.mailbox = reinterpret_cast<const double*>(h->mailbox[h->rank]);
    p.epoch = reinterpret_cast<const unsigned long long*>(h->epoch);
    p.error = reinterpret_cast<int*>(h->error);
    p.stats_out = stats;
    p.timeout_cycles = h->timeout_cycles > 0 ? (long long)h->timeout_cycles : 120000000000ll;   // ~60 s
    p.world = h->world;
    p.rank = h->rank;
    for (int q = 0; q < h->world; ++q) p.peer_mailbox[q] = reinterpret_cast<double*>(h->mailbox[q]);
  }
  return p;
}
// All threads of the CTA call this (it contains a barrier).  nf comes back as the sum of num_foreground over the
// ranks (integers: the same on every rank and in every CTA).  A peer that does not arrive within the time-out poisons
// the result with NaN (the step's losses and gradients become NaN: it cannot be used silently) and raises *error.
// Only the CTAs that start before the peers have arrived (the first wave) pay for the system-scope poll: the first
// one through publishes the sum in the scratch half of this rank's own mailbox, and every later CTA takes it from
// there with one device-scope acquire.
__device__ __forceinline__ void peer_poll_nf(const PeerPoll& P, double& nf) {
  __shared__ double s_p0[8];
  const int tid = threadIdx.x;
  unsigned long long* pub = reinterpret_cast<unsigned long long*>(const_cast<double*>(P.mailbox)) + 64;
  // fast path (every CTA after the first wave): two independent loads per thread, one barrier.  The published word
  // carries its own validity (epoch << 32 | sum; 0xffffffff = poisoned), so no acquire / dependent second load.
  const unsigned long long ep = *reinterpret_cast<const volatile unsigned long long*>(P.epoch);
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(pub) : "memory");
  const bool fast = (w >> 32) == (ep & 0xffffffffull);
  if (!__syncthreads_or(fast ? 0 : 1)) {   // (CTA-uniform: warps may have read the word at different times)
    const unsigned v = (unsigned)(w & 0xffffffffull);
    nf = (v == 0xffffffffu) ? __longlong_as_double(0x7ff8000000000000ll) : (double)v;
    return;
  }
  if (tid < P.world) {
    const unsigned long long* src =
        reinterpret_cast<const unsigned long long*>(P.mailbox) + ((int)(ep & 1ull) * 8 + tid) * 4 + 3;
    const long long t0 = clock64();
    unsigned long long word = 0ull;
    bool ok = true;
    for (;;) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
      if ((word >> 32) == (ep & 0xffffffffull)) break;
      if (clock64() - t0 > P.timeout_cycles) { ok = false; break; }
    }
    if (ok) {
      s_p0[tid] = (double)(unsigned)(word & 0xffffffffull);
    } else {
      *P.error = 1;
      s_p0[tid] = __longlong_as_double(0x7ff8000000000000ll);
    }
  }
  __syncthreads();
  double a = 0.0;
  for (int p = 0; p < P.world; ++p) a += s_p0[p];
  nf = a;
  if (tid == 0) {   // (several CTAs may publish at once: the same word; a poisoned result is published as well, so
                    //  that the rest of the grid does not wait for the time-out again)
    const unsigned long long lo = (a == a) ? (unsigned long long)(unsigned)(long long)a : 0xffffffffull;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(pub), "l"((ep << 32) | lo) : "memory");
  }
}
// One CTA of the loss kernel (the caller picks it) posts this rank's complete record to every peer; threads
// 0..world-1 each serve one peer.  Not on anybody's critical path: the record is read at the END of the peers' kernels.
__device__ __forceinline__ void peer_post_full(const PeerPoll& P, double v0, double v1) {
  const int tid = threadIdx.x;
  if (tid < P.world) {
    const unsigned long long ep = *reinterpret_cast<const volatile unsigned long long*>(P.epoch);
    double* dst = P.peer_mailbox[tid] + ((int)(ep & 1ull) * 8 + P.rank) * 4;
    dst[0] = v0;
    dst[1] = v1;
    __threadfence_system();
    unsigned long long* fl = reinterpret_cast<unsigned long long*>(dst + 2);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(ep) : "memory");
  }
}
// The last CTA's thread 0: the sum of v1 (S_batch) over the ranks' complete records, in rank order.
__device__ __forceinline__ double peer_sum_v1(const PeerPoll& P) {
  const unsigned long long ep = *reinterpret_cast<const volatile unsigned long long*>(P.epoch);
  double sb = 0.0;
  for (int p = 0; p < P.world; ++p) {
    const double* src = P.mailbox + ((int)(ep & 1ull) * 8 + p) * 4;
    const unsigned long long* fin = reinterpret_cast<const unsigned long long*>(src + 2);
    const long long t0 = clock64();
    unsigned long long seen = 0ull;
    bool ok = true;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fin) : "memory");
      if (seen == ep) break;
      if (clock64() - t0 > P.timeout_cycles) { ok = false; break; }
    }
    if (!ok) *P.error = 1;
    sb += ok ? *reinterpret_cast<const volatile double*>(src + 1) : __longlong_as_double(0x7ff8000000000000ll);
  }
  return sb;
}

// ---- per-tile partials -> scalars --------------------------------------------------------------
// Called by every thread of an NT-thread CTA after it has reduced its own five sums into registers of
// warp lane 0 (acc_* already warp-reduced).  Writes the tile's slot, elects the last CTA of the grid and lets it
// fold all slots in a fixed order (run-to-run deterministic), producing the scalars of include/fsg_dense.h.
template <int NT = kLossBlock, bool PEER = false>
__device__ __forceinline__ void finish_tile(float acc_cls, float acc_reg, float acc_wl, float acc_l, float max_l,
                                            int n, int tile, int T, int N, float* partials, unsigned* counter,
                                            double* scalars, double nf_d, float c_cls, float c_reg, float c_gam,
                                            const PeerPoll peer = PeerPoll{}) {
  __shared__ float s_part[NT / 32][5];
  __shared__ double s_tot[NT / 32][5];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (lane == 0) {
    s_part[wid][0] = acc_cls; s_part[wid][1] = acc_reg; s_part[wid][2] = acc_wl;
    s_part[wid][3] = acc_l; s_part[wid][4] = max_l;
  }
  __syncthreads();
  if (tid == 0) {
    float c = 0.f, g = 0.f, w = 0.f, l = 0.f, mx = 0.f;
    for (int k = 0; k < NT / 32; ++k) {
      c += s_part[k][0]; g += s_part[k][1]; w += s_part[k][2]; l += s_part[k][3];
      mx = fmaxf(mx, s_part[k][4]);
    }
    float* P = partials + ((int64_t)n * T + tile) * kPartialStride;
    P[0] = c; P[1] = g; P[2] = w; P[3] = l; P[4] = mx;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == (unsigned)(T * N) - 1u);
  }
  __syncthreads();
  if (!s_last) return;

  __threadfence();
  double t_cls = 0.0, t_reg = 0.0, t_wl = 0.0, t_l = 0.0, t_mx = 0.0;
  for (int img = wid; img < N; img += NT / 32) {
    double c = 0.0, g = 0.0, w = 0.0, l = 0.0;
    float mx = 0.f;
    for (int b = lane; b < T; b += 32) {
      const float* P = partials + ((int64_t)img * T + b) * kPartialStride;
      c += (double)__ldcg(P + 0); g += (double)__ldcg(P + 1); w += (double)__ldcg(P + 2);
      l += (double)__ldcg(P + 3); mx = fmaxf(mx, __ldcg(P + 4));
    }
    c = warp_sum_d(c); g = warp_sum_d(g); w = warp_sum_d(w); l = warp_sum_d(l); mx = warp_max(mx);
    if (lane == 0) {
      scalars[FSG_SCALARS_HEADER + img] = w;
      t_cls += c; t_reg += g; t_wl += w; t_l += l; t_mx += (double)mx;
    }
  }
  if (lane == 0) {
    s_tot[wid][0] = t_cls; s_tot[wid][1] = t_reg; s_tot[wid][2] = t_wl; s_tot[wid][3] = t_l; s_tot[wid][4] = t_mx;
  }
  __syncthreads();
  if (tid == 0) {
    double v[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < NT / 32; ++k)
      for (int q = 0; q < 5; ++q) v[q] += s_tot[k][q];
    const double nfc = nf_d > 1.0 ? nf_d : 1.0;
    for (int q = 0; q < 5; ++q) scalars[q] = v[q];
    scalars[5] = v[0] / nfc;
    scalars[6] = v[1] / nfc;
    scalars[7] = -v[2];
    scalars[8] = (double)c_cls * scalars[5] + (double)c_reg * scalars[6] + (double)c_gam * scalars[7];
    scalars[9] = nf_d;
    if (PEER && peer.world > 1) {   // the global sums become stats[0..1] (the peers posted their records long ago)
      peer.stats_out[1] = peer_sum_v1(peer);
      peer.stats_out[0] = nf_d;
    }
    *counter = 0u;
  }
}

}  // namespace fsg
