// Bet / weight statistics of GANTrainer.calc_log_metrics (ImbalanceDetection/train_net.py:1104-1121), on the device.
//
// The reference walks the gambler's per-level betting maps -- AFTER gambler_loss multiplied the picky mask into them
// in place (gambler_heads.py:568-569) -- with a host sync per level (`if torch.max(b) > max_bets_all_layers`, :1109)
// and takes torch.median over the N*R normalised weights (:1120, a full sort).  Here:
//   pass 0   one sweep over (bet, mask): sum and max of bet*mask, sum and max of w_hat, and the histogram of the top
//            11 bits of w_hat's bit pattern (w_hat > 0, so bit patterns order like the values)
//   pass 1-2 two more sweeps narrow the median's bin (11 + 10 bits); w_hat is recomputed from (bet, mask, S[n]) with
//            exactly the arithmetic of the loss kernel, never materialised
// Sums go through one partial per CTA in a fixed slot, folded in a fixed order by the last CTA (run-to-run
// deterministic); histogram counts are integers.  torch.median returns the LOWER median: the element of rank
// (n-1)/2 (0-based) of the sorted values.
#include "common.cuh"

namespace fsg {

constexpr int kStatThreads = 256;
constexpr int kStatBins = 2048;

struct StatLevels {
  const float* ptr[FSG_MAX_LEVELS];
  int64_t off[FSG_MAX_LEVELS + 1];
  int HW[FSG_MAX_LEVELS];
  int A, num_levels;
};

struct StatArgs {
  const float* bets;        // flat (N,R) or NULL (then lv)
  const int64_t* mask;      // (N,R) or NULL (mask == 1)
  const double* stats;      // [2+N] of the step (S_batch, S[n])
  int N;
  int64_t R;
  float T;
  int nmode;
  unsigned* hist;           // 3 * kStatBins, zero-initialised
  unsigned* sel;            // [0] prefix bits found so far, [1] rank still wanted inside the prefix
  double* part;             // (gridDim.x, 4) partial sums / maxima of pass 0
  unsigned* done;           // 3 counters, zero-initialised
  double* out;              // 8 doubles, see fsg_bet_stats
};

__device__ __forceinline__ float stat_bet(const StatArgs& A, const StatLevels& lv, int n, int64_t r) {
  if (A.bets) return A.bets[(int64_t)n * A.R + r];
  int l = 0;
  while (l + 1 < lv.num_levels && r >= lv.off[l + 1]) ++l;
  const int local = (int)(r - lv.off[l]);
  const int hw = local / lv.A, a = local - hw * lv.A;
  return lv.ptr[l][((int64_t)n * lv.A + a) * lv.HW[l] + hw];
}

// pass = 0, 1, 2: digit widths 11, 11, 10 of the 32-bit pattern
template <int PASS>
__global__ void __launch_bounds__(kStatThreads) bet_stats_kernel(const StatArgs A, const StatLevels lv) {
  __shared__ unsigned s_hist[kStatBins];
  __shared__ double s_red[kStatThreads / 32][4];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int b = tid; b < kStatBins; b += kStatThreads) s_hist[b] = 0u;
  __syncthreads();
  constexpr int kShift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  constexpr unsigned kMaskBits = PASS == 2 ? 1023u : 2047u;
  const unsigned prefix = PASS == 0 ? 0u : A.sel[0];
  const int64_t total = (int64_t)A.N * A.R;
  double sum_b = 0.0, sum_w = 0.0;
  float max_b = -__int_as_float(0x7f800000), max_w = 0.f;
  // (warp-uniform trip count: the histogram update uses full-mask warp votes)
  for (int64_t i0 = (int64_t)blockIdx.x * kStatThreads; i0 < total; i0 += (int64_t)gridDim.x * kStatThreads) {
    const int64_t i = i0 + tid;
    const bool live = i < total;
    unsigned bits = 0u;
    if (live) {
      const int n = (int)(i / A.R);
      const int64_t r = i - (int64_t)n * A.R;
      const float m = A.mask ? (float)A.mask[i] : 1.f;
      const float bm = __fmul_rn(stat_bet(A, lv, n, r), m);              // gambler_heads.py:569
      float inv_S = 1.f;
      if (A.nmode == FSG_NORM_IMAGE) inv_S = __frcp_rn((float)A.stats[FSG_STATS_HEADER + n]);
      else if (A.nmode == FSG_NORM_BATCH) inv_S = __frcp_rn((float)A.stats[1]);
      const float w_hat = __fadd_rn(bm, A.T) * inv_S;                    // :304, :308-311 (as the loss kernel computes it)
      bits = __float_as_uint(w_hat);
      if (PASS == 0) {
        sum_b += (double)bm; sum_w += (double)w_hat;
        max_b = fmaxf(max_b, bm); max_w = fmaxf(max_w, w_hat);
      }
    }
    if (PASS == 0) {
      warp_hist_add(s_hist, bits >> kShift, live);
    } else {
      const bool in = live && (bits >> (kShift + (PASS == 1 ? 11 : 10))) == prefix;
      warp_hist_add(s_hist, (bits >> kShift) & kMaskBits, in);
    }
  }
  __syncthreads();
  unsigned* gh = A.hist + PASS * kStatBins;
  for (int b = tid; b < kStatBins; b += kStatThreads)
    if (s_hist[b]) atomicAdd(&gh[b], s_hist[b]);
  if (PASS == 0) {
    sum_b = warp_sum_d(sum_b); sum_w = warp_sum_d(sum_w);
    max_b = warp_max(max_b); max_w = warp_max(max_w);
    if (lane == 0) { s_red[wid][0] = sum_b; s_red[wid][1] = sum_w; s_red[wid][2] = max_b; s_red[wid][3] = max_w; }
  }
  __syncthreads();
  if (tid == 0) {
    if (PASS == 0) {
      double sb = 0.0, sw = 0.0, mb = s_red[0][2], mw = 0.0;
      for (int w = 0; w < kStatThreads / 32; ++w) {
        sb += s_red[w][0]; sw += s_red[w][1];
        mb = s_red[w][2] > mb ? s_red[w][2] : mb; mw = s_red[w][3] > mw ? s_red[w][3] : mw;
      }
      double* P = A.part + (int64_t)blockIdx.x * 4;
      P[0] = sb; P[1] = sw; P[2] = mb; P[3] = mw;
    }
    __threadfence();
    s_last = (atomicAdd(&A.done[PASS], 1u) == gridDim.x - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: fold the partial sums (pass 0) and locate the median's digit
  if (PASS == 0 && tid == 0) {
    double sb = 0.0, sw = 0.0, mb = -1e300, mw = 0.0;
    for (unsigned c = 0; c < gridDim.x; ++c) {
      const double* P = A.part + (int64_t)c * 4;
      sb += __ldcg(P); sw += __ldcg(P + 1);
      const double b2 = __ldcg(P + 2), w2 = __ldcg(P + 3);
      mb = b2 > mb ? b2 : mb; mw = w2 > mw ? w2 : mw;
    }
    // train_net.py:1106-1107: max_bets_all_layers starts at 0 and is only replaced by a larger level maximum
    if (!(mb > 0.0)) mb = 0.0;
    A.out[0] = sb; A.out[1] = mb; A.out[2] = sb / (double)total;
    A.out[3] = sw; A.out[4] = mw; A.out[5] = sw / (double)total;
  }
  if (tid == 0) {
    // ascending walk over the bins until the wanted rank falls inside one (positive floats: bits order like values)
    unsigned want = PASS == 0 ? (unsigned)((total - 1) / 2) : A.sel[1];
    unsigned acc = 0u;
    int bin = 0;
    for (; bin < kStatBins; ++bin) {
      const unsigned h = __ldcg(&gh[bin]);
      if (acc + h > want) break;
      acc += h;
    }
    const int width = PASS == 2 ? 10 : 11;
    const unsigned np = (prefix << width) | (unsigned)bin;
    A.sel[0] = np;
    A.sel[1] = want - acc;
    if (PASS == 2) A.out[6] = (double)__uint_as_float(np);   // all 32 bits fixed: the median itself
    A.done[PASS] = 0u;
  }
}

struct StatWs {
  size_t off_hist, off_done, off_sel, off_zero_end, off_part, total;
};
static StatWs stat_ws_layout(int grid) {
  StatWs w;
  size_t o = 0;
  w.off_hist = o; o += sizeof(unsigned) * 3 * kStatBins;
  w.off_done = o; o += 16;
  w.off_sel = o;  o += 16;
  w.off_zero_end = o;
  w.off_part = o; o += sizeof(double) * 4 * (size_t)grid;
  w.total = o;
  return w;
}
constexpr int kStatGrid = 148 * 4;

}  // namespace fsg

using namespace fsg;

extern "C" size_t fsg_bet_stats_workspace_bytes(void) { return stat_ws_layout(kStatGrid).total; }

extern "C" int fsg_bet_stats(const float* bets, const fsg_bet_levels* h_bet_levels, const int64_t* mask, int N,
                             int64_t R, const fsg_loss_params* hp, const double* stats, double* out, void* workspace,
                             size_t workspace_bytes, fsg_stream_t stream) {
  if (!hp || N <= 0 || R <= 0 || !stats || !out) return FSG_ERR_INVALID_ARG;
  if ((bets != nullptr) == (h_bet_levels != nullptr)) return FSG_ERR_INVALID_ARG;   // exactly one of the two
  if ((int64_t)N * R >= ((int64_t)1 << 32)) return FSG_ERR_UNSUPPORTED;
  StatLevels lv = {};
  if (h_bet_levels) {
    if (h_bet_levels->num_levels <= 0 || h_bet_levels->num_levels > FSG_MAX_LEVELS || h_bet_levels->A <= 0)
      return FSG_ERR_INVALID_ARG;
    int64_t off = 0;
    lv.A = h_bet_levels->A;
    lv.num_levels = h_bet_levels->num_levels;
    for (int l = 0; l < FSG_MAX_LEVELS; ++l) {
      lv.off[l] = off;
      if (l < lv.num_levels) {
        const int64_t hw = (int64_t)h_bet_levels->H[l] * h_bet_levels->W[l];
        if (hw < 0 || hw > (1 << 30) || (hw > 0 && !h_bet_levels->bets[l])) return FSG_ERR_INVALID_ARG;
        lv.ptr[l] = h_bet_levels->bets[l];
        lv.HW[l] = (int)hw;
        off += hw * lv.A;
      }
    }
    lv.off[FSG_MAX_LEVELS] = off;
    if (off != R) return FSG_ERR_INVALID_ARG;
  }
  const StatWs w = stat_ws_layout(kStatGrid);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return FSG_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  FSG_CUDA_TRY(cudaMemsetAsync(ws, 0, w.off_zero_end, s));
  StatArgs a;
  a.bets = bets; a.mask = mask; a.stats = stats; a.N = N; a.R = R; a.T = hp->temperature; a.nmode = hp->norm_mode;
  a.hist = (unsigned*)(ws + w.off_hist); a.sel = (unsigned*)(ws + w.off_sel); a.part = (double*)(ws + w.off_part);
  a.done = (unsigned*)(ws + w.off_done); a.out = out;
  int64_t grid = ceil_div((int64_t)N * R, kStatThreads * 8);
  if (grid > kStatGrid) grid = kStatGrid;
  if (grid < 1) grid = 1;
  bet_stats_kernel<0><<<(unsigned)grid, kStatThreads, 0, s>>>(a, lv);
  FSG_LAUNCH_CHECK();
  bet_stats_kernel<1><<<(unsigned)grid, kStatThreads, 0, s>>>(a, lv);
  FSG_LAUNCH_CHECK();
  bet_stats_kernel<2><<<(unsigned)grid, kStatThreads, 0, s>>>(a, lv);
  FSG_LAUNCH_CHECK();
  return FSG_OK;
}
