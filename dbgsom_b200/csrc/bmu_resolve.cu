// K1 back end: exact float64 re-score of the candidates kept by either front end.
//
// One warp per sample.  For each candidate prototype j the squared distance
// sum_d (x_id - w_jd)^2 is evaluated in float64 from the float32 sample and the float64 master
// prototype, so the winner equals the reference's float64 search (sklearn ArgKmin64 behind
// BaseSom._get_winning_neurons, dbgsom/BaseSom.py:455-457) wherever best and runner-up differ by
// more than rounding; exact ties go to the lowest index like sklearn's heap
// (sklearn/utils/_heap.pyx:46).  Samples whose candidate table overflowed are scored against
// all M prototypes.
#include "common.cuh"

namespace dbgsom {

namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ double sqdist_f64(const float* __restrict__ x, const double* __restrict__ w, int D,
                                             int lane) {
  double acc = 0.0;
  // D % 4 == 0 and 16-byte aligned rows are guaranteed by the C API
  for (int d = lane * 4; d < D; d += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(x + d);
    const double2 w0 = *reinterpret_cast<const double2*>(w + d);
    const double2 w1 = *reinterpret_cast<const double2*>(w + d + 2);
    const double a = (double)xv.x - w0.x, b = (double)xv.y - w0.y;
    const double c = (double)xv.z - w1.x, e = (double)xv.w - w1.y;
    acc = fma(a, a, acc);
    acc = fma(b, b, acc);
    acc = fma(c, c, acc);
    acc = fma(e, e, acc);
  }
  return warp_sum(acc);
}

struct Top2 {
  double d1, d2;
  int i1, i2;
  __device__ __forceinline__ void init() {
    d1 = d2 = __longlong_as_double(0x7ff0000000000000LL);
    i1 = i2 = -1;
  }
  // strict ordering by (distance, index): reproduces "lowest index wins ties"
  __device__ __forceinline__ static bool before(double da, int ia, double db, int ib) {
    return da < db || (da == db && (unsigned)ia < (unsigned)ib);
  }
  __device__ __forceinline__ void offer(double d, int j) {
    if (before(d, j, d1, i1)) {
      d2 = d1; i2 = i1; d1 = d; i1 = j;
    } else if (before(d, j, d2, i2)) {
      d2 = d; i2 = j;
    }
  }
};

template <int NB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) bmu_resolve_kernel(
    const float* __restrict__ X, int64_t N, int D, int64_t ldx, const double* __restrict__ W, int M,
    const int32_t* __restrict__ cand_idx, const uint8_t* __restrict__ cand_count, int want_dist,
    int32_t* __restrict__ idx_out, double* __restrict__ dist_out, unsigned long long* __restrict__ stats,
    // flagged-sample policy (tensor back end only; xnorm16 == nullptr disables the shortcut)
    const float* __restrict__ xnorm16, const float* __restrict__ wmax, float bound_coef, float inv_scale2,
    float tie_rel) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  unsigned long long n_amb = 0, n_ovf = 0, n_cand = 0, n_full = 0;

  for (int64_t row = warp0; row < N; row += nwarps) {
    const int cnt = cand_count[row];
    const bool overflow = cnt == DBGSOM_CAND_OVERFLOW;
    if (!overflow) n_cand += cnt;
    // a single candidate for a single winner is already final (written by the front end)
    if (NB == 1 && cnt == 1 && !want_dist) continue;
    const float* x = X + row * ldx;
    Top2 top;
    top.init();
    if (overflow) {
      ++n_ovf;
      bool full = true;
      if (NB == 1 && xnorm16 != nullptr) {
        // More than kMaxCand prototypes lie within 2 * bound of the best approximate score, hence
        // within 4 * bound of the true minimum: the best/second-best gap is below that.  If this is
        // under the tie tolerance the approximate winner is as good as any (see dbgsom_b200.h).
        const int jb = idx_out[row];
        const double db = sqdist_f64(x, W + (int64_t)jb * D, D, lane);
        const float bound = tensor_score_bound(xnorm16[row], wmax, bound_coef) * inv_scale2;
        if (4.0 * (double)bound <= (double)tie_rel * db) {
          top.offer(db, jb);
          full = false;
        }
      }
      if (full) {
        ++n_full;
        top.init();
        for (int j = 0; j < M; ++j) top.offer(sqdist_f64(x, W + (int64_t)j * D, D, lane), j);
      }
    } else {
      if (cnt > NB) ++n_amb;
      const int my = lane < kMaxCand ? cand_idx[row * kMaxCand + lane] : -1;
      for (int q = 0; q < cnt; ++q) {
        const int j = __shfl_sync(kFullMask, my, q);
        top.offer(sqdist_f64(x, W + (int64_t)j * D, D, lane), j);
      }
    }
    if (lane == 0) {
      idx_out[row * NB] = top.i1;
      if (NB == 2) idx_out[row * NB + 1] = top.i2;
      if (want_dist) {
        dist_out[row * NB] = sqrt(top.d1);
        if (NB == 2) dist_out[row * NB + 1] = sqrt(top.d2);
      }
    }
  }
  if (stats != nullptr && lane == 0 && (n_amb | n_ovf | n_cand | n_full)) {
    atomicAdd(stats + 0, n_amb);
    atomicAdd(stats + 1, n_ovf);
    atomicAdd(stats + 2, n_cand);
    atomicAdd(stats + 3, n_full);
  }
}

}  // namespace

int launch_bmu_resolve(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s) {
  int64_t blocks = ceil_div<int64_t>(a.N, WARPS_PER_BLOCK);
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride beyond 16 resident CTAs per SM
  auto* stats = reinterpret_cast<unsigned long long*>(a.d_stats);
  const bool shortcut = a.backend == DBGSOM_BMU_TENSOR && !a.strict;
  const float* xn = shortcut ? a.d_xnorm16 : nullptr;
  const float coef = tensor_bound_coef(a.n_pass, a.bound_scale);
  const float inv_s2 = a.scale > 0.f ? 1.f / (a.scale * a.scale) : 1.f;
  const float tie = a.tie_rel > 0.f ? a.tie_rel : 1e-6f;
  if (a.n_bmu == 1)
    bmu_resolve_kernel<1><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
        a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
        a.d_wmax, coef, inv_s2, tie);
  else
    bmu_resolve_kernel<2><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
        a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
        a.d_wmax, coef, inv_s2, tie);
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace dbgsom
