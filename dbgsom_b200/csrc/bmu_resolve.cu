// K1 back end: exact float64 re-score of the candidates kept by either front end.
//
// One warp per sample.  For each candidate prototype j the squared distance
// sum_d (x_id - w_jd)^2 is evaluated in float64 from the float32 sample and the float64 master
// prototype, so the winner equals the reference's float64 search (sklearn ArgKmin64 behind
// BaseSom._get_winning_neurons, dbgsom/BaseSom.py:455-457) wherever best and runner-up differ by
// more than rounding; exact ties go to the lowest index like sklearn's heap
// (sklearn/utils/_heap.pyx:46).  Samples whose candidate table overflowed (and whose bound is not
// under the tie tolerance) are queued and scored against all M prototypes by a second kernel
// that shares every prototype row between eight queued samples.
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
    float tie_rel, int32_t* __restrict__ rescan_count, int32_t* __restrict__ rescan_rows) {
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
        if (lane == 0) rescan_rows[atomicAdd(rescan_count, 1)] = (int32_t)row;
        continue;  // written by bmu_rescan_kernel
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

// Full re-score of queued samples: one CTA takes eight samples (in shared memory) and streams all
// M float64 prototype rows once for the eight of them; warps stripe the prototypes, lanes the features.
constexpr int RS_ROWS = 8;
template <int NB>
__global__ void __launch_bounds__(256) bmu_rescan_kernel(const float* __restrict__ X, int64_t ldx, int D,
                                                        const double* __restrict__ W, int M,
                                                        const int32_t* __restrict__ rescan_count,
                                                        const int32_t* __restrict__ rescan_rows, int want_dist,
                                                        int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  extern __shared__ __align__(16) float xs[];  // [RS_ROWS][D]
  __shared__ double m_d[8][RS_ROWS][2];
  __shared__ int m_i[8][RS_ROWS][2];
  __shared__ int row_id[RS_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *rescan_count;
  const int groups = ceil_div(n, RS_ROWS);
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < RS_ROWS) {
      const int q = g * RS_ROWS + threadIdx.x;
      row_id[threadIdx.x] = q < n ? rescan_rows[q] : -1;
    }
    __syncthreads();
    for (int e = threadIdx.x * 4; e < RS_ROWS * D; e += 256 * 4) {
      const int r = e / D, d = e % D;
      const int rid = row_id[r];
      const float4 v = rid >= 0 ? *reinterpret_cast<const float4*>(X + (int64_t)rid * ldx + d)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(xs + e) = v;
    }
    __syncthreads();
    Top2 top[RS_ROWS];
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) top[r].init();
    for (int j = warp; j < M; j += 8) {
      double acc[RS_ROWS];
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r) acc[r] = 0.0;
      const double* w = W + (int64_t)j * D;
      for (int d = lane * 4; d < D; d += 128) {
        const double2 w0 = *reinterpret_cast<const double2*>(w + d);
        const double2 w1 = *reinterpret_cast<const double2*>(w + d + 2);
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + r * D + d);
          const double a = (double)xv.x - w0.x, b = (double)xv.y - w0.y;
          const double c = (double)xv.z - w1.x, e2 = (double)xv.w - w1.y;
          acc[r] = fma(a, a, acc[r]);
          acc[r] = fma(b, b, acc[r]);
          acc[r] = fma(c, c, acc[r]);
          acc[r] = fma(e2, e2, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r) top[r].offer(warp_sum(acc[r]), j);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < RS_ROWS; ++r) {
        m_d[warp][r][0] = top[r].d1; m_d[warp][r][1] = top[r].d2;
        m_i[warp][r][0] = top[r].i1; m_i[warp][r][1] = top[r].i2;
      }
    }
    __syncthreads();
    if (threadIdx.x < RS_ROWS && row_id[threadIdx.x] >= 0) {
      const int r = threadIdx.x;
      Top2 t;
      t.init();
      for (int q = 0; q < 8; ++q) {
        if (m_i[q][r][0] >= 0) t.offer(m_d[q][r][0], m_i[q][r][0]);
        if (m_i[q][r][1] >= 0) t.offer(m_d[q][r][1], m_i[q][r][1]);
      }
      const int64_t row = row_id[r];
      idx_out[row * NB] = t.i1;
      if (NB == 2) idx_out[row * NB + 1] = t.i2;
      if (want_dist) {
        dist_out[row * NB] = sqrt(t.d1);
        if (NB == 2) dist_out[row * NB + 1] = sqrt(t.d2);
      }
    }
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
  DBGSOM_CUDA_TRY(cudaMemsetAsync(ws.rescan_count, 0, sizeof(int32_t), s));
  const size_t rs_smem = (size_t)RS_ROWS * a.D * sizeof(float);
  const int rs_grid = 148 * 2;
  if (a.n_bmu == 1) {
    bmu_resolve_kernel<1><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
        a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
        a.d_wmax, coef, inv_s2, tie, ws.rescan_count, ws.rescan_rows);
    DBGSOM_LAUNCH_CHECK();
    DBGSOM_CUDA_TRY(cudaFuncSetAttribute(bmu_rescan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    bmu_rescan_kernel<1><<<rs_grid, 256, rs_smem, s>>>(a.d_X, a.ldx, a.D, a.d_W, a.M, ws.rescan_count, ws.rescan_rows,
                                                      a.want_dist, a.d_idx, a.d_dist);
  } else {
    bmu_resolve_kernel<2><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
        a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
        a.d_wmax, coef, inv_s2, tie, ws.rescan_count, ws.rescan_rows);
    DBGSOM_LAUNCH_CHECK();
    DBGSOM_CUDA_TRY(cudaFuncSetAttribute(bmu_rescan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    bmu_rescan_kernel<2><<<rs_grid, 256, rs_smem, s>>>(a.d_X, a.ldx, a.D, a.d_W, a.M, ws.rescan_count, ws.rescan_rows,
                                                      a.want_dist, a.d_idx, a.d_dist);
  }
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace dbgsom
