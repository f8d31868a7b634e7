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
    const float* __restrict__ xnorm16, const float* __restrict__ wmax, float bound_coef, float acc_coef, float inv_scale2,
    float tie_rel, int32_t* __restrict__ rescan_count, int32_t* __restrict__ rescan_rows, int gap_is_proven) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  unsigned long long n_amb = 0, n_ovf = 0, n_cand = 0, n_full = 0;

  // 32 rows per warp step: every lane reads one row's candidate count (one coalesced load), then the warp
  // works through the rows that need a re-score together.  In a training epoch > 99 % of the rows have a
  // single candidate, which is already final (written by the front end); one warp per ROW spent most of this
  // kernel's time discovering that.
  for (int64_t base = warp0 * 32; base < N; base += nwarps * 32) {
    const int64_t my_row = base + lane;
    const int my_cnt = my_row < N ? cand_count[my_row] : 1;
    if (my_row < N && my_cnt != DBGSOM_CAND_OVERFLOW) n_cand += my_cnt;
    unsigned todo = __ballot_sync(kFullMask, my_row < N && !(NB == 1 && my_cnt == 1 && !want_dist));
    while (todo) {
      const int src_lane = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t row = base + src_lane;
      const int cnt = __shfl_sync(kFullMask, my_cnt, src_lane);
      const bool overflow = cnt == DBGSOM_CAND_OVERFLOW;
      const float* x = X + row * ldx;
      Top2 top;
      top.init();
      if (overflow) {
        ++n_ovf;
        bool full = true;
        if (NB == 1 && xnorm16 != nullptr) {
          // More than kMaxCand prototypes lie within 2 * bound of the best approximate score s1.  With s2 the
          // second smallest approximate score and B the bound: every exact squared distance is >= s1 - B and
          // the two prototypes behind s1, s2 are exactly <= s2 + B, so the exact best / second-best gap is at
          // most (s2 - s1) + 2B and the exact minimum is at least db - 2B (db = exact distance of the approximate
          // winner).  If that gap is under the tie tolerance the approximate winner is as good as any (see
          // dbgsom_b200.h); the candidate search leaves s2 - s1 in the row's first candidate slot.
          const int jb = idx_out[row];
          const double db = sqdist_f64(x, W + (int64_t)jb * D, D, lane);
          const double bound = (double)(tensor_score_bound(xnorm16[row], wmax, bound_coef, acc_coef) * inv_scale2);
          // (with per-tile bounds -- bmu_tc.cu, TB -- the slot already holds "second smallest upper bound minus
          // smallest lower bound" of the exact scores, a proven bound on the gap by itself)
          const double gap = (double)(__int_as_float(cand_idx[row * kMaxCand]) * inv_scale2) + (gap_is_proven ? 0.0 : 2.0 * bound);
          if (jb >= 0 && gap <= (double)tie_rel * (db - 2.0 * bound)) {
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
  }
  // n_cand was counted per lane (its own rows), the other three per warp (uniform)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_xor_sync(kFullMask, n_cand, o);
  if (stats != nullptr && lane == 0 && (n_amb | n_ovf | n_cand | n_full)) {
    atomicAdd(stats + 0, n_amb);
    atomicAdd(stats + 1, n_ovf);
    atomicAdd(stats + 2, n_cand);
    atomicAdd(stats + 3, n_full);
  }
}

// Full re-score of queued samples: one CTA takes R samples, converts them to float64 in shared memory
// once, and streams all M float64 prototype rows once for the R of them.  A warp scores RS_P prototypes
// per iteration against all R samples (R x RS_P accumulators in registers), so every shared-memory load
// of a sample element feeds RS_P subtract/FMA pairs and every prototype element R of them: the float64
// pipe is the limiter, not the load ports (the first version converted float -> double per use and was
// bound by the conversion unit at ~20 % of the float64 rate).  The R x RS_P partial sums of the 32 lanes
// are summed with one transposed butterfly (31 shuffle steps for 32 values; separate warp sums took 160
// and were a third of the kernel's instructions); lane (r * RS_P + p) * 32 / (R * RS_P) then owns the
// pair (sample r, prototype j0 + p) and keeps its own running top-2.
constexpr int RS_P = 4;
// XT: element type of the staged samples.  double for D <= 3200; float beyond (converted per use on the otherwise
// idle conversion unit), which keeps R = 8 samples per pass over the prototypes within shared memory -- with 16384 x
// 4096 float64 prototypes the pass is bound by L2 bandwidth, so rows per pass is what counts.
template <int NB, int R, typename XT>
__global__ void __launch_bounds__(256) bmu_rescan_kernel(const float* __restrict__ X, int64_t ldx, int D,
                                                        const double* __restrict__ W, int M,
                                                        const int32_t* __restrict__ rescan_count,
                                                        const int32_t* __restrict__ rescan_rows, int want_dist,
                                                        int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  constexpr int V = R * RS_P;   // values per lane; 32 / V lanes end up holding each total
  static_assert(V <= 32 && 32 % V == 0, "R * RS_P must divide the warp");
  extern __shared__ __align__(16) unsigned char xs_raw[];
  XT* xs = reinterpret_cast<XT*>(xs_raw);  // [R][D]
  __shared__ double m_d[8][V][2];
  __shared__ int m_i[8][V][2];
  __shared__ int row_id[R];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int vi = lane / (32 / V);            // the (sample, prototype slot) pair this lane owns after the reduction
  const int my_p = vi % RS_P;
  const bool leader = lane % (32 / V) == 0;
  const int n = *rescan_count;
  const int groups = ceil_div(n, R);
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < R) {
      const int q = g * R + threadIdx.x;
      row_id[threadIdx.x] = q < n ? rescan_rows[q] : -1;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < R * D; e += 256) {
      const int r = e / D, d = e % D;
      const int rid = row_id[r];
      xs[e] = rid >= 0 ? (XT)X[(int64_t)rid * ldx + d] : (XT)0;
    }
    __syncthreads();
    Top2 mine;
    mine.init();
    for (int j0 = warp * RS_P; j0 < M; j0 += 8 * RS_P) {
      double acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.0;
      for (int d = lane * 2; d < D; d += 64) {  // D % 4 == 0: pairs never straddle the row end
        double2 w[RS_P];
#pragma unroll
        for (int p = 0; p < RS_P; ++p) {
          const int j = j0 + p < M ? j0 + p : M - 1;
          w[p] = *reinterpret_cast<const double2*>(W + (int64_t)j * D + d);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          double2 xv;
          if constexpr (sizeof(XT) == 8) {
            xv = *reinterpret_cast<const double2*>(xs + r * D + d);
          } else {
            const float2 xf = *reinterpret_cast<const float2*>(xs + r * D + d);
            xv = make_double2((double)xf.x, (double)xf.y);
          }
#pragma unroll
          for (int p = 0; p < RS_P; ++p) {
            const double a = xv.x - w[p].x, b = xv.y - w[p].y;
            acc[r * RS_P + p] = fma(a, a, acc[r * RS_P + p]);
            acc[r * RS_P + p] = fma(b, b, acc[r * RS_P + p]);
          }
        }
      }
      const double tot = reduce_rows<V>(acc, lane);
      if (leader && j0 + my_p < M) mine.offer(tot, j0 + my_p);
    }
    if (leader) {
      m_d[warp][vi][0] = mine.d1; m_d[warp][vi][1] = mine.d2;
      m_i[warp][vi][0] = mine.i1; m_i[warp][vi][1] = mine.i2;
    }
    __syncthreads();
    if (threadIdx.x < R && row_id[threadIdx.x] >= 0) {
      const int r = threadIdx.x;
      Top2 t;
      t.init();
      for (int q = 0; q < 8; ++q) {
        for (int p = 0; p < RS_P; ++p) {
          if (m_i[q][r * RS_P + p][0] >= 0) t.offer(m_d[q][r * RS_P + p][0], m_i[q][r * RS_P + p][0]);
          if (m_i[q][r * RS_P + p][1] >= 0) t.offer(m_d[q][r * RS_P + p][1], m_i[q][r * RS_P + p][1]);
        }
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

template <int NB, int R, typename XT>
int launch_rescan(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s) {
  const size_t smem = (size_t)R * a.D * sizeof(XT);
  auto kern = bmu_rescan_kernel<NB, R, XT>;
  DBGSOM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  kern<<<148 * per_sm, 256, smem, s>>>(a.d_X, a.ldx, a.D, a.d_W, a.M, ws.rescan_count, ws.rescan_rows, a.want_dist,
                                       a.d_idx, a.d_dist);
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace

int launch_bmu_resolve(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s) {
  int64_t blocks = ceil_div<int64_t>(a.N, WARPS_PER_BLOCK * 32);  // a warp takes 32 rows per step
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride beyond 16 resident CTAs per SM
  auto* stats = reinterpret_cast<unsigned long long*>(a.d_stats);
  const bool shortcut = a.backend == DBGSOM_BMU_TENSOR && !a.strict;
  const float* xn = shortcut ? a.d_xnorm16 : nullptr;
  const float coef = tensor_bound_coef(a.n_pass, a.bound_scale);
  const float acc_coef = tensor_acc_coef_args(a);
  const float inv_s2 = a.scale > 0.f ? 1.f / (a.scale * a.scale) : 1.f;
  const float tie = a.tie_rel > 0.f ? a.tie_rel : 1e-6f;
  DBGSOM_CUDA_TRY(cudaMemsetAsync(ws.rescan_count, 0, sizeof(int32_t), s));
  const bool wide = (size_t)8 * a.D * sizeof(double) > 200 * 1024;  // D > 3200: samples staged as float
  if (a.n_bmu == 1) {
    bmu_resolve_kernel<1><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
        a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
        a.d_wmax, coef, acc_coef, inv_s2, tie, ws.rescan_count, ws.rescan_rows, tile_bounds_active(a) ? 1 : 0);
    DBGSOM_LAUNCH_CHECK();
    return wide ? launch_rescan<1, 8, float>(a, ws, s) : launch_rescan<1, 8, double>(a, ws, s);
  }
  bmu_resolve_kernel<2><<<(unsigned)blocks, WARPS_PER_BLOCK * 32, 0, s>>>(
      a.d_X, a.N, a.D, a.ldx, a.d_W, a.M, ws.cand_idx, ws.cand_count, a.want_dist, a.d_idx, a.d_dist, stats, xn,
      a.d_wmax, coef, acc_coef, inv_s2, tie, ws.rescan_count, ws.rescan_rows, 0);
  DBGSOM_LAUNCH_CHECK();
  return wide ? launch_rescan<2, 8, float>(a, ws, s) : launch_rescan<2, 8, double>(a, ws, s);
}

}  // namespace dbgsom
