// K1 front end on CUDA cores: fp32 candidate search by direct differences.
//
// Replaces the distance evaluation inside BaseSom._get_winning_neurons (dbgsom/BaseSom.py:446-464)
// for shapes where the tensor-core path does not pay (small M or N) and serves as the on-device
// cross-check of the tcgen05 front end.  Scores are d~2_ij = sum_d (x_id - w32_jd)^2 accumulated
// in fp32; their relative error is bounded by (D+4) * 2^-24 and the prototypes themselves are
// rounded from float64 (distance changes by at most 2^-24 ||w||), which gives the acceptance
// window of RowTracker<SQ_DOMAIN=true>.
#include "common.cuh"

namespace dbgsom {

namespace {

constexpr int BM = 64;   // sample rows per CTA
constexpr int BN = 64;   // prototypes per column tile
constexpr int BK = 32;   // features per shared-memory stage
constexpr int PAD = 4;   // keeps 16-byte alignment of the transposed tiles
constexpr int THREADS = 256;

template <int NB>
__global__ void __launch_bounds__(THREADS) bmu_cand_simt_kernel(const float* __restrict__ X, int64_t N, int D,
                                                               int64_t ldx, const float* __restrict__ W, int M,
                                                               const float* __restrict__ wmax, int32_t* __restrict__ idx_out,
                                                               int32_t* __restrict__ cand_idx,
                                                               uint8_t* __restrict__ cand_count) {
  __shared__ __align__(16) float Xs[BK][BM + PAD];
  __shared__ __align__(16) float Ws[BK][BN + PAD];
  __shared__ float S[BM][BN + 1];
  __shared__ int ring_idx[BM][kMaxCand];
  __shared__ float ring_val[BM][kMaxCand];

  const int tid = threadIdx.x;
  const int tx = tid % 16;  // 4 prototypes each
  const int ty = tid / 16;  // 4 rows each
  const int64_t row0 = (int64_t)blockIdx.x * BM;

  RowTracker<NB, true> trk;
  if (tid < BM) {
    CandBound b;
    b.rel = (float)(D + 4) * 5.9604645e-8f * 1.05f;  // (D+4) * 2^-24 on the squared distance, i.e.
                                                     // half of it per side on the distance, twice (both scores)
    b.abs_d = 2.1f * 5.9604645e-8f * wmax[1];        // both prototypes rounded to fp32
    b.abs_s = 0.f;
    trk.init(b);
  }

  // loader mapping: thread -> (row/prototype r = tid / 8, feature quad kq = tid % 8)
  const int lr = tid / 8;        // 0..31, two passes cover 64 rows
  const int lk = (tid % 8) * 4;  // 0,4,..,28

  for (int col0 = 0; col0 < M; col0 += BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < D; k0 += BK) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int r = lr + 32 * p;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + lk;
        if (k < D) {  // D % 4 == 0 is guaranteed by the C API
          if (row0 + r < N) xv = *reinterpret_cast<const float4*>(X + (row0 + r) * ldx + k);
          if (col0 + r < M) wv = *reinterpret_cast<const float4*>(W + (int64_t)(col0 + r) * D + k);
        }
        Xs[lk + 0][r] = xv.x; Xs[lk + 1][r] = xv.y; Xs[lk + 2][r] = xv.z; Xs[lk + 3][r] = xv.w;
        Ws[lk + 0][r] = wv.x; Ws[lk + 1][r] = wv.y; Ws[lk + 2][r] = wv.z; Ws[lk + 3][r] = wv.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d = av[i] - bv[j];
            acc[i][j] = fmaf(d, d, acc[i][j]);
          }
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        S[ty * 4 + i][tx * 4 + j] = c < M ? acc[i][j] : __int_as_float(0x7f800000);
      }
    __syncthreads();
    if (tid < BM) {
      float a1 = __int_as_float(0x7f800000), a2 = a1;
      for (int c = 0; c < BN; ++c) {
        const float s = S[tid][c];
        if (NB == 2) a2 = fminf(a2, fmaxf(a1, s));
        a1 = fminf(a1, s);
      }
      trk.observe(a1, a2);
      if (a1 <= trk.thr) {
        for (int c = 0; c < BN; ++c) {
          const float s = S[tid][c];
          if (s <= trk.thr) trk.offer(s, col0 + c, ring_idx[tid], ring_val[tid]);
        }
      }
    }
    __syncthreads();
  }

  if (tid < BM && row0 + tid < N) {
    const int64_t row = row0 + tid;
    int out[kMaxCand];
    int best;
    const int cnt = trk.finish(ring_idx[tid], ring_val[tid], out, &best);
    const int valid = cnt == DBGSOM_CAND_OVERFLOW ? 0 : cnt;
#pragma unroll
    for (int q = 0; q < kMaxCand; ++q) cand_idx[row * kMaxCand + q] = q < valid ? out[q] : -1;
    cand_count[row] = (uint8_t)cnt;
    idx_out[row * NB] = best;  // provisional winner; final for unambiguous rows when NB == 1
  }
}

}  // namespace

int launch_bmu_cand_simt(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s) {
  const dim3 grid((unsigned)ceil_div<int64_t>(a.N, BM));
  if (a.n_bmu == 1)
    bmu_cand_simt_kernel<1><<<grid, THREADS, 0, s>>>(a.d_X, a.N, a.D, a.ldx, a.d_W32, a.M, a.d_wmax, a.d_idx,
                                                    ws.cand_idx, ws.cand_count);
  else
    bmu_cand_simt_kernel<2><<<grid, THREADS, 0, s>>>(a.d_X, a.N, a.D, a.ldx, a.d_W32, a.M, a.d_wmax, a.d_idx,
                                                    ws.cand_idx, ws.cand_count);
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace dbgsom
