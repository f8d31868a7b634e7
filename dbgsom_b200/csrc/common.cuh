// Shared helpers for the dbgsom_b200 kernels (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdlib>

#include "../../include/dbgsom_b200.h"

#define DBGSOM_CUDA_TRY(expr)                  \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

#define DBGSOM_LAUNCH_CHECK()                  \
  do {                                         \
    cudaError_t _e = cudaGetLastError();       \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

namespace dbgsom {

constexpr int kMaxCand = DBGSOM_MAX_CAND;
constexpr unsigned kFullMask = 0xffffffffu;

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) {
  return ceil_div(a, b) * b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

// Sum each of U per-lane values over the warp with a transposed butterfly (the number of live
// values halves while the lane distance halves): 6 double shuffles for U = 4 instead of 20.
// Afterwards lane l holds the total of row row_of_lane(l); every row is held by 32 / U lanes.
template <int U>
__device__ __forceinline__ int row_of_lane(int lane) {
  return lane / (32 / U);
}
template <int U>
__device__ __forceinline__ int lane_of_row(int u) {
  return u * (32 / U);
}
template <int U, int OFF>
__device__ __forceinline__ double reduce_rows_step(const double (&v)[U], int lane) {
  if constexpr (U == 1) {
    double b = v[0];
#pragma unroll
    for (int o = OFF; o > 0; o >>= 1) b += __shfl_xor_sync(kFullMask, b, o);
    return b;
  } else {
    const bool up = lane & OFF;
    double a[U / 2];
#pragma unroll
    for (int i = 0; i < U / 2; ++i)
      a[i] = (up ? v[U / 2 + i] : v[i]) + __shfl_xor_sync(kFullMask, up ? v[i] : v[U / 2 + i], OFF);
    return reduce_rows_step<U / 2, OFF / 2>(a, lane);
  }
}
template <int U>
__device__ __forceinline__ double reduce_rows(const double (&v)[U], int lane) {
  return reduce_rows_step<U, 16>(v, lane);
}

// streaming 128-bit load that does not pollute L1 (data is read exactly once)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Candidate tracking shared by both BMU front ends.
//
// One thread owns one sample row and sees the approximate scores s~_j of all prototypes tile by
// tile (ascending j).  With m = the NB-th smallest score seen so far, a prototype is a candidate
// when s~_j <= accept(m), where accept() widens m by the error bound of the front end (both the
// candidate's and the incumbent's score may be off by the bound).  Per tile the owner first
// folds the tile's smallest scores into m (`observe`, branch-free), then offers the elements
// that pass accept(m) (`offer`).  Updating m BEFORE offering matters: prototypes are stored in
// map order, so on a smooth map the scores fall monotonically towards the best region and a
// per-element running minimum would accept almost every element on the way down.  accept(m)
// only shrinks as m shrinks, so offering against the current value keeps a superset of the
// final set; stale entries are filtered in `finish`.  The kMaxCand smallest accepted entries
// live in a small table in shared memory; `evicted` remembers the best score that did not fit,
// so an entry that would still qualify at the end but was dropped is detected (-> overflow).
//
// Bound model: accept(m) = (sqrt(max(m,0)) * (1 + rel) + abs_d)^2   when SQ_DOMAIN (scores are
// squared distances with a relative error `rel` and a distance-domain slack abs_d), or
// accept(m) = m + abs_s (scores with an absolute error bound abs_s / 2).
// ---------------------------------------------------------------------------------------------
struct CandBound {
  float rel;    // sq_domain: relative slack on the distance
  float abs_d;  // sq_domain: absolute slack on the distance
  float abs_s;  // linear domain: absolute slack on the score
};

template <int NB, bool SQ_DOMAIN>
struct RowTracker {
  float m1, m2, thr, evicted;
  uint32_t n_app;
  CandBound b;

  __device__ __forceinline__ void init(const CandBound& bound) {
    m1 = 3.0e38f;
    m2 = 3.0e38f;
    thr = 3.0e38f;  // finite: padded prototypes carry +inf scores and are never accepted
    evicted = __int_as_float(0x7f800000);
    n_app = 0;
    b = bound;
  }
  __device__ __forceinline__ float accept(float m) const {
    if (m >= 1.0e38f) return 3.0e38f;
    if (SQ_DOMAIN) {
      float r = sqrtf(fmaxf(m, 0.f)) * (1.f + b.rel) + b.abs_d;
      return r * r * (1.f + 4.8e-7f);
    }
    return m + b.abs_s;
  }
  // fold the two smallest scores of a tile (a1 <= a2; a2 unused when NB == 1) into the running minima
  __device__ __forceinline__ void observe(float a1, float a2) {
    if (NB == 2) m2 = fminf(fmaxf(m1, a1), fminf(m2, a2));
    m1 = fminf(m1, a1);
    thr = accept(NB == 1 ? m1 : m2);
  }
  // called only for s <= thr.  The table keeps the kMaxCand smallest (score, index) pairs offered; equal
  // scores are ordered by index so the lowest index among exact duplicates always survives.
  // Deliberately not inlined: it is rare, and inlined copies thrash the instruction cache.
  __device__ __noinline__ void offer(float s, int j, int* tab_idx, float* tab_val) {
    if (n_app < (uint32_t)kMaxCand) {
      tab_idx[n_app] = j;
      tab_val[n_app] = s;
    } else {
      int worst = 0;
      float wv = tab_val[0];
      int wj = tab_idx[0];
#pragma unroll
      for (int q = 1; q < kMaxCand; ++q) {
        const float v = tab_val[q];
        const int jq = tab_idx[q];
        if (v > wv || (v == wv && jq > wj)) {
          wv = v;
          wj = jq;
          worst = q;
        }
      }
      if (s < wv || (s == wv && j < wj)) {
        evicted = fminf(evicted, wv);
        tab_idx[worst] = j;
        tab_val[worst] = s;
      } else {
        evicted = fminf(evicted, s);
      }
    }
    ++n_app;
  }
  // final: compact valid candidates to out_idx[0..count) ; returns count or DBGSOM_CAND_OVERFLOW
  __device__ __forceinline__ int finish(const int* tab_idx, const float* tab_val, int* out_idx,
                                        int* best_idx) const {
    const int have = n_app < (uint32_t)kMaxCand ? (int)n_app : kMaxCand;
    int cnt = 0, bi = -1;
    float bv = __int_as_float(0x7f800000);
    for (int q = 0; q < have; ++q) {
      const float v = tab_val[q];
      const int j = tab_idx[q];
      if (v <= thr) {
        out_idx[cnt++] = j;
        if (v < bv || (v == bv && j < bi)) {
          bv = v;
          bi = j;
        }
      }
    }
    *best_idx = bi;
    if (evicted <= thr) return DBGSOM_CAND_OVERFLOW;
    return cnt;
  }
};

// Bound on |approximate - exact score| of the tensor front end for one sample, in the scaled score
// units of the shadows:  coef * ||x'|| * max||u||  for the input rounding, plus fp32 accumulation
// and bias rounding at the 2^-22 level.  Worst case (Cauchy-Schwarz): coef = 2^-9 for one pass,
// 1.5 * 2^-20 for three.  The defaults sit below that on purpose -- rounding errors of 2D terms do not
// align -- with measured head room (tools/calibrate_bound.py, winners against the fp32 path on real
// training trajectories): one pass  0.25   * 2^-9   (first wrong winners appear at ~0.03 * 2^-9),
//                         three     0.0625 * 2^-19  (none seen down to 0.004 * 2^-19).
// bound_scale = 1 selects the worst-case coefficient.
// The accumulation term grows with the length of the fp32 accumulation chain in tensor memory (one step per MMA
// of 16 products): 2.4e-7 was calibrated at D = 256, three passes (48 steps).  Measured with
// tools/check_backends.py at D = 4096 (768 steps; 150k rows x 1024 prototypes x 3 epochs, winners against the fp32
// SIMT search and exact float64 scores): 9 wrong winners with exact relative gaps up to 8e-6 at 1x, 5 at 2x, none
// at 4x and beyond.  Default: scaled with the SQUARE ROOT of the chain length (4x at D = 4096) -- the largest
// factor at which flagged rows can still be proven near-ties; from ~6x every flagged row of a collapsed map goes
// to the float64 re-scan (1 s per epoch at 200k x 4096 x 16384).  strict != 0 scales LINEARLY (16x at D = 4096): a
// 4x margin over the smallest factor without wrong winners, at that price.  The streamed CTA-pair form avoids
// all this by cutting the chain (segmented accumulation, bmu_tc.cu: tensor_acc_coef_args); this function serves
// the one-chain forms.
__device__ __forceinline__ float tensor_score_bound(float xnorm, const float* __restrict__ wmax, float coef, float acc_coef) {
  const float xw = xnorm * wmax[0];
  return xw * coef + acc_coef * (xw + wmax[2]);
}
inline float tensor_acc_coef(int n_pass, int64_t ld16, int strict) {
  const double steps = (double)(ld16 / 16) * (n_pass == 3 ? 3.0 : 1.0);
  double c = 2.4e-7 * (steps > 48.0 ? (strict ? steps / 48.0 : sqrt(steps / 48.0)) : 1.0);
  if (const char* e = getenv("DBGSOM_ACC_SCALE")) c *= atof(e);  // calibration switch
  return (float)c;
}
// accumulation coefficient of the kernel form that runs for these arguments (bmu_tc.cu)
float tensor_acc_coef_args(const dbgsom_bmu_args& a);
// true if the candidate search runs with per-tile error bounds for these arguments (bmu_tc.cu): a flagged row's first
// candidate slot then holds a proven bound on the exact best / second-best gap instead of the raw score gap
bool tile_bounds_active(const dbgsom_bmu_args& a);
__host__ __device__ inline float tensor_bound_coef(int n_pass, float bound_scale) {
  if (n_pass == 1) return (bound_scale > 0.f ? bound_scale : 0.25f) * 1.953125e-3f;
  return (bound_scale > 0.f ? bound_scale : 0.0625f) * 1.9073486e-6f;
}

// workspace layout of the BMU search:
//   [cand_idx int32 N*kMaxCand][cand_count uint8 N (padded)][rescan_count int32 (padded)][rescan_rows int32 N]
struct BmuWorkspace {
  int32_t* cand_idx;
  uint8_t* cand_count;
  int32_t* rescan_count;
  int32_t* rescan_rows;
  __host__ static size_t bytes(int64_t N) {
    return round_up<size_t>((size_t)N * kMaxCand * sizeof(int32_t), 256) + round_up<size_t>((size_t)N, 256) + 256 +
           round_up<size_t>((size_t)N * sizeof(int32_t), 256);
  }
  __host__ static BmuWorkspace carve(void* base, int64_t N) {
    BmuWorkspace w;
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    w.cand_idx = reinterpret_cast<int32_t*>(p);
    p += round_up<size_t>((size_t)N * kMaxCand * sizeof(int32_t), 256);
    w.cand_count = p;
    p += round_up<size_t>((size_t)N, 256);
    w.rescan_count = reinterpret_cast<int32_t*>(p);
    p += 256;
    w.rescan_rows = reinterpret_cast<int32_t*>(p);
    return w;
  }
};

// internal launchers (defined in the .cu files, called from capi.cu)
int launch_bmu_cand_simt(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s);
int launch_bmu_cand_tensor(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s);
int launch_bmu_resolve(const dbgsom_bmu_args& a, const BmuWorkspace& ws, cudaStream_t s);

}  // namespace dbgsom
