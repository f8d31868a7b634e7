// K2: sample weights + per-BMU segmented accumulation (HBM-bound; reads X exactly once).
//
// Replaces, per epoch (dbgsom/BaseSom.py):
//   _calculate_exp_similarity                      :533-538   k_i = 1 - sqrt(1 - exp(-d_i^2 / V))  (float64)
//   np.argsort(winners) / np.unique(return_index)  :488-489   -> counting sort by winner
//   numba_voronoi_set_centers                      :1028-1055 -> Sk_j = sum k_i x_i, sk_j = sum k_i
//   neuron_activations                             :500-503   -> n_j
//   numba_quantization_error                       :1058-1073 -> E_j = sum d_i
//
// Pipeline (all on one stream):
//   histogram (counts per winner [+ class histogram]) -> exclusive scan (segment offsets, n_j as
//   float64) -> scatter (sample permutation grouped by winner) -> segmented accumulate.
// The accumulate kernel walks the permutation in fixed-size chunks; a team of warps owns a
// chunk, keeps the current segment's float64 prototype in registers (so the exact distance
// ||x_i - w_b|| costs no extra memory traffic), accumulates k_i x_i in float64 registers and
// flushes with float64 atomics whenever the segment changes.  Everything after the fp32 load
// of x is float64, so the result is independent of the (atomic-ordered) permutation up to
// 1e-16 and matches the reference to ~1e-13; its convergence test (sum ||dW|| < 1e-5,
// dbgsom/BaseSom.py:519-522) needs that: with fp32 weights or partial sums the prototypes jitter
// by ~1e-7 relative from epoch to epoch and the fit stops at a different epoch (or never).
// B200 has the float64 rate for this (~3 DFMA per sample element, far below the HBM time).  The M x D table is never privatised in
// shared memory (4 MB at 4096 x 256); traffic is N*(4D+8) + a few M*D words.
#include "common.cuh"

namespace dbgsom {

namespace {

// ------------------------------------------------------------------------------------------ histogram
constexpr int HIST_THREADS = 256;
constexpr int HIST_SMEM_BINS = 8192;

// counts[b] (and class_hist[b, y]) via shared-memory privatisation when the bin count is small
// (few bins => heavy contention on global atomics), plain global atomics otherwise.
template <bool SMEM>
__global__ void __launch_bounds__(HIST_THREADS) hist_kernel(const int32_t* __restrict__ bmu, int64_t N, int M,
                                                           const int32_t* __restrict__ labels, int n_classes,
                                                           int32_t* __restrict__ counts,
                                                           int32_t* __restrict__ class_hist, int64_t rows_per_block) {
  extern __shared__ int32_t sh[];
  const int nbins_c = labels ? M * n_classes : 0;
  if (SMEM) {
    for (int i = threadIdx.x; i < M + nbins_c; i += HIST_THREADS) sh[i] = 0;
    __syncthreads();
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += HIST_THREADS) {
    const int b = bmu[i];
    if ((unsigned)b >= (unsigned)M) continue;  // defensive: a NaN row has no winner
    if (SMEM) {
      atomicAdd(&sh[b], 1);
      if (labels) atomicAdd(&sh[M + b * n_classes + labels[i]], 1);
    } else {
      atomicAdd(&counts[b], 1);
      if (labels) atomicAdd(&class_hist[(int64_t)b * n_classes + labels[i]], 1);
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += HIST_THREADS)
      if (sh[i]) atomicAdd(&counts[i], sh[i]);
    for (int i = threadIdx.x; i < nbins_c; i += HIST_THREADS)
      if (sh[M + i]) atomicAdd(&class_hist[i], sh[M + i]);
  }
}

// ------------------------------------------------------------------------------------------ scan
// single CTA: offsets[j] = sum_{i<j} counts[i], offsets[M] = total; n_j -> part_n as float64; cursor = 0
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const int32_t* __restrict__ counts, int M,
                                                           int32_t* __restrict__ offsets,
                                                           int32_t* __restrict__ cursor, double* __restrict__ part_n) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_sh;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < M; base += SCAN_THREADS) {
    const int j = base + threadIdx.x;
    const int32_t c = j < M ? counts[j] : 0;
    int32_t v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(kFullMask, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int32_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(kFullMask, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;  // inclusive totals of warps
    }
    __syncthreads();
    const int32_t carry = carry_sh;
    const int32_t excl = carry + (warp ? warp_tot[warp - 1] : 0) + v - c;
    if (j < M) {
      offsets[j] = excl;
      cursor[j] = 0;
      part_n[j] = (double)c;
    }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry_sh = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[M] = carry_sh;
}

// ------------------------------------------------------------------------------------------ scatter
// perm[offsets[b] + slot] = i, slots handed out per winner; lanes of a warp that share a winner
// take consecutive slots from one atomic (few prototypes => thousands of samples per address).
constexpr int SCAT_THREADS = 256;
__global__ void __launch_bounds__(SCAT_THREADS) scatter_kernel(const int32_t* __restrict__ bmu, int64_t N, int M,
                                                              const int32_t* __restrict__ offsets,
                                                              int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * SCAT_THREADS;
  const int64_t n_round = round_up<int64_t>(N, 32);
  for (int64_t i = (int64_t)blockIdx.x * SCAT_THREADS + threadIdx.x; i < n_round; i += stride) {
    int b = i < N ? bmu[i] : -1;
    if ((unsigned)b >= (unsigned)M) b = -1;
    const unsigned peers = __match_any_sync(kFullMask, b);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1));
    int base = 0;
    if (lane == leader && b >= 0) base = atomicAdd(&cursor[b], __popc(peers));
    base = __shfl_sync(kFullMask, base, leader);
    if (b >= 0) perm[offsets[b] + base + rank] = (int32_t)i;
  }
}

// ------------------------------------------------------------------------------------------ accumulate
constexpr int ACC_THREADS = 256;
constexpr int ACC_CHUNK = 128;  // sorted positions per team task

// VPL : float4 per lane per row slab;  WPR : warps cooperating on one row (team size);
// U   : rows in flight per team (their loads are issued together and their sample weights are
//       evaluated in one go, lane u doing row u, so the float64 exp/sqrt costs 1/U per row).
// A team handles columns [tw * 128 * VPL, (tw + 1) * 128 * VPL) of each row with warp tw.
// All arithmetic is float64: distances by direct differences, k = 1 - sqrt(1 - exp(-d^2 / V))
// literally as dbgsom/BaseSom.py:536-537, sums in float64 registers.
template <int VPL, int WPR, int U>
__global__ void __launch_bounds__(ACC_THREADS) accumulate_kernel(
    const float* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ offsets, const double* __restrict__ W, int M, double inv_var,
    double* __restrict__ part) {
  constexpr int TEAMS = ACC_THREADS / 32 / WPR;
  __shared__ double red[TEAMS][2][WPR][U];  // cross-warp partial squared distances (double buffered)

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int team = warp / WPR;
  const int tw = warp % WPR;
  const int col0 = tw * 128 * VPL + lane * 4;
  const int32_t total = offsets[M];  // samples that have a winner (== N without NaN rows)

  double* __restrict__ Sk = part;
  double* __restrict__ sk = part + (int64_t)M * D;
  double* __restrict__ En = sk + 2 * (int64_t)M;

  const int64_t n_tasks = ceil_div<int64_t>(total, ACC_CHUNK);
  int parity = 0;
  for (int64_t task = (int64_t)blockIdx.x * TEAMS + team; task < n_tasks; task += (int64_t)gridDim.x * TEAMS) {
    const int32_t p0 = (int32_t)(task * ACC_CHUNK);
    const int32_t p1 = p0 + ACC_CHUNK < total ? p0 + ACC_CHUNK : total;
    // segment containing p0: largest j with offsets[j] <= p0
    int lo = 0, hi = M;  // invariant offsets[lo] <= p0 < offsets[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (offsets[mid] <= p0) lo = mid; else hi = mid;
    }
    int seg = lo;
    int32_t seg_end = offsets[seg + 1];

    double w[VPL][4], acc[VPL][4];
    double run_k = 0.0, run_d = 0.0;
    auto load_w = [&]() {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = col0 + v * 128;
        if (c < D) {
          const double2 a = *reinterpret_cast<const double2*>(W + (int64_t)seg * D + c);
          const double2 b = *reinterpret_cast<const double2*>(W + (int64_t)seg * D + c + 2);
          w[v][0] = a.x; w[v][1] = a.y; w[v][2] = b.x; w[v][3] = b.y;
        } else {
          w[v][0] = w[v][1] = w[v][2] = w[v][3] = 0.0;
        }
        acc[v][0] = acc[v][1] = acc[v][2] = acc[v][3] = 0.0;
      }
      run_k = 0.0;
      run_d = 0.0;
    };
    auto flush = [&]() {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = col0 + v * 128;
        if (c < D) {
          double* dst = Sk + (int64_t)seg * D + c;
          atomicAdd(dst + 0, acc[v][0]);
          atomicAdd(dst + 1, acc[v][1]);
          atomicAdd(dst + 2, acc[v][2]);
          atomicAdd(dst + 3, acc[v][3]);
        }
      }
      if (tw == 0 && lane == 0) {
        atomicAdd(sk + seg, run_k);
        atomicAdd(En + seg, run_d);
      }
    };
    load_w();

    int32_t p = p0;
    while (p < p1) {
      if (p >= seg_end) {  // entering a new segment (skipping empty ones)
        flush();
        do {
          ++seg;
          seg_end = offsets[seg + 1];
        } while (p >= seg_end);
        load_w();
      }
      const int32_t lim = p1 < seg_end ? p1 : seg_end;
      const int nb = lim - p < U ? lim - p : U;  // rows of this batch, all in the current segment

      float4 x[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < nb) {
          const int64_t r = perm[p + u];
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c = col0 + v * 128;
            x[u][v] = c < D ? ld_stream_f4(X + r * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      double d2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double t = 0.0;
        if (u < nb) {
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const double a = (double)x[u][v].x - w[v][0], b = (double)x[u][v].y - w[v][1];
            const double c = (double)x[u][v].z - w[v][2], e = (double)x[u][v].w - w[v][3];
            t = fma(a, a, t);
            t = fma(b, b, t);
            t = fma(c, c, t);
            t = fma(e, e, t);
          }
          t = warp_sum(t);
        }
        d2[u] = t;
      }
      if (WPR > 1) {
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < U; ++u) red[team][parity][tw][u] = d2[u];
        }
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(WPR * 32));
#pragma unroll
        for (int u = 0; u < U; ++u) {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < WPR; ++q) t += red[team][parity][q][u];
          d2[u] = t;
        }
        parity ^= 1;
      }
      // lane u evaluates the weight and the distance of row u
      double my_d2 = d2[0];
#pragma unroll
      for (int u = 1; u < U; ++u)
        if ((lane % U) == u) my_d2 = d2[u];
      const double my_dist = sqrt(my_d2);
      const double my_k = 1.0 - sqrt(1.0 - exp(-inv_var * (my_dist * my_dist)));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < nb) {
          const double k = __shfl_sync(kFullMask, my_k, u);
          run_k += k;
          run_d += __shfl_sync(kFullMask, my_dist, u);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            acc[v][0] = fma(k, (double)x[u][v].x, acc[v][0]);
            acc[v][1] = fma(k, (double)x[u][v].y, acc[v][1]);
            acc[v][2] = fma(k, (double)x[u][v].z, acc[v][2]);
            acc[v][3] = fma(k, (double)x[u][v].w, acc[v][3]);
          }
        }
      }
      p += nb;
    }
    flush();
  }
}

template <int VPL, int WPR, int U>
int launch_accumulate(const dbgsom_accumulate_args& a, const int32_t* perm, const int32_t* offsets, cudaStream_t s) {
  constexpr int TEAMS = ACC_THREADS / 32 / WPR;
  int64_t blocks = ceil_div<int64_t>(ceil_div<int64_t>(a.N, ACC_CHUNK), TEAMS);
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  accumulate_kernel<VPL, WPR, U><<<(unsigned)blocks, ACC_THREADS, 0, s>>>(
      a.d_X, a.N, a.D, a.ldx, perm, offsets, a.d_W, a.M, a.inv_total_variance, a.d_part);
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace

struct AccWorkspace {
  int32_t *counts, *offsets, *cursor, *perm;
  static size_t bytes(int64_t N, int M) {
    return 3 * round_up<size_t>((size_t)(M + 1) * 4, 256) + round_up<size_t>((size_t)N * 4, 256);
  }
  static AccWorkspace carve(void* base, int64_t N, int M) {
    AccWorkspace w;
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    const size_t sm = round_up<size_t>((size_t)(M + 1) * 4, 256);
    w.counts = reinterpret_cast<int32_t*>(p);
    w.offsets = reinterpret_cast<int32_t*>(p + sm);
    w.cursor = reinterpret_cast<int32_t*>(p + 2 * sm);
    w.perm = reinterpret_cast<int32_t*>(p + 3 * sm);
    return w;
  }
};

size_t accumulate_workspace_bytes(int64_t N, int M) { return AccWorkspace::bytes(N, M); }

int run_accumulate(const dbgsom_accumulate_args& a, cudaStream_t s) {
  const AccWorkspace ws = AccWorkspace::carve(a.d_workspace, a.N, a.M);
  const int64_t M = a.M, D = a.D;
  DBGSOM_CUDA_TRY(cudaMemsetAsync(a.d_part, 0, (size_t)(M * D + 3 * M) * sizeof(double), s));
  DBGSOM_CUDA_TRY(cudaMemsetAsync(ws.counts, 0, (size_t)(M + 1) * 4, s));
  const bool with_labels = a.d_labels != nullptr && a.d_class_hist != nullptr && a.n_classes > 0;
  if (with_labels) DBGSOM_CUDA_TRY(cudaMemsetAsync(a.d_class_hist, 0, (size_t)M * a.n_classes * 4, s));

  // histogram
  {
    const int64_t bins = M + (with_labels ? M * a.n_classes : 0);
    const bool smem = bins <= HIST_SMEM_BINS;
    const int64_t rows_per_block = smem ? 16384 : 4096;
    const unsigned blocks = (unsigned)ceil_div<int64_t>(a.N, rows_per_block);
    if (smem)
      hist_kernel<true><<<blocks, HIST_THREADS, (size_t)bins * 4, s>>>(a.d_bmu, a.N, a.M, with_labels ? a.d_labels : nullptr,
                                                                     a.n_classes, ws.counts, a.d_class_hist, rows_per_block);
    else
      hist_kernel<false><<<blocks, HIST_THREADS, 0, s>>>(a.d_bmu, a.N, a.M, with_labels ? a.d_labels : nullptr,
                                                        a.n_classes, ws.counts, a.d_class_hist, rows_per_block);
    DBGSOM_LAUNCH_CHECK();
  }
  scan_kernel<<<1, SCAN_THREADS, 0, s>>>(ws.counts, a.M, ws.offsets, ws.cursor, a.d_part + M * D + M);
  DBGSOM_LAUNCH_CHECK();
  {
    int64_t blocks = ceil_div<int64_t>(a.N, SCAT_THREADS * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    scatter_kernel<<<(unsigned)blocks, SCAT_THREADS, 0, s>>>(a.d_bmu, a.N, a.M, ws.offsets, ws.cursor, ws.perm);
    DBGSOM_LAUNCH_CHECK();
  }
  const int D4 = a.D;
  if (D4 <= 128) return launch_accumulate<1, 1, 4>(a, ws.perm, ws.offsets, s);
  if (D4 <= 256) return launch_accumulate<2, 1, 4>(a, ws.perm, ws.offsets, s);
  if (D4 <= 512) return launch_accumulate<4, 1, 2>(a, ws.perm, ws.offsets, s);
  if (D4 <= 1024) return launch_accumulate<4, 2, 2>(a, ws.perm, ws.offsets, s);
  if (D4 <= 2048) return launch_accumulate<4, 4, 2>(a, ws.perm, ws.offsets, s);
  if (D4 <= 4096) return launch_accumulate<4, 8, 2>(a, ws.perm, ws.offsets, s);
  return DBGSOM_E_UNSUPPORTED;
}

}  // namespace dbgsom
