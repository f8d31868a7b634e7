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
// The accumulate kernel gives every team of warps one contiguous range of the permutation; the
// team keeps the current segment's float64 prototype in registers (so the exact distance
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
constexpr int ACC_STAGES = 2;

__device__ __forceinline__ uint32_t acc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void acc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ACC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ACC_DONE;\n"
      "bra ACC_WAIT;\n"
      "ACC_DONE:\n"
      "}\n" ::"r"(acc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// float -> double without the XU pipe.  ncu showed the first version of this kernel bound by
// F2F.F64.F32 (sm__inst_executed_pipe_xu at 109 % of peak: ~4 conversions per clock and SM); the
// same value comes from two integer multiplies and two logic ops on the ALU/FMA pipes: the float's
// exponent/mantissa field shifted right by 3 plus the bias difference (1023 - 127) << 20.  Exact
// for normal floats; zeros and denormals come out with magnitude < 1.2e-38 instead of exactly
// themselves, far below anything a float64 sum of the data can resolve; inputs are finite
// (check_array), so the inf/NaN encodings never occur.
__device__ __forceinline__ double f32_as_f64(float f) {
  const uint32_t b = __float_as_uint(f);
  const uint32_t t = b & 0x7fffffffu;
  const uint32_t hi = (__umulhi(t, 0x20000000u) + 0x38000000u) | (b & 0x80000000u);
  return __hiloint2double((int)hi, (int)(b << 29));
}

// Sum each of U per-lane values over the warp with a transposed butterfly (the number of live
// values halves while the lane distance halves): 6 double shuffles for U = 4 instead of 20.
// Afterwards lane l holds the total of row row_of_lane(l); every row is held by 32 / U lanes.
template <int U>
__device__ __forceinline__ int row_of_lane(int lane) {
  return U == 4 ? ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1) : U == 2 ? (lane >> 4) & 1 : 0;
}
template <int U>
__device__ __forceinline__ double reduce_rows(const double (&v)[U], int lane) {
  double b;
  if (U == 4) {
    const bool h16 = lane & 16, h8 = lane & 8;
    const double a0 = (h16 ? v[2] : v[0]) + __shfl_xor_sync(kFullMask, h16 ? v[0] : v[2], 16);
    const double a1 = (h16 ? v[3] : v[1]) + __shfl_xor_sync(kFullMask, h16 ? v[1] : v[3], 16);
    b = (h8 ? a1 : a0) + __shfl_xor_sync(kFullMask, h8 ? a0 : a1, 8);
    b += __shfl_xor_sync(kFullMask, b, 4);
  } else if (U == 2) {
    const bool h16 = lane & 16;
    b = (h16 ? v[1] : v[0]) + __shfl_xor_sync(kFullMask, h16 ? v[0] : v[1], 16);
    b += __shfl_xor_sync(kFullMask, b, 8);
    b += __shfl_xor_sync(kFullMask, b, 4);
  } else {
    b = v[0];
#pragma unroll
    for (int o = 16; o > 2; o >>= 1) b += __shfl_xor_sync(kFullMask, b, o);
  }
  b += __shfl_xor_sync(kFullMask, b, 2);
  b += __shfl_xor_sync(kFullMask, b, 1);
  return b;
}
template <int U>
__device__ __forceinline__ int lane_of_row(int u) {
  return U == 4 ? (u >> 1) * 16 + (u & 1) * 8 : U == 2 ? u * 16 : 0;
}

// Position of a team in the sorted sample sequence: next position, its segment, the segment's end.
struct SegCursor {
  int32_t p, seg, seg_end;
  // rows of the next batch: at most U, never across a segment boundary or the end of the range
  __device__ __forceinline__ int next_batch(const int32_t* __restrict__ offsets, int32_t p_end, int U) {
    while (p >= seg_end) {  // skip empty segments
      ++seg;
      seg_end = offsets[seg + 1];
    }
    const int32_t lim = p_end < seg_end ? p_end : seg_end;
    return lim - p < U ? lim - p : U;
  }
};

// VPL : float4 per lane per row slab;  WPR : warps cooperating on one row (team size);
// U   : rows per batch (their sample weights are evaluated together, lane u doing row u, so the
//       float64 exp/sqrt costs 1/U per row).
// Every team owns one contiguous range of the sorted sequence.  Rows travel HBM -> shared memory by
// 1-D bulk async copies (cp.async.bulk, one instruction per row, completion on an mbarrier) into a
// ring of ACC_STAGES batches per team, so the copies of the next batches are in flight while the
// current one is reduced and no registers are tied up by loads.  A team handles columns
// [tw * 128 * VPL, (tw + 1) * 128 * VPL) of each row with warp tw.
// All arithmetic is float64: distances by direct differences, k = 1 - sqrt(1 - exp(-d^2 / V))
// literally as dbgsom/BaseSom.py:536-537, sums in float64 registers.
template <int VPL, int WPR, int U>
__global__ void __launch_bounds__(ACC_THREADS) accumulate_kernel(
    const float* __restrict__ X, int64_t N, int D, int64_t ldx, const int32_t* __restrict__ perm,
    const int32_t* __restrict__ offsets, const double* __restrict__ W, int M, double inv_var,
    double* __restrict__ part) {
  constexpr int TEAMS = ACC_THREADS / 32 / WPR;
  extern __shared__ __align__(16) float rows_smem[];  // [TEAMS][ACC_STAGES][U][D]
  __shared__ __align__(8) uint64_t full_bar[TEAMS][ACC_STAGES];
  __shared__ double red[TEAMS][2][WPR][U];  // cross-warp partial squared distances (double buffered)

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int team = warp / WPR;
  const int tw = warp % WPR;
  const int col0 = tw * 128 * VPL + lane * 4;
  const int32_t total = offsets[M];  // samples that have a winner (== N without NaN rows)
  const bool issuer = tw == 0 && lane == 0;
  float* ring = rows_smem + (size_t)team * ACC_STAGES * U * D;
  const uint32_t row_bytes = (uint32_t)D * 4u;

  if (issuer) {
    for (int s = 0; s < ACC_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(acc_smem_u32(&full_bar[team][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double* __restrict__ Sk = part;
  double* __restrict__ sk = part + (int64_t)M * D;
  double* __restrict__ En = sk + 2 * (int64_t)M;

  const int64_t n_teams = (int64_t)gridDim.x * TEAMS;
  const int64_t team_id = (int64_t)blockIdx.x * TEAMS + team;
  const int32_t p_begin = (int32_t)((int64_t)total * team_id / n_teams);
  const int32_t p_end = (int32_t)((int64_t)total * (team_id + 1) / n_teams);
  if (p_begin >= p_end) return;  // whole team exits together (no block-wide barrier follows)

  // segment containing p_begin: largest j with offsets[j] <= p_begin
  int lo = 0, hi = M;  // invariant offsets[lo] <= p_begin < offsets[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= p_begin) lo = mid; else hi = mid;
  }
  SegCursor use{p_begin, lo, offsets[lo + 1]};  // consumer position
  SegCursor pre = use;                          // prefetch position (runs ACC_STAGES - 1 batches ahead)

  auto issue = [&](int stage) {  // all lanes advance the cursor, one lane issues the copies
    if (pre.p >= p_end) return;
    const int nb = pre.next_batch(offsets, p_end, U);
    if (issuer) {
      uint64_t* bar = &full_bar[team][stage];
      // order the team's earlier generic-proxy reads of this slot before the async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(acc_smem_u32(bar)), "r"(nb * row_bytes)
                   : "memory");
      for (int u = 0; u < nb; ++u) {
        const float* src = X + (int64_t)perm[pre.p + u] * ldx;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                acc_smem_u32(ring + ((size_t)stage * U + u) * D)),
            "l"(src), "r"(row_bytes), "r"(acc_smem_u32(bar))
            : "memory");
      }
    }
    pre.p += nb;
  };
#pragma unroll
  for (int s = 0; s < ACC_STAGES - 1; ++s) issue(s);

  double w[VPL][4], acc[VPL][4];
  double run_k = 0.0, run_d = 0.0;
  int seg = -1;
  auto load_w = [&](int sgm) {
    seg = sgm;
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

  int stage = 0, parity = 0;
  uint32_t phase = 0;
  while (use.p < p_end) {
    const int nb = use.next_batch(offsets, p_end, U);
    if (use.seg != seg) {
      if (seg >= 0) flush();
      load_w(use.seg);
    }
    issue((stage + ACC_STAGES - 1) % ACC_STAGES);  // refill the slot consumed in the previous iteration
    acc_mbar_wait(&full_bar[team][stage], phase);
    const float* rows = ring + (size_t)stage * U * D;

    double part_d2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double t = 0.0;
      if (u < nb) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = col0 + v * 128;
          if (c < D) {
            const float4 x = *reinterpret_cast<const float4*>(rows + u * D + c);
            // one conversion in four goes through the XU pipe (F2F, ~27 clk per warp instruction, otherwise
            // idle), the rest through the ALU/FMA pipes: neither saturates
            const double a = (double)x.x - w[v][0], b = f32_as_f64(x.y) - w[v][1];
            const double cc = f32_as_f64(x.z) - w[v][2], e = f32_as_f64(x.w) - w[v][3];
            t = fma(a, a, t);
            t = fma(b, b, t);
            t = fma(cc, cc, t);
            t = fma(e, e, t);
          }
        }
      }
      part_d2[u] = t;
    }
    double my_d2 = reduce_rows<U>(part_d2, lane);  // lane l: squared distance of row row_of_lane(l)
    if (WPR > 1) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lane == lane_of_row<U>(u)) red[team][parity][tw][u] = my_d2;
      asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(WPR * 32));
      my_d2 = 0.0;
#pragma unroll
      for (int q = 0; q < WPR; ++q) my_d2 += red[team][parity][q][row_of_lane<U>(lane)];
      parity ^= 1;
    }
    // each lane evaluates the weight and the distance of its row (32 / U lanes per row, redundantly)
    const double my_dist = sqrt(my_d2);
    const double my_k = 1.0 - sqrt(1.0 - exp(-inv_var * (my_dist * my_dist)));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u < nb) {
        const double k = __shfl_sync(kFullMask, my_k, lane_of_row<U>(u));
        run_k += k;
        run_d += __shfl_sync(kFullMask, my_dist, lane_of_row<U>(u));
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = col0 + v * 128;
          if (c < D) {
            const float4 x = *reinterpret_cast<const float4*>(rows + u * D + c);
            acc[v][0] = fma(k, f32_as_f64(x.x), acc[v][0]);
            acc[v][1] = fma(k, f32_as_f64(x.y), acc[v][1]);
            acc[v][2] = fma(k, f32_as_f64(x.z), acc[v][2]);
            acc[v][3] = fma(k, f32_as_f64(x.w), acc[v][3]);
          }
        }
      }
    }
    use.p += nb;
    // the slot may be overwritten by the next issue(): every lane of the team is done reading it
    if (WPR > 1)
      asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(WPR * 32));
    else
      __syncwarp();
    if (++stage == ACC_STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
  if (seg >= 0) flush();
}

template <int VPL, int WPR, int U>
int launch_accumulate(const dbgsom_accumulate_args& a, const int32_t* perm, const int32_t* offsets, cudaStream_t s) {
  constexpr int TEAMS = ACC_THREADS / 32 / WPR;
  const size_t smem = (size_t)TEAMS * ACC_STAGES * U * a.D * sizeof(float);
  auto kern = accumulate_kernel<VPL, WPR, U>;
  DBGSOM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 2048));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = 148 * per_sm;
  const int64_t useful = ceil_div<int64_t>(ceil_div<int64_t>(a.N, 4 * U), TEAMS);  // >= 4 batches per team
  if (blocks > useful) blocks = useful;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, ACC_THREADS, smem, s>>>(a.d_X, a.N, a.D, a.ldx, perm, offsets, a.d_W, a.M,
                                                  a.inv_total_variance, a.d_part);
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
