// Non-negative sparse coding of samples over the normalised prototypes on the device.
//
// Replaces the scikit-learn call inside BaseSom.transform (dbgsom/BaseSom.py:241-268; also the first step of
// SomClassifier.predict_proba, dbgsom/SomClassifier.py:178-220): for every sample the LARS-lasso path with
// positive coefficients down to alpha = 0.  The per-sample routine is the step-by-step restatement of
// sklearn's `_lars_path_solver` in lars_core.cuh; here one CUDA thread runs one sample.  The path is
// sequential and data dependent (2-7 active atoms on trained maps), so the parallelism is across samples:
// per-thread state (covariances, coefficients, index permutation, Cholesky factor) lives in a global
// scratch laid out element-major ([element][thread]) so that the lanes of a warp touch consecutive addresses.
#include "common.cuh"
#include "lars_core.cuh"

namespace dbgsom {

namespace {

constexpr int LARS_THREADS = 128;

__global__ void __launch_bounds__(LARS_THREADS) sparse_code_kernel(const double* __restrict__ gram,
                                                                  const double* __restrict__ cov, int64_t N, int M,
                                                                  int n_features, int max_iter, int A,
                                                                  const int32_t* __restrict__ rows,
                                                                  double* __restrict__ scratch,
                                                                  int32_t* __restrict__ idx_scratch,
                                                                  double* __restrict__ code,
                                                                  int32_t* __restrict__ status) {
  const int64_t n_threads = (int64_t)gridDim.x * LARS_THREADS;
  const int64_t t = (int64_t)blockIdx.x * LARS_THREADS + threadIdx.x;
  const LarsMem S{scratch + t, n_threads};
  const LarsIdx I{idx_scratch + t, n_threads};
  for (int64_t q = t; q < N; q += n_threads) {
    const int64_t i = rows ? rows[q] : q;  // second launch: only the samples that outgrew the first capacity
    status[i] = lars_lasso_positive(gram, M, cov + i * M, 1, n_features, max_iter, A, S, I, code + i * M, 1);
  }
}

}  // namespace

size_t sparse_code_workspace_bytes(int64_t n_threads, int M, int A) {
  return (size_t)n_threads * ((size_t)lars_scratch_doubles(M, A) * sizeof(double) + (size_t)M * sizeof(int32_t)) + 256;
}

int run_sparse_code(const double* gram, const double* cov, int64_t N, int M, int n_features, int max_iter, int A,
                    const int32_t* rows, double* code, int32_t* status, void* workspace, size_t workspace_bytes,
                    cudaStream_t s) {
  // as many threads as the workspace holds (multiples of a CTA), at most one per sample and 16 CTAs per SM
  const size_t per_thread = (size_t)lars_scratch_doubles(M, A) * sizeof(double) + (size_t)M * sizeof(int32_t);
  int64_t blocks = (int64_t)((workspace_bytes - 256) / (per_thread * LARS_THREADS));
  const int64_t want = ceil_div<int64_t>(N, LARS_THREADS);
  if (blocks > want) blocks = want;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) return DBGSOM_E_WORKSPACE;
  const int64_t n_threads = blocks * LARS_THREADS;
  double* scratch = reinterpret_cast<double*>(workspace);
  int32_t* idx_scratch = reinterpret_cast<int32_t*>(scratch + n_threads * lars_scratch_doubles(M, A));
  sparse_code_kernel<<<(unsigned)blocks, LARS_THREADS, 0, s>>>(gram, cov, N, M, n_features, max_iter, A, rows, scratch,
                                                              idx_scratch, code, status);
  DBGSOM_LAUNCH_CHECK();
  return DBGSOM_OK;
}

}  // namespace dbgsom
