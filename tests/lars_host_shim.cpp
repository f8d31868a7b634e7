// TEST-ONLY host instantiation of dbgsom_b200/csrc/lars_core.cuh: lets the CPU suite check the restatement of
// scikit-learn's LARS-lasso path against scikit-learn itself.  Never built into libdbgsom_b200.so.
#include <cstdlib>
#include <vector>

#include "../dbgsom_b200/csrc/lars_core.cuh"

extern "C" int lars_host(const double* gram, int M, const double* cov, long long N, int n_features, int max_iter, int A,
                         double* code, int* status) {
  std::vector<double> scratch(dbgsom::lars_scratch_doubles(M, A));
  std::vector<int32_t> idx(M);
  for (long long i = 0; i < N; ++i) {
    dbgsom::LarsMem S{scratch.data(), 1};
    dbgsom::LarsIdx I{idx.data(), 1};
    status[i] = dbgsom::lars_lasso_positive(gram, M, cov + i * M, 1, n_features, max_iter, A, S, I, code + i * M, 1);
  }
  return 0;
}
