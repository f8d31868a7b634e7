// Non-negative LARS-lasso path of ONE sample over a dictionary given by its Gram matrix
// (SURVEY.md section 8(f) rank 3).
//
// Replaces, for BaseSom.transform (dbgsom/BaseSom.py:241-268) and SomClassifier.predict_proba
// (dbgsom/SomClassifier.py:178-220), what scikit-learn 1.9.0 executes per sample behind
//   SparseCoder(dictionary=normalize(W), positive_code=True, transform_alpha=0,
//               transform_algorithm="lasso_lars").transform(normalize(X)):
// sklearn/decomposition/_dict_learning.py `_sparse_encode_precomputed` -> LassoLars(alpha=0,
// fit_intercept=False, precompute=gram, fit_path=False, positive=True, max_iter=1000) ->
// sklearn/linear_model/_least_angle.py `_lars_path_solver` (:415-897) in Gram mode with
// method="lasso", positive=True, alpha_min=0, return_path=False.  This file restates that routine
// step by step (line references are to _least_angle.py), including its stopping rules, the rounding
// of the equiangular correlations to 15 decimals (:776), `min_pos`, the tiny32 guards and the Cholesky
// append / delete updates (sklearn/utils/arrayfuncs.pyx `cholesky_delete`).  The reference swaps rows
// and columns of a private copy of the Gram matrix; here the swaps are carried by the index permutation
// `idx` (position -> atom), which is equivalent because every swap is a symmetric row + column swap.
//
// The routine is written against an accessor (strided per-thread state) so that one CUDA thread runs
// one sample with coalesced state; it is also instantiated on the host by tests/lars_host_shim.cu, which
// exists only to check this restatement against scikit-learn without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LARS_HD __host__ __device__ __forceinline__
#else
#define LARS_HD inline
#endif

namespace dbgsom {

enum LarsStatus : int32_t {
  LARS_OK = 0,
  LARS_CAPACITY = 1,   // the active set outgrew the Cholesky capacity: rerun with a larger one
  LARS_EARLY_STOP = 2, // "alpha is increasing" bail-out of the reference (:733-747); result is valid
  LARS_DEGENERATE = 4, // a degenerate regressor was dropped (:703-724); result follows the reference
  LARS_MULTI_DROP = 8, // several coefficients crossed zero in the same step (ties in z)
};

// per-thread strided arrays: element e of this thread lives at base[e * stride]
struct LarsMem {
  double* base;
  int64_t stride;
  LARS_HD double& operator()(int64_t e) const { return base[e * stride]; }
};
struct LarsIdx {
  int32_t* base;
  int64_t stride;
  LARS_HD int32_t& operator()(int64_t e) const { return base[e * stride]; }
};

// Layout of the double scratch of one thread: covp[M] | coef[M] | prev[M] | ced[M] | L[A*A] | ls[A] | tmp[A]
LARS_HD int64_t lars_scratch_doubles(int M, int A) { return 4 * (int64_t)M + (int64_t)A * A + 2 * (int64_t)A; }

// gram [M, M] row-major (shared, read only); cov_in[j * cov_stride] = <dictionary_j, x>;
// code_out[j * code_stride] receives the coefficients.  Returns an OR of LarsStatus flags.
LARS_HD int32_t lars_lasso_positive(const double* __restrict__ gram, int M, const double* __restrict__ cov_in,
                                    int64_t cov_stride, int n_features, int max_iter, int A, LarsMem S, LarsIdx idx,
                                    double* __restrict__ code_out, int64_t code_stride) {
  const double eps = 2.220446049250313e-16;       // np.finfo(float).eps
  const double tiny32 = 1.1754943508222875e-38;   // np.finfo(np.float32).tiny
  const double eq_tol = 1.1920928955078125e-07;   // np.finfo(np.float32).eps
  const double dbl_max = 1.7976931348623157e308;
  const int64_t oCov = 0, oCoef = M, oPrev = 2 * (int64_t)M, oCed = 3 * (int64_t)M, oL = 4 * (int64_t)M;
  const int64_t oLs = oL + (int64_t)A * A, oTmp = oLs + A;
#define COV(p) S(oCov + (p))
#define COEF(j) S(oCoef + (j))
#define PREV(j) S(oPrev + (j))
#define CED(p) S(oCed + (p))
#define LL(r, c) S(oL + (int64_t)(r) * A + (c))
#define LS(k) S(oLs + (k))
#define TMP(k) S(oTmp + (k))
  for (int p = 0; p < M; ++p) {
    COV(p) = cov_in[p * cov_stride];
    COEF(p) = 0.0;
    PREV(p) = 0.0;
    idx(p) = p;
  }
  int n_active = 0, n_iter = 0;
  bool drop = false;
  double alpha = 0.0, prev_alpha = 0.0;
  int32_t status = LARS_OK;
  const int max_features = max_iter < M ? max_iter : M;

  while (true) {
    // :633-647  largest remaining covariance (first occurrence, like np.argmax)
    double C = 0.0;
    int C_pos = n_active;
    if (n_active < M) {
      C = COV(n_active);
      for (int p = n_active + 1; p < M; ++p) {
        const double v = COV(p);
        if (v > C) {
          C = v;
          C_pos = p;
        }
      }
    }
    // :655-668  stopping on alpha
    alpha = C / (double)n_features;
    if (alpha <= 0.0 + eq_tol) {
      if (fabs(alpha - 0.0) > eq_tol) {
        if (n_iter > 0) {
          const double ss = (prev_alpha - 0.0) / (prev_alpha - alpha);
          for (int j = 0; j < M; ++j) COEF(j) = PREV(j) + ss * (COEF(j) - PREV(j));
        }
        alpha = 0.0;
      }
      break;
    }
    if (n_iter >= max_iter || n_active >= M) break;  // :670-671

    if (!drop) {
      // :672-731  append the winner to the Cholesky factor of the active Gram block
      if (n_active >= A || n_active >= max_features) {
        status |= LARS_CAPACITY;
        break;
      }
      const int m = n_active;
      {
        const double t = COV(C_pos);
        COV(C_pos) = COV(m);
        COV(m) = t;
        const int32_t ti = idx(C_pos);
        idx(C_pos) = idx(m);
        idx(m) = ti;
      }
      const int64_t gm = (int64_t)idx(m) * M;
      const double c = gram[gm + idx(m)];
      for (int k = 0; k < m; ++k) LL(m, k) = gram[gm + idx(k)];
      // forward substitution: L[:m,:m] w = L[m,:m]
      for (int r = 0; r < m; ++r) {
        double acc = LL(m, r);
        for (int k = 0; k < r; ++k) acc -= LL(r, k) * LL(m, k);
        LL(m, r) = acc / LL(r, r);
      }
      double v = 0.0;
      for (int k = 0; k < m; ++k) v += LL(m, k) * LL(m, k);
      double diag = sqrt(fabs(c - v));
      if (diag < eps) diag = eps;
      LL(m, m) = diag;
      if (diag < 1e-7) {
        // :703-724  degenerate regressor: its covariance is zeroed and swapped back; the index swap stays
        status |= LARS_DEGENERATE;
        COV(m) = 0.0;
        const double t = COV(C_pos);
        COV(C_pos) = COV(m);
        COV(m) = t;
        continue;
      }
      n_active += 1;
    }

    if (n_iter > 0 && prev_alpha < alpha) {  // :733-747
      status |= LARS_EARLY_STOP;
      break;
    }

    // :749-775  least squares direction: (L L^T) ls = 1, AA = 1 / sqrt(sum ls)
    double AA;
    {
      int reg = -1;  // -1: plain solve; >= 0: diagonal regularised by 2^reg * eps (cumulative, :764-772)
      while (true) {
        for (int r = 0; r < n_active; ++r) {
          double acc = 1.0;
          for (int k = 0; k < r; ++k) acc -= LL(r, k) * TMP(k);
          TMP(r) = acc / LL(r, r);
        }
        for (int r = n_active - 1; r >= 0; --r) {
          double acc = TMP(r);
          for (int k = r + 1; k < n_active; ++k) acc -= LL(k, r) * LS(k);
          LS(r) = acc / LL(r, r);
        }
        double sum = 0.0;
        for (int k = 0; k < n_active; ++k) sum += LS(k);
        if (reg < 0) {
          if (n_active == 1 && LS(0) == 0.0) {
            LS(0) = 1.0;
            AA = 1.0;
            break;
          }
          AA = 1.0 / sqrt(sum);
        } else {
          AA = 1.0 / sqrt(sum > eps ? sum : eps);
        }
        if (isfinite(AA)) break;
        if (reg < 0) {  // keep the exact diagonal (the reference regularises a copy of L)
          for (int k = 0; k < n_active; ++k) CED(k) = LL(k, k);
        }
        reg += 1;
        const double bump = ldexp(eps, reg);
        for (int k = 0; k < n_active; ++k) LL(k, k) += bump;
      }
      if (reg >= 0)
        for (int k = 0; k < n_active; ++k) LL(k, k) = CED(k);
      for (int k = 0; k < n_active; ++k) LS(k) *= AA;
    }

    // :783-796  correlation of the inactive atoms with the equiangular direction, rounded to 15 decimals,
    //           and the step to the next atom joining
    double g1 = dbl_max;
    for (int p = n_active; p < M; ++p) {
      const int64_t gp = (int64_t)idx(p) * M;  // Gram is symmetric: row of the inactive atom
      double acc = 0.0;
      for (int k = 0; k < n_active; ++k) acc += gram[gp + idx(k)] * LS(k);
      acc = rint(acc * 1e15) / 1e15;
      CED(p) = acc;
      const double q = (C - COV(p)) / (AA - acc + tiny32);
      if (0.0 < q && q < g1) g1 = q;
    }
    double gamma = g1 < C / AA ? g1 : C / AA;

    // :804-817  a coefficient crossing zero first?
    drop = false;
    double z_pos = dbl_max;
    for (int k = 0; k < n_active; ++k) {
      const double z = -COEF(idx(k)) / (LS(k) + tiny32);
      if (0.0 < z && z < z_pos) z_pos = z;
    }
    int drop_k = -1;
    if (z_pos < gamma) {
      for (int k = 0; k < n_active; ++k) {
        const double z = -COEF(idx(k)) / (LS(k) + tiny32);
        if (z == z_pos) {
          if (drop_k >= 0) status |= LARS_MULTI_DROP;
          drop_k = k;  // the reference walks ties from the back; a single crossing is the only regular case
        }
      }
      gamma = z_pos;
      drop = true;
    }
    n_iter += 1;

    // :832-842  new coefficients (inactive ones are zero), correlations
    for (int j = 0; j < M; ++j) {
      PREV(j) = COEF(j);
      COEF(j) = 0.0;
    }
    prev_alpha = alpha;
    for (int k = 0; k < n_active; ++k) COEF(idx(k)) = PREV(idx(k)) + gamma * LS(k);
    for (int p = n_active; p < M; ++p) COV(p) -= gamma * CED(p);

    if (drop) {
      // :845-890  remove the atom at active position drop_k: Cholesky delete (Givens rotations, as
      // arrayfuncs.cholesky_delete), shift it behind the active block, recompute its covariance
      const int n = n_active, go = drop_k;
      for (int i = go; i < n - 1; ++i)
        for (int k = 0; k <= i + 1; ++k) LL(i, k) = LL(i + 1, k);
      for (int i = go; i < n - 1; ++i) {
        // drotg on (a = L[i, i], b = L[i, i + 1]) of the shifted matrix
        double a = LL(i, i), b = LL(i, i + 1), cs, sn, r;
        const double roe = fabs(a) > fabs(b) ? a : b;
        const double scale = fabs(a) + fabs(b);
        if (scale == 0.0) {
          cs = 1.0;
          sn = 0.0;
          r = 0.0;
        } else {
          const double as = a / scale, bs = b / scale;
          r = scale * sqrt(as * as + bs * bs);
          if (roe < 0.0) r = -r;
          cs = a / r;
          sn = b / r;
        }
        LL(i, i) = r;
        if (LL(i, i) < 0.0) {  // diagonals cannot be negative
          LL(i, i) = fabs(LL(i, i));
          cs = -cs;
          sn = -sn;
        }
        LL(i, i + 1) = 0.0;
        // drot on the column pair (i, i + 1) of the rows below
        for (int rr = i + 1; rr < n - 1; ++rr) {
          const double x = LL(rr, i), y = LL(rr, i + 1);
          LL(rr, i) = cs * x + sn * y;
          LL(rr, i + 1) = cs * y - sn * x;
        }
      }
      n_active -= 1;
      const int32_t dropped = idx(go);
      for (int i = go; i < n_active; ++i) idx(i) = idx(i + 1);
      idx(n_active) = dropped;
      // temp = Cov_copy[drop_idx] - dot(Gram_copy[drop_idx], coef)
      double acc = 0.0;
      const int64_t gd = (int64_t)dropped * M;
      for (int j = 0; j < M; ++j) acc += gram[gd + j] * COEF(j);
      COV(n_active) = cov_in[dropped * cov_stride] - acc;
    }
  }
  for (int j = 0; j < M; ++j) code_out[j * code_stride] = COEF(j);
#undef COV
#undef COEF
#undef PREV
#undef CED
#undef LL
#undef LS
#undef TMP
  return status;
}

}  // namespace dbgsom
