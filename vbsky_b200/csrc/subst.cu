// Substitution models (host): closed-form HKY/K80 spectrum standing in for
// msprime.HKY + numpy.linalg.eigh of substitution.py:73-87.
#include <cmath>

#include "common.h"

using namespace vbsky;

extern "C" int vbsky_hky(double kappa, vbsky_subst_model* out) {
  VB_CHECK_ARG(out, "out is NULL");
  VB_CHECK_ARG(kappa > 0 && std::isfinite(kappa), "kappa must be positive and finite (got %g)", kappa);
  // Off-diagonal rates after msprime's scaling (largest off-diagonal row sum = 1) with uniform
  // base frequencies: transversion beta = 1/(2+kappa), transition alpha = kappa/(2+kappa);
  // the reference then sets the diagonal to -1 (substitution.py:77).
  const double beta = 1.0 / (2.0 + kappa);
  const double alpha = kappa * beta;
  const double h = 0.5, r = std::sqrt(0.5);
  // Orthonormal eigenvectors as columns of P (state order A,C,G,T):
  //   v0 = (1,0,-1,0)/sqrt2  [A vs G]   lambda = -2(alpha+beta)
  //   v1 = (0,1,0,-1)/sqrt2  [C vs T]   lambda = -2(alpha+beta)
  //   v2 = (1,-1,1,-1)/2     [purine vs pyrimidine]  lambda = -4 beta
  //   v3 = (1,1,1,1)/2       lambda = 0
  // (ascending eigenvalue order, as eigh returns them).
  const double V[4][4] = {{r, 0, -r, 0}, {0, r, 0, -r}, {h, -h, h, -h}, {h, h, h, h}};
  const double lam[4] = {-2.0 * (alpha + beta), -2.0 * (alpha + beta), -4.0 * beta, 0.0};
  for (int k = 0; k < 4; ++k) {
    out->d[k] = lam[k];
    out->pi[k] = 0.25;
    for (int i = 0; i < 4; ++i) {
      out->P[i * 4 + k] = V[k][i];
      out->Pinv[k * 4 + i] = V[k][i];
    }
  }
  return VBSKY_OK;
}

extern "C" int vbsky_hky_dkappa(double kappa, double* dd) {
  VB_CHECK_ARG(dd, "dd is NULL");
  VB_CHECK_ARG(kappa > 0 && std::isfinite(kappa), "kappa must be positive and finite (got %g)", kappa);
  // d/dkappa of the spectrum above; the eigenvectors do not depend on kappa, so dQ/dkappa = P diag(dd) Pinv
  const double s = 1.0 / ((2.0 + kappa) * (2.0 + kappa));
  dd[0] = dd[1] = -2.0 * s, dd[2] = 4.0 * s, dd[3] = 0.0;
  return VBSKY_OK;
}

extern "C" int vbsky_subst_rate_matrix(const vbsky_subst_model* m, double* q) {
  VB_CHECK_ARG(m && q, "NULL argument");
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += m->P[i * 4 + k] * m->d[k] * m->Pinv[k * 4 + j];
      q[i * 4 + j] = s;
    }
  return VBSKY_OK;
}
