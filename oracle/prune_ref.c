/*
 * CPU oracle (TEST INFRASTRUCTURE, not product code): C/OpenMP restatement of the two scans of
 * /root/reference/vbsky/prune.py and of its eq.-(9) gradient, one site pattern at a time exactly as
 * the reference evaluates it under vmap.  Used as the timed CPU baseline (bench.py cpu_baseline /
 * --impl reference, kind "port") and cross-checked against the NumPy oracle in tests/.
 *
 *   expm_action          substitution.py:64-70  (incl. clip to [1e-40, 1])
 *   postorder partials   prune.py:15-49         (parents n..2n-2 in index order, rescale by the sum)
 *   preorder partials    prune.py:52-96         (reversed postorder[:-1])
 *   log-likelihood       prune.py:128-131
 *   gradient             prune.py:134-150       (einsum "nu,uv,nv->n" times exp(lpre+lpost-logP))
 *   data term            bdsky.py:331-343       (sum over patterns weighted by counts)
 *
 * exp(t*d) does not depend on the site pattern, so XLA evaluates it once per node outside the
 * vmapped axis; it is hoisted here in the same way.  Everything else is per site.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double clip(double x) { return x < 1e-40 ? 1e-40 : (x > 1.0 ? 1.0 : x); }

/* right=True: P @ (e * (Pinv @ p)) ; right=False: ((p @ P) * e) @ Pinv */
static inline void expm_action(const double* P, const double* Pinv, const double* e, const double* p, int right, double* out) {
  double tmp[4];
  if (right) {
    for (int k = 0; k < 4; ++k) {
      double s = 0;
      for (int j = 0; j < 4; ++j) s += Pinv[k * 4 + j] * p[j];
      tmp[k] = e[k] * s;
    }
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += P[i * 4 + k] * tmp[k];
      out[i] = clip(s);
    }
  } else {
    for (int k = 0; k < 4; ++k) {
      double s = 0;
      for (int j = 0; j < 4; ++j) s += p[j] * P[j * 4 + k];
      tmp[k] = s * e[k];
    }
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += tmp[k] * Pinv[k * 4 + i];
      out[i] = clip(s);
    }
  }
}

void vbsky_oracle_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

int vbsky_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* patterns: [S][n] 4-bit state masks (bit0=A..bit3=T) == the 0/1 tip partials; counts [S].
 * Outputs: site_ll [S] (may be NULL), *ll_out, grad [2n-2] (may be NULL => value only). */
int vbsky_oracle_prune(int n, const int32_t* parent_child, const int32_t* child_parent, const int32_t* postorder,
                       const int32_t* siblings, const double* bl, const double* pi, const double* P, const double* d,
                       const double* Pinv, int64_t S, const uint8_t* patterns, const double* counts, double* site_ll,
                       double* ll_out, double* grad) {
  const int N = 2 * n - 1;
  double Q[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += P[i * 4 + k] * d[k] * Pinv[k * 4 + j];
      Q[i * 4 + j] = s;
    }
  double* E = (double*)malloc(sizeof(double) * 4 * (N - 1));
  for (int c = 0; c < N - 1; ++c)
    for (int k = 0; k < 4; ++k) E[c * 4 + k] = exp(bl[c] * d[k]);
  double ll_tot = 0.0;
  int nthreads = vbsky_oracle_threads();
  double* gacc = grad ? (double*)calloc((size_t)nthreads * (N - 1), sizeof(double)) : NULL;
#pragma omp parallel reduction(+ : ll_tot)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double* post = (double*)malloc(sizeof(double) * 4 * N);
    double* pre = (double*)malloc(sizeof(double) * 4 * N);
    double* lpost = (double*)malloc(sizeof(double) * N);
    double* lpre = (double*)malloc(sizeof(double) * N);
    double* g = gacc ? gacc + (size_t)tid * (N - 1) : NULL;
#pragma omp for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
      const uint8_t* pat = patterns + s * n;
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 4; ++j) post[i * 4 + j] = (double)((pat[i] >> j) & 1);
        lpost[i] = 0.0;
      }
      /* prune.py:24-38 */
      for (int parent = n; parent < N; ++parent) {
        const int c0 = parent_child[parent * 2], c1 = parent_child[parent * 2 + 1];
        double a0[4], a1[4], pp[4];
        expm_action(P, Pinv, E + c0 * 4, post + c0 * 4, 1, a0);
        expm_action(P, Pinv, E + c1 * 4, post + c1 * 4, 1, a1);
        double c = 0;
        for (int j = 0; j < 4; ++j) pp[j] = a0[j] * a1[j], c += pp[j];
        for (int j = 0; j < 4; ++j) post[parent * 4 + j] = pp[j] / c;
        lpost[parent] = log(c) + (lpost[c0] + lpost[c1]);
      }
      double like = 0;
      for (int j = 0; j < 4; ++j) like += post[(N - 1) * 4 + j] * pi[j];
      const double logP = log(like) + lpost[N - 1];
      if (site_ll) site_ll[s] = logP;
      ll_tot += counts[s] * logP;
      if (!g) continue;
      /* prune.py:62-95 */
      for (int j = 0; j < 4; ++j) pre[(N - 1) * 4 + j] = pi[j];
      lpre[N - 1] = 0.0;
      for (int idx = N - 2; idx >= 0; --idx) {
        const int child = postorder[idx];
        const int parent = child_parent[child], sib = siblings[child];
        double A0[4], A[4], B[4];
        expm_action(P, Pinv, E + sib * 4, post + sib * 4, 1, A0);
        for (int j = 0; j < 4; ++j) A[j] = pre[parent * 4 + j] * A0[j];
        expm_action(P, Pinv, E + child * 4, A, 0, B);
        double c = 0;
        for (int j = 0; j < 4; ++j) c += B[j];
        for (int j = 0; j < 4; ++j) pre[child * 4 + j] = B[j] / c;
        lpre[child] = log(c) + lpre[parent] + lpost[sib];
      }
      /* prune.py:146-148 */
      for (int i = 0; i < N - 1; ++i) {
        double v = 0;
        for (int u = 0; u < 4; ++u) {
          double qv = 0;
          for (int w = 0; w < 4; ++w) qv += Q[u * 4 + w] * post[i * 4 + w];
          v += pre[i * 4 + u] * qv;
        }
        g[i] += counts[s] * v * exp(lpre[i] + lpost[i] - logP);
      }
    }
    free(post), free(pre), free(lpost), free(lpre);
  }
  *ll_out = ll_tot;
  if (grad) {
    for (int i = 0; i < N - 1; ++i) {
      double s = 0;
      for (int t = 0; t < nthreads; ++t) s += gacc[(size_t)t * (N - 1) + i];
      grad[i] = s;
    }
    free(gacc);
  }
  free(E);
  return 0;
}
