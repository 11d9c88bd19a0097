// bdsky.loglik (bdsky.py:268-345) for all (tree, sample) units: chain_forward -> bdsky -> prune ->
// chain_backward, enqueued on one stream with no host synchronisation.
#include "common.h"

namespace vbsky {

__global__ void combine_kernel(int U, int flags, const double* lt, const double* ld, double* ll, double* out_t, double* out_d) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const double a = (flags & VBSKY_TERM_TREE) ? lt[u] : 0.0;
  const double b = (flags & VBSKY_TERM_DATA) ? ld[u] : 0.0;
  ll[u] = a + b;  // bdsky.py:327,343
  if (out_t) out_t[u] = a;
  if (out_d) out_d[u] = b;
}

struct LoglikPlan {
  size_t heights, bl, xs, times, lam, psi, mu, dbl, dxs, dtimes, dlam, dpsi, dmu, lt, ld, dh, bdsky, prune, total;
  size_t bdsky_bytes, prune_bytes;
};

static int loglik_plan(const vbsky_layout* lay, const vbsky_tips* tips, int U, int m, LoglikPlan* pl) {
  const size_t N = 2 * (size_t)lay->n - 1;
  const size_t mm = (size_t)std::max(m, 1);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = (off + bytes + 255) / 256 * 256;
    return o;
  };
  const size_t d = sizeof(double);
  pl->heights = take(U * N * d), pl->bl = take(U * (N - 1) * d), pl->xs = take(U * N * d);
  pl->times = take(U * (mm + 1) * d), pl->lam = take(U * mm * d), pl->psi = take(U * mm * d), pl->mu = take(U * mm * d);
  pl->dbl = take(U * (N - 1) * d), pl->dxs = take(U * N * d), pl->dtimes = take(U * (mm + 1) * d);
  pl->dlam = take(U * mm * d), pl->dpsi = take(U * mm * d), pl->dmu = take(U * mm * d);
  pl->lt = take(U * d), pl->ld = take(U * d), pl->dh = take(U * N * d);
  pl->bdsky_bytes = vbsky_bdsky_workspace_bytes(U, lay->n);
  pl->bdsky = take(pl->bdsky_bytes);
  pl->prune_bytes = std::max(vbsky_prune_workspace_bytes(lay, tips, U, 1), vbsky_prune_workspace_bytes_f32(lay, tips, U, 1));
  if (pl->prune_bytes == 0) return VBSKY_EINVAL;
  pl->prune = take(pl->prune_bytes);
  pl->total = off;
  return VBSKY_OK;
}

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_loglik_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t m) {
  if (!lay || !tips || U < 1 || m < 0) {
    set_error(VBSKY_EINVAL, "bad argument to vbsky_loglik_workspace_bytes");
    return 0;
  }
  LoglikPlan pl;
  if (loglik_plan(lay, tips, U, m, &pl) != VBSKY_OK) return 0;
  return pl.total;
}

extern "C" int vbsky_loglik_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                           const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                                           const vbsky_subst_model* model, int32_t flags, double* ll, double* ll_tree,
                                           double* ll_data, vbsky_param_grads* grads, void* ws, size_t ws_bytes,
                                           void* stream) {
  VB_CHECK_ARG(lay && tips && unit_tree && in && model && ll && ws, "NULL argument");
  VB_CHECK_ARG(U >= 1, "U must be >= 1");
  const bool tree = flags & VBSKY_TERM_TREE, data = flags & VBSKY_TERM_DATA;
  // bdsky.py:282-284 shape asserts: one proportion per internal non-root node is implied by the [U][n-2] layout;
  // len(R) == len(s) == len(delta) == m is implied by the [U][m] layout.
  if (tree) VB_CHECK_ARG(m >= 1 && in->R && in->delta && in->s && in->rho, "the tree-prior term needs m >= 1 and R, delta, s, rho");
  if (grads) {
    VB_CHECK_ARG((grads->proportions || lay->n == 2) && grads->root_proportion && grads->origin && grads->clock_rate, "NULL gradient output");
    if (tree) VB_CHECK_ARG(grads->R && grads->delta && grads->s && grads->rho, "NULL skyline gradient output");
  }
  LoglikPlan pl;
  int rc = loglik_plan(lay, tips, U, m, &pl);
  if (rc != VBSKY_OK) return rc;
  if (ws_bytes < pl.total) return set_error(VBSKY_ENOMEM, "workspace too small: %zu < %zu bytes", ws_bytes, pl.total);
  char* b = (char*)ws;
  auto D = [&](size_t off) { return (double*)(b + off); };
  cudaStream_t st = (cudaStream_t)stream;
  const int n = lay->n;

  rc = vbsky_heights_forward(lay, U, unit_tree, tree ? m : 0, in, D(pl.heights), D(pl.bl), tree ? D(pl.xs) : nullptr,
                             tree ? D(pl.times) : nullptr, D(pl.lam), D(pl.psi), D(pl.mu), stream);
  if (rc != VBSKY_OK) return rc;
  if (tree) {
    rc = vbsky_bdsky_value_and_grad(U, n, m, D(pl.lam), D(pl.psi), D(pl.mu), in->rho, D(pl.xs), D(pl.times),
                                    (flags & VBSKY_CONDITION_ON_SURVIVAL) ? 1 : 0, D(pl.lt), grads ? D(pl.dlam) : nullptr,
                                    D(pl.dpsi), D(pl.dmu), grads ? grads->rho : nullptr, D(pl.dxs), D(pl.dtimes),
                                    b + pl.bdsky, pl.bdsky_bytes, stream);
    if (rc != VBSKY_OK) return rc;
  } else if (grads && grads->rho && m >= 1) {
    VB_CUDA(cudaMemsetAsync(grads->rho, 0, (size_t)U * m * sizeof(double), st));
  }
  if (data) {
    rc = (flags & VBSKY_PRUNE_F32)
             ? vbsky_prune_value_and_grad_f32(lay, tips, U, unit_tree, D(pl.bl), model, nullptr, D(pl.ld),
                                              grads ? D(pl.dbl) : nullptr, b + pl.prune, pl.prune_bytes, stream)
             : vbsky_prune_value_and_grad(lay, tips, U, unit_tree, D(pl.bl), model, nullptr, D(pl.ld),
                                          grads ? D(pl.dbl) : nullptr, b + pl.prune, pl.prune_bytes, stream);
    if (rc != VBSKY_OK) return rc;
  }
  combine_kernel<<<(U + 127) / 128, 128, 0, st>>>(U, flags, D(pl.lt), D(pl.ld), ll, ll_tree, ll_data);
  VB_CUDA(cudaGetLastError());
  if (grads) {
    vbsky_param_grads g = *grads;
    if (!tree) g.R = g.delta = g.s = nullptr;
    rc = vbsky_heights_backward(lay, U, unit_tree, tree ? m : 0, in, D(pl.heights), data ? D(pl.dbl) : nullptr,
                                tree ? D(pl.dxs) : nullptr, tree ? D(pl.dtimes) : nullptr, tree ? D(pl.dlam) : nullptr,
                                tree ? D(pl.dpsi) : nullptr, tree ? D(pl.dmu) : nullptr, &g, b + pl.dh,
                                (size_t)U * (2 * (size_t)n - 1) * sizeof(double), stream);
    if (rc != VBSKY_OK) return rc;
    if (!tree && grads->R && m >= 1) {
      VB_CUDA(cudaMemsetAsync(grads->R, 0, (size_t)U * m * sizeof(double), st));
      VB_CUDA(cudaMemsetAsync(grads->delta, 0, (size_t)U * m * sizeof(double), st));
      VB_CUDA(cudaMemsetAsync(grads->s, 0, (size_t)U * m * sizeof(double), st));
    }
  }
  return VBSKY_OK;
}
