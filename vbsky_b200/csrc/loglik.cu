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
  size_t xchg_count;  // doubles of the contiguous [ld (U) | dbl (U x (2n-2))] region: what site-block ranks add up
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
  // the data term's partial results sit back to back so that one all-reduce covers them (site-pattern-block sharding)
  pl->ld = off, off += U * d;
  pl->dbl = take(U * (N - 1) * d);
  pl->xchg_count = (size_t)U + (size_t)U * (N - 1);
  pl->dxs = take(U * N * d), pl->dtimes = take(U * (mm + 1) * d);
  pl->dlam = take(U * mm * d), pl->dpsi = take(U * mm * d), pl->dmu = take(U * mm * d);
  pl->lt = take(U * d), pl->dh = take(U * N * d);
  pl->bdsky_bytes = vbsky_bdsky_workspace_bytes(U, lay->n);
  pl->bdsky = take(pl->bdsky_bytes);
  pl->prune_bytes = std::max(vbsky_prune_workspace_bytes(lay, tips, U, 1), vbsky_prune_workspace_bytes_f32(lay, tips, U, 1));
  if (pl->prune_bytes == 0) return VBSKY_EINVAL;
  pl->prune = take(pl->prune_bytes);
  pl->total = off;
  return VBSKY_OK;
}

// The tree prior (bdsky_kernel: one CTA per unit, a serial m-step recursion -- latency-bound, ~0.1-0.2 ms whatever U is)
// does not depend on the data term, so it runs on a side stream next to tables_kernel / prune_kernel: fork after the
// height chain, join before the call returns its stream to the caller.  One side stream + two events per (host thread,
// device), created on first use; event record / wait are capturable, so CUDA-graph capture of the fused call still works.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static int side_stream(int device, SideStream** out) {
  static thread_local std::vector<std::pair<int, SideStream>> pool;
  for (auto& e : pool)
    if (e.first == device) {
      *out = &e.second;
      return VBSKY_OK;
    }
  SideStream ss;
  VB_CUDA(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
  VB_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
  pool.push_back({device, ss});
  *out = &pool.back().second;
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

static int check_loglik_args(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, const int32_t* unit_tree, int32_t m,
                             const vbsky_params* in, const vbsky_subst_model* model, int32_t flags, const vbsky_param_grads* grads,
                             void* ws) {
  VB_CHECK_ARG(lay && tips && unit_tree && in && model && ws, "NULL argument");
  VB_CHECK_ARG(U >= 1, "U must be >= 1");
  const bool tree = flags & VBSKY_TERM_TREE;
  // bdsky.py:282-284 shape asserts: one proportion per internal non-root node is implied by the [U][n-2] layout;
  // len(R) == len(s) == len(delta) == m is implied by the [U][m] layout.
  if (tree) VB_CHECK_ARG(m >= 1 && in->R && in->delta && in->s && in->rho, "the tree-prior term needs m >= 1 and R, delta, s, rho");
  if (grads) {
    VB_CHECK_ARG((grads->proportions || lay->n == 2) && grads->root_proportion && grads->origin && grads->clock_rate, "NULL gradient output");
    if (tree) VB_CHECK_ARG(grads->R && grads->delta && grads->s && grads->rho, "NULL skyline gradient output");
  }
  return VBSKY_OK;
}

// First half: height chain, tree prior (+ its gradient), data term over the tiles [tile_begin, tile_end) -> partial
// [ld | dbl] in the workspace.
extern "C" int vbsky_loglik_forward_partial(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, const int32_t* unit_tree,
                                            int32_t m, const vbsky_params* in, const vbsky_subst_model* model, int32_t flags,
                                            int32_t tile_begin, int32_t tile_end, int32_t want_grad, vbsky_param_grads* grads,
                                            void* ws, size_t ws_bytes, void* stream) {
  int rc = check_loglik_args(lay, tips, U, unit_tree, m, in, model, flags, want_grad ? grads : nullptr, ws);
  if (rc != VBSKY_OK) return rc;
  if (want_grad) VB_CHECK_ARG(grads != nullptr, "grads is NULL");
  const bool tree = flags & VBSKY_TERM_TREE, data = flags & VBSKY_TERM_DATA;
  if (data && (flags & VBSKY_PRUNE_F32)) VB_CHECK_ARG(tile_begin == 0 && tile_end == tips->tiles, "the fp32 data term does not take a tile sub-range");
  LoglikPlan pl;
  rc = loglik_plan(lay, tips, U, m, &pl);
  if (rc != VBSKY_OK) return rc;
  if (ws_bytes < pl.total) return set_error(VBSKY_ENOMEM, "workspace too small: %zu < %zu bytes", ws_bytes, pl.total);
  char* b = (char*)ws;
  auto D = [&](size_t off) { return (double*)(b + off); };
  cudaStream_t st = (cudaStream_t)stream;
  const int n = lay->n;
  rc = vbsky_heights_forward(lay, U, unit_tree, tree ? m : 0, in, D(pl.heights), D(pl.bl), tree ? D(pl.xs) : nullptr,
                             tree ? D(pl.times) : nullptr, D(pl.lam), D(pl.psi), D(pl.mu), stream);
  if (rc != VBSKY_OK) return rc;
  SideStream* side = nullptr;
  if (tree && data && getenv("VBSKY_NO_SIDE_STREAM") == nullptr) {
    rc = side_stream(lay->device, &side);
    if (rc != VBSKY_OK) return rc;
    VB_CUDA(cudaEventRecord(side->fork, st));
    VB_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  if (tree) {
    rc = vbsky_bdsky_value_and_grad(U, n, m, D(pl.lam), D(pl.psi), D(pl.mu), in->rho, D(pl.xs), D(pl.times),
                                    (flags & VBSKY_CONDITION_ON_SURVIVAL) ? 1 : 0, D(pl.lt), want_grad ? D(pl.dlam) : nullptr,
                                    D(pl.dpsi), D(pl.dmu), want_grad ? grads->rho : nullptr, D(pl.dxs), D(pl.dtimes),
                                    b + pl.bdsky, pl.bdsky_bytes, side ? (void*)side->stream : stream);
    if (side) {  // join is enqueued after the data term below; record it now so that an early return cannot leave a dangling fork
      cudaError_t e = cudaEventRecord(side->join, side->stream);
      if (rc == VBSKY_OK && e != cudaSuccess) rc = set_error(VBSKY_ECUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    }
    if (rc != VBSKY_OK) {
      if (side) cudaStreamWaitEvent(st, side->join, 0);
      return rc;
    }
  } else if (want_grad && grads->rho && m >= 1) {
    VB_CUDA(cudaMemsetAsync(grads->rho, 0, (size_t)U * m * sizeof(double), st));
  }
  if (data) {
    rc = (flags & VBSKY_PRUNE_F32)
             ? vbsky_prune_value_and_grad_f32(lay, tips, U, unit_tree, D(pl.bl), model, nullptr, D(pl.ld),
                                              want_grad ? D(pl.dbl) : nullptr, b + pl.prune, pl.prune_bytes, stream)
             : vbsky_prune_value_and_grad_tiles(lay, tips, U, unit_tree, D(pl.bl), model, tile_begin, tile_end, nullptr, D(pl.ld),
                                                want_grad ? D(pl.dbl) : nullptr, b + pl.prune, pl.prune_bytes, stream);
  }
  if (side) VB_CUDA(cudaStreamWaitEvent(st, side->join, 0));  // the caller's stream owns every result from here on
  return rc;
}

extern "C" double* vbsky_loglik_exchange_buffer(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t m, void* ws,
                                                int64_t* count) {
  LoglikPlan pl;
  if (!lay || !tips || !ws || U < 1 || loglik_plan(lay, tips, U, m, &pl) != VBSKY_OK) {
    set_error(VBSKY_EINVAL, "bad argument to vbsky_loglik_exchange_buffer");
    return nullptr;
  }
  if (count) *count = (int64_t)pl.xchg_count;
  return (double*)((char*)ws + pl.ld);
}

// Second half: combine the terms and run the adjoint of the height chain on the (complete) [ld | dbl].
extern "C" int vbsky_loglik_backward_complete(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, const int32_t* unit_tree,
                                              int32_t m, const vbsky_params* in, const vbsky_subst_model* model, int32_t flags,
                                              double* ll, double* ll_tree, double* ll_data, vbsky_param_grads* grads, void* ws,
                                              size_t ws_bytes, void* stream) {
  int rc = check_loglik_args(lay, tips, U, unit_tree, m, in, model, flags, grads, ws);
  if (rc != VBSKY_OK) return rc;
  VB_CHECK_ARG(ll != nullptr, "ll is NULL");
  const bool tree = flags & VBSKY_TERM_TREE, data = flags & VBSKY_TERM_DATA;
  LoglikPlan pl;
  rc = loglik_plan(lay, tips, U, m, &pl);
  if (rc != VBSKY_OK) return rc;
  if (ws_bytes < pl.total) return set_error(VBSKY_ENOMEM, "workspace too small: %zu < %zu bytes", ws_bytes, pl.total);
  char* b = (char*)ws;
  auto D = [&](size_t off) { return (double*)(b + off); };
  cudaStream_t st = (cudaStream_t)stream;
  const int n = lay->n;
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

extern "C" int vbsky_loglik_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                           const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                                           const vbsky_subst_model* model, int32_t flags, double* ll, double* ll_tree,
                                           double* ll_data, vbsky_param_grads* grads, void* ws, size_t ws_bytes,
                                           void* stream) {
  VB_CHECK_ARG(ll != nullptr, "NULL argument");
  int rc = vbsky_loglik_forward_partial(lay, tips, U, unit_tree, m, in, model, flags, 0, tips ? tips->tiles : 0, grads != nullptr,
                                        grads, ws, ws_bytes, stream);
  if (rc != VBSKY_OK) return rc;
  return vbsky_loglik_backward_complete(lay, tips, U, unit_tree, m, in, model, flags, ll, ll_tree, ll_data, grads, ws, ws_bytes, stream);
}
