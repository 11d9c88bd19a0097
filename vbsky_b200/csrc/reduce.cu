// The per-evaluation reduction of the ensemble ELBO: one kernel turns the per-unit outputs of
// vbsky_loglik_value_and_grad into the all-reduce payload [ELBO, d/d global parameters, d/d per-tree parameters],
// with a fixed summation order (bitwise reproducible for a given unit partition).
#include "common.h"

namespace vbsky {

struct ReduceParams {
  const int32_t* unit_tree;
  const double* ll;
  const double* g_prop;   // [U][n-2]
  const double* g_rprop;  // [U]
  const double* g_origin;
  const double* g_clock;  // [U]
  const double* g_sky[4];  // R, delta, s, rho: [U][m] (may be null => zeros)
  double* out;
  double scale;
  int U, T, n, m;
};

// Blocks 0..T-1: tree t sums d/d(proportions, root_proportion) over its units (unit_tree is non-decreasing, so they are
// the contiguous range found by binary search), one thread per column, units in index order.
// Blocks T..: one warp per global column (ll, origin, clock_rate, R[m], delta[m], s[m], rho[m]); lanes stride the units,
// then a fixed butterfly.
__global__ void reduce_elbo_kernel(const ReduceParams p) {
  const int d = p.n - 1;
  if ((int)blockIdx.x < p.T) {
    const int t = blockIdx.x;
    int lo = 0, hi = p.U;
    while (lo < hi) {  // first u with unit_tree[u] >= t
      const int mid = (lo + hi) >> 1;
      if (p.unit_tree[mid] < t) lo = mid + 1; else hi = mid;
    }
    const int first = lo;
    hi = p.U;
    while (lo < hi) {  // first u with unit_tree[u] > t
      const int mid = (lo + hi) >> 1;
      if (p.unit_tree[mid] <= t) lo = mid + 1; else hi = mid;
    }
    const int last = lo;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      double acc = 0.0;
      if (j < p.n - 2) {
        for (int u = first; u < last; ++u) acc += p.g_prop[(size_t)u * (p.n - 2) + j];
      } else {
        for (int u = first; u < last; ++u) acc += p.g_rprop[u];
      }
      p.out[3 + 4 * p.m + (size_t)t * d + j] = acc * p.scale;
    }
    return;
  }
  const int warps = blockDim.x >> 5;
  const int col = ((int)blockIdx.x - p.T) * warps + (threadIdx.x >> 5);
  const int ncol = 3 + 4 * p.m;
  if (col >= ncol) return;
  const int lane = threadIdx.x & 31;
  const double* src;
  int stride = 1, off = 0;
  if (col == 0) src = p.ll;
  else if (col == 1) src = p.g_origin;
  else if (col == 2) src = p.g_clock;
  else {
    src = p.g_sky[(col - 3) / p.m];
    stride = p.m, off = (col - 3) % p.m;
  }
  double acc = 0.0;
  if (src != nullptr)
    for (int u = lane; u < p.U; u += 32) acc += src[(size_t)u * stride + off];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) p.out[col] = acc * p.scale;
}

}  // namespace vbsky

using namespace vbsky;

extern "C" int64_t vbsky_reduce_elbo_count(int32_t T, int32_t n, int32_t m) { return 3 + 4 * (int64_t)m + (int64_t)T * (n - 1); }

extern "C" int vbsky_reduce_elbo(int32_t U, int32_t T, int32_t n, int32_t m, const int32_t* unit_tree, const double* ll,
                                 const vbsky_param_grads* grads, double scale, double* out, void* stream) {
  VB_CHECK_ARG(unit_tree && ll && grads && out, "NULL argument");
  VB_CHECK_ARG(U >= 1 && T >= 1 && n >= 2 && m >= 0, "bad sizes U=%d T=%d n=%d m=%d", U, T, n, m);
  VB_CHECK_ARG((grads->proportions || n == 2) && grads->root_proportion && grads->origin && grads->clock_rate, "NULL gradient input");
  ReduceParams p;
  p.unit_tree = unit_tree, p.ll = ll, p.g_prop = grads->proportions, p.g_rprop = grads->root_proportion;
  p.g_origin = grads->origin, p.g_clock = grads->clock_rate;
  p.g_sky[0] = grads->R, p.g_sky[1] = grads->delta, p.g_sky[2] = grads->s, p.g_sky[3] = grads->rho;
  p.out = out, p.scale = scale, p.U = U, p.T = T, p.n = n, p.m = m;
  const int warps = 4;
  const int gblocks = (3 + 4 * m + warps - 1) / warps;
  reduce_elbo_kernel<<<T + gblocks, 32 * warps, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}
