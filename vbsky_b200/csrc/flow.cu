// Local variational family of the reference on the device: for every
// (tree, sample) unit, reparameterised draws of the per-tree parameters
//     x = clip(expit(mu + exp(log_sigma) * eps), 1e-7, 1 - 1e-7)
// (Transform(dim, Compose(DiagonalAffine, ZeroOne)) over a MeanField base: fasta.py:63-65,378-384,
// prob/transform.py:130-202, prob/distribution.py:52-65), their log-density log q(x) exactly as
// TransformedDistribution.log_pdf evaluates it (through the inverse map, transform.py:56-63), and the
// adjoint w.r.t. (mu, log_sigma) of a loss that depends on x and on log q (optim.py:54-65).
#include <cmath>

#include "common.h"

namespace vbsky {

struct FlowParams {
  const int32_t* unit_tree;
  const double *mu, *log_sigma;  // [T][d]
  const double* eps;             // [U][d]
  double* x;                     // [U][d]
  double* logq;                  // [U]
  const double *dx, *dlogq;      // backward inputs [U][d], [U]
  double *dmu, *dlog_sigma;      // backward outputs, per unit [U][d]
  int U, d;
};

struct FlowPoint {
  double x, logq_term, dx_dz, dlq_dmu, dlq_dls, sig_eps;
};

__device__ __forceinline__ FlowPoint flow_point(double mu, double ls, double eps) {
  const double sigma = exp(ls);
  const double z = sigma * eps + mu;           // DiagonalAffine.direct
  const double s = 1.0 / (1.0 + exp(-z));      // expit
  const double lo = 1e-7, hi = 1.0 - 1e-7;
  const double p = fmin(fmax(s, lo), hi);      // ZeroOne.direct
  const double gate = (s >= lo && s <= hi) ? 1.0 : 0.0;
  // log_pdf path: inverse map, base density, log-det-jacobian at the recovered point
  const double zr = log(p / (1.0 - p));        // ZeroOne.inverse
  const double er = (zr - mu) * exp(-ls);      // DiagonalAffine.inverse
  const double sr = 1.0 / (1.0 + exp(-(sigma * er + mu)));
  FlowPoint f;
  f.x = p;
  f.logq_term = (-0.5 * er * er - 0.9189385332046727418) - ls - (log(sr) + log1p(-sr));
  f.dx_dz = gate * s * (1.0 - s);
  // d logq / d zr = -er/sigma - (1 - 2 sr);  d zr / d z = gate * s(1-s) / (p(1-p))
  const double G = gate * s * (1.0 - s) / (p * (1.0 - p));
  const double dzr = -er / sigma - (1.0 - 2.0 * sr);
  f.dlq_dmu = dzr * G + er / sigma;
  f.dlq_dls = dzr * G * sigma * eps + er * er - 1.0;
  f.sig_eps = sigma * eps;
  return f;
}

__global__ void flow_forward_kernel(const FlowParams p) {
  // one warp per unit: lanes stride over the d dimensions, fixed-shape shuffle reduction of log q
  const int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (u >= p.U) return;
  const int t = p.unit_tree[u];
  double acc = 0.0;
  for (int j = lane; j < p.d; j += 32) {
    const FlowPoint f = flow_point(p.mu[(size_t)t * p.d + j], p.log_sigma[(size_t)t * p.d + j], p.eps[(size_t)u * p.d + j]);
    p.x[(size_t)u * p.d + j] = f.x;
    acc += f.logq_term;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) p.logq[u] = acc;
}

__global__ void flow_backward_kernel(const FlowParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.U * p.d) return;
  const int u = (int)(idx / p.d), j = (int)(idx - (long long)u * p.d);
  const int t = p.unit_tree[u];
  const FlowPoint f = flow_point(p.mu[(size_t)t * p.d + j], p.log_sigma[(size_t)t * p.d + j], p.eps[idx]);
  const double gx = p.dx ? p.dx[idx] : 0.0, gq = p.dlogq ? p.dlogq[u] : 0.0;
  p.dmu[idx] = gx * f.dx_dz + gq * f.dlq_dmu;
  p.dlog_sigma[idx] = gx * f.dx_dz * f.sig_eps + gq * f.dlq_dls;
}

}  // namespace vbsky

using namespace vbsky;

extern "C" int vbsky_local_flow_sample(int32_t U, int32_t d, const int32_t* unit_tree, const double* mu,
                                       const double* log_sigma, const double* eps, double* x, double* logq, void* stream) {
  VB_CHECK_ARG(unit_tree && mu && log_sigma && eps && x && logq, "NULL argument");
  VB_CHECK_ARG(U >= 1 && d >= 1, "bad shape U=%d d=%d", U, d);
  FlowParams p{};
  p.unit_tree = unit_tree, p.mu = mu, p.log_sigma = log_sigma, p.eps = eps, p.x = x, p.logq = logq, p.U = U, p.d = d;
  flow_forward_kernel<<<(U + 3) / 4, 128, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_local_flow_backward(int32_t U, int32_t d, const int32_t* unit_tree, const double* mu,
                                         const double* log_sigma, const double* eps, const double* dx, const double* dlogq,
                                         double* dmu, double* dlog_sigma, void* stream) {
  VB_CHECK_ARG(unit_tree && mu && log_sigma && eps && dmu && dlog_sigma, "NULL argument");
  VB_CHECK_ARG(U >= 1 && d >= 1, "bad shape U=%d d=%d", U, d);
  FlowParams p{};
  p.unit_tree = unit_tree, p.mu = mu, p.log_sigma = log_sigma, p.eps = eps, p.dx = dx, p.dlogq = dlogq;
  p.dmu = dmu, p.dlog_sigma = dlog_sigma, p.U = U, p.d = d;
  const long long tot = (long long)U * d;
  flow_backward_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

// ------------------------------------------------------------------------------------------------
// Optimiser step of SeqData.loop (fasta.py:388, 421-433): jax.example_libraries.optimizers.adagrad
// (step_size, momentum = 0.9) applied to nan_to_num(g), fused over a flat parameter block:
//   g_sq += g^2;  m = (1 - momentum) * g * (g_sq > 0 ? 1/sqrt(g_sq) : 0) + momentum * m;  x -= step_size * m.
// jax (absent from this image) is restated from its published source.
namespace vbsky {
__global__ void adagrad_kernel(long long n, double* __restrict__ x, const double* __restrict__ g, double* __restrict__ g_sq,
                               double* __restrict__ m, double step_size, double momentum) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double gi = g[i];
    // jnp.nan_to_num: nan -> 0, +-inf -> +-largest finite
    if (gi != gi) gi = 0.0;
    else if (isinf(gi)) gi = gi > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    const double acc = __dadd_rn(g_sq[i], __dmul_rn(gi, gi));
    const double inv = acc > 0 ? 1.0 / sqrt(acc) : 0.0;
    const double mi = __dadd_rn(__dmul_rn(1.0 - momentum, __dmul_rn(gi, inv)), __dmul_rn(momentum, m[i]));
    g_sq[i] = acc, m[i] = mi;
    x[i] = __dsub_rn(x[i], __dmul_rn(step_size, mi));
  }
}
}  // namespace vbsky

extern "C" int vbsky_adagrad_update(int64_t n, double* x, const double* g, double* g_sq, double* m, double step_size,
                                    double momentum, void* stream) {
  VB_CHECK_ARG(n >= 0 && (n == 0 || (x && g && g_sq && m)), "NULL argument");
  if (n == 0) return VBSKY_OK;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  vbsky::adagrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, x, g, g_sq, m, step_size, momentum);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}
