// Host-buffer entry points: what a caller holding NumPy arrays binds (the reference's jitted `step`
// receives its arguments that way, fasta.py:461-471).  A session owns device staging buffers and
// the workspace; each call copies inputs host -> device, runs the device entry point, copies the
// results back and synchronises the stream.
#include "common.h"

using namespace vbsky;

struct vbsky_session {
  const vbsky_layout* lay = nullptr;
  const vbsky_tips* tips = nullptr;
  int32_t U_max = 0, m_max = 0;
  int32_t* d_unit_tree = nullptr;
  double* d_in = nullptr;    // packed parameter rows
  double* d_out = nullptr;   // packed results
  void* d_ws = nullptr;
  size_t ws_bytes = 0, in_doubles = 0, out_doubles = 0;
};

extern "C" void vbsky_session_destroy(vbsky_session* s) {
  if (!s) return;
  cudaFree(s->d_unit_tree), cudaFree(s->d_in), cudaFree(s->d_out), cudaFree(s->d_ws);
  delete s;
}

extern "C" int vbsky_session_create(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U_max, int32_t m_max,
                                    vbsky_session** out) {
  VB_CHECK_ARG(lay && tips && out, "NULL argument");
  VB_CHECK_ARG(U_max >= 1 && m_max >= 0, "bad capacity U_max=%d m_max=%d", U_max, m_max);
  VB_CHECK_DEVICE(lay);
  *out = nullptr;
  auto* s = new vbsky_session();
  s->lay = lay, s->tips = tips, s->U_max = U_max, s->m_max = m_max;
  const size_t n = lay->n, nb = 2 * n - 2, mm = (size_t)std::max(m_max, 1);
  // inputs: max(bl rows, parameter rows); outputs: ll x3 + gradients
  s->in_doubles = (size_t)U_max * std::max(nb, (n - 2) + 4 + 4 * mm + (mm + 1));
  s->out_doubles = (size_t)U_max * (3 + std::max(nb, (n - 2) + 3 + 4 * mm + (mm + 1)));
  s->ws_bytes = std::max(vbsky_prune_workspace_bytes(lay, tips, U_max, 1), vbsky_loglik_workspace_bytes(lay, tips, U_max, m_max));
  if (s->ws_bytes == 0) {
    delete s;
    return VBSKY_EINVAL;
  }
  cudaError_t e = cudaMalloc((void**)&s->d_unit_tree, U_max * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_in, s->in_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_out, s->out_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_ws, s->ws_bytes);
  if (e != cudaSuccess) {
    vbsky_session_destroy(s);
    return set_error(VBSKY_ENOMEM, "session allocation failed: %s", cudaGetErrorString(e));
  }
  *out = s;
  return VBSKY_OK;
}

extern "C" int vbsky_prune_value_and_grad_host(vbsky_session* s, int32_t U, const int32_t* unit_tree, const double* bl,
                                               const vbsky_subst_model* model, double* ll, double* dll_dbl, void* stream_) {
  VB_CHECK_ARG(s && unit_tree && bl && model && ll, "NULL argument");
  VB_CHECK_ARG(U >= 1 && U <= s->U_max, "U=%d outside the session's capacity %d", U, s->U_max);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nb = 2 * (size_t)s->lay->n - 2;
  double* d_ll = s->d_out;
  double* d_dll = s->d_out + (size_t)3 * s->U_max;
  VB_CUDA(cudaMemcpyAsync(s->d_unit_tree, unit_tree, U * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  VB_CUDA(cudaMemcpyAsync(s->d_in, bl, U * nb * sizeof(double), cudaMemcpyHostToDevice, stream));
  int rc = vbsky_prune_value_and_grad(s->lay, s->tips, U, s->d_unit_tree, s->d_in, model, nullptr, d_ll,
                                      dll_dbl ? d_dll : nullptr, s->d_ws, s->ws_bytes, stream_);
  if (rc != VBSKY_OK) return rc;
  VB_CUDA(cudaMemcpyAsync(ll, d_ll, U * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (dll_dbl) VB_CUDA(cudaMemcpyAsync(dll_dbl, d_dll, U * nb * sizeof(double), cudaMemcpyDeviceToHost, stream));
  VB_CUDA(cudaStreamSynchronize(stream));
  return VBSKY_OK;
}

extern "C" int vbsky_loglik_value_and_grad_host(vbsky_session* s, int32_t U, const int32_t* unit_tree, int32_t m,
                                                const vbsky_params* in, const vbsky_subst_model* model, int32_t flags,
                                                double* ll, double* ll_tree, double* ll_data, vbsky_param_grads* grads,
                                                void* stream_) {
  VB_CHECK_ARG(s && unit_tree && in && model && ll, "NULL argument");
  VB_CHECK_ARG(U >= 1 && U <= s->U_max && m >= 0 && m <= s->m_max, "U=%d, m=%d outside the session's capacity (%d, %d)", U, m, s->U_max, s->m_max);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t n = s->lay->n, np = n - 2, mm = (size_t)m;
  // device-side packing: [proportions U*(n-2)] [root_prop U] [origin U] [origin_start U] [clock U] [R|delta|s|rho U*m each] [grid U*(m+1)]
  double* d = s->d_in;
  vbsky_params din{};
  struct Field { const double* src; const double** dst; size_t cols; };
  Field f[10] = {{in->proportions, &din.proportions, np}, {in->root_proportion, &din.root_proportion, 1},
                 {in->origin, &din.origin, 1},           {in->origin_start, &din.origin_start, 1},
                 {in->clock_rate, &din.clock_rate, 1},   {in->R, &din.R, mm},
                 {in->delta, &din.delta, mm},            {in->s, &din.s, mm},
                 {in->rho, &din.rho, mm},                {in->grid, &din.grid, mm ? mm + 1 : 0}};
  for (auto& x : f) {
    if (x.src && x.cols) {
      VB_CUDA(cudaMemcpyAsync(d, x.src, (size_t)U * x.cols * sizeof(double), cudaMemcpyHostToDevice, stream));
      *x.dst = d;
      d += (size_t)U * x.cols;
    }
  }
  VB_CUDA(cudaMemcpyAsync(s->d_unit_tree, unit_tree, U * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  double* o = s->d_out;
  double *d_ll = o, *d_lt = o + U, *d_ld = o + 2 * (size_t)U;
  o += 3 * (size_t)s->U_max;
  vbsky_param_grads dg{};
  struct GField { double* host; double** dev; size_t cols; };
  GField g[9];
  if (grads) {
    GField gg[9] = {{grads->proportions, &dg.proportions, np}, {grads->root_proportion, &dg.root_proportion, 1},
                    {grads->origin, &dg.origin, 1},           {grads->clock_rate, &dg.clock_rate, 1},
                    {grads->R, &dg.R, mm},                    {grads->delta, &dg.delta, mm},
                    {grads->s, &dg.s, mm},                    {grads->rho, &dg.rho, mm},
                    {in->grid ? grads->grid : nullptr, &dg.grid, mm ? mm + 1 : 0}};
    for (int i = 0; i < 9; ++i) {
      g[i] = gg[i];
      if (g[i].host && g[i].cols) {
        *g[i].dev = o;
        o += (size_t)U * g[i].cols;
      }
    }
  }
  int rc = vbsky_loglik_value_and_grad(s->lay, s->tips, U, s->d_unit_tree, m, &din, model, flags, d_ll, d_lt, d_ld,
                                       grads ? &dg : nullptr, s->d_ws, s->ws_bytes, stream_);
  if (rc != VBSKY_OK) return rc;
  VB_CUDA(cudaMemcpyAsync(ll, d_ll, U * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (ll_tree) VB_CUDA(cudaMemcpyAsync(ll_tree, d_lt, U * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (ll_data) VB_CUDA(cudaMemcpyAsync(ll_data, d_ld, U * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (grads)
    for (int i = 0; i < 9; ++i)
      if (g[i].host && g[i].cols)
        VB_CUDA(cudaMemcpyAsync(g[i].host, *g[i].dev, (size_t)U * g[i].cols * sizeof(double), cudaMemcpyDeviceToHost, stream));
  VB_CUDA(cudaStreamSynchronize(stream));
  return VBSKY_OK;
}
