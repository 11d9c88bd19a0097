// XLA FFI handlers: the C ABI of include/vbsky_b200.h exposed as jax.ffi custom calls (CUDA platform).
//
// One handler per reference callee that the new kernels replace:
//   VbskyPrune           prune.prune_loglik vmapped over sites x counts + prune_loglik_jvp   (prune.py:99-150, bdsky.py:331-343)
//   VbskyBdsky           bdsky._tree_loglik and its gradient                                 (bdsky.py:18-216)
//   VbskyHeightsForward  TreeData.height_transform + the glue of bdsky.loglik                (tree_data.py:287-367, bdsky.py:292-322)
//   VbskyHeightsBackward its adjoint
//   VbskyLoglik          all of bdsky.loglik for a batch of samples, value + every gradient  (bdsky.py:268-345)
//   VbskyReduceElbo      the all-reduce payload of one ensemble evaluation
// XLA calls a handler on its own thread with device buffers and its CUDA stream; a handler only enqueues work on that
// stream and returns an ffi::Error (the C ABI never throws, never synchronises).  Handles (vbsky_layout*, vbsky_tips*)
// are created once on the Python side through ctypes and passed as int64 attributes; the substitution model travels
// as a 40-double operand [pi(4), P(16), d(4), Pinv(16)] (the SubstitutionModel pytree of substitution.py:51-55).
// Workspaces are extra result buffers sized by the *_workspace_bytes queries, so XLA owns every allocation.
//
// Built only where jaxlib's headers exist:   make -C vbsky_b200/csrc ffi XLA_FFI_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
// (JAX is not installable in this image; tests/test_ffi_source.py checks this file against the header instead.)
#include <cstdint>
#include <cstring>

#include <cuda_runtime_api.h>

#include "vbsky_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

ffi::Error Status(int rc) {
  if (rc == VBSKY_OK) return ffi::Error::Success();
  return ffi::Error(rc == VBSKY_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, vbsky_last_error());
}

const vbsky_layout* Layout(int64_t h) { return reinterpret_cast<const vbsky_layout*>(static_cast<intptr_t>(h)); }
const vbsky_tips* Tips(int64_t h) { return reinterpret_cast<const vbsky_tips*>(static_cast<intptr_t>(h)); }

// The 40-double model operand lives on the device (it is a jnp array); the C ABI takes the struct by host pointer.
// It is tiny and constant across a fit, so the Python side also passes it as four attribute spans.
vbsky_subst_model Model(ffi::Span<const double> pi, ffi::Span<const double> P, ffi::Span<const double> d, ffi::Span<const double> Pinv) {
  vbsky_subst_model m;
  std::memcpy(m.pi, pi.begin(), 4 * sizeof(double));
  std::memcpy(m.P, P.begin(), 16 * sizeof(double));
  std::memcpy(m.d, d.begin(), 4 * sizeof(double));
  std::memcpy(m.Pinv, Pinv.begin(), 16 * sizeof(double));
  return m;
}

// ---------------------------------------------------------------------------------------------------------- prune
ffi::Error PruneImpl(cudaStream_t stream, int64_t layout, int64_t tips, ffi::Span<const double> pi, ffi::Span<const double> P,
                     ffi::Span<const double> d, ffi::Span<const double> Pinv, S32 unit_tree, F64 bl, RF64 ll, RF64 dll, RU8 ws) {
  const int32_t U = static_cast<int32_t>(unit_tree.element_count());
  const vbsky_subst_model model = Model(pi, P, d, Pinv);
  return Status(vbsky_prune_value_and_grad(Layout(layout), Tips(tips), U, unit_tree.typed_data(), bl.typed_data(), &model,
                                           /*site_ll=*/nullptr, ll->typed_data(), dll->typed_data(), ws->untyped_data(),
                                           ws->size_bytes(), stream));
}

// ---------------------------------------------------------------------------------------------------------- bdsky
ffi::Error BdskyImpl(cudaStream_t stream, int32_t n, int32_t condition_on_survival, F64 lam, F64 psi, F64 mu, F64 rho, F64 xs,
                     F64 times, RF64 lp, RF64 dlam, RF64 dpsi, RF64 dmu, RF64 drho, RF64 dxs, RF64 dtimes, RU8 ws) {
  const auto dims = lam.dimensions();
  const int32_t U = static_cast<int32_t>(dims[0]), m = static_cast<int32_t>(dims[1]);
  return Status(vbsky_bdsky_value_and_grad(U, n, m, lam.typed_data(), psi.typed_data(), mu.typed_data(), rho.typed_data(),
                                           xs.typed_data(), times.typed_data(), condition_on_survival, lp->typed_data(),
                                           dlam->typed_data(), dpsi->typed_data(), dmu->typed_data(), drho->typed_data(),
                                           dxs->typed_data(), dtimes->typed_data(), ws->untyped_data(), ws->size_bytes(), stream));
}

// --------------------------------------------------------------------------------------------------- height chain
vbsky_params Params(F64& proportions, F64& root_proportion, F64& origin, F64& origin_start, F64& clock_rate, F64& R, F64& delta,
                    F64& s, F64& rho) {
  vbsky_params p{};  // grid = NULL: equidistant skyline intervals (the reference's default, bdsky.py:314)
  p.proportions = proportions.typed_data(), p.root_proportion = root_proportion.typed_data();
  p.origin = origin.typed_data(), p.origin_start = origin_start.typed_data(), p.clock_rate = clock_rate.typed_data();
  p.R = R.typed_data(), p.delta = delta.typed_data(), p.s = s.typed_data(), p.rho = rho.typed_data();
  return p;
}

ffi::Error HeightsForwardImpl(cudaStream_t stream, int64_t layout, S32 unit_tree, F64 proportions, F64 root_proportion, F64 origin,
                              F64 origin_start, F64 clock_rate, F64 R, F64 delta, F64 s, F64 rho, RF64 heights, RF64 bl_scaled,
                              RF64 xs, RF64 times, RF64 lam, RF64 psi, RF64 mu) {
  const int32_t U = static_cast<int32_t>(unit_tree.element_count());
  const int32_t m = static_cast<int32_t>(R.dimensions()[1]);
  const vbsky_params in = Params(proportions, root_proportion, origin, origin_start, clock_rate, R, delta, s, rho);
  return Status(vbsky_heights_forward(Layout(layout), U, unit_tree.typed_data(), m, &in, heights->typed_data(),
                                      bl_scaled->typed_data(), xs->typed_data(), times->typed_data(), lam->typed_data(),
                                      psi->typed_data(), mu->typed_data(), stream));
}

ffi::Error HeightsBackwardImpl(cudaStream_t stream, int64_t layout, S32 unit_tree, F64 proportions, F64 root_proportion, F64 origin,
                               F64 origin_start, F64 clock_rate, F64 R, F64 delta, F64 s, F64 rho, F64 heights, F64 d_bl_scaled,
                               F64 d_xs, F64 d_times, F64 d_lam, F64 d_psi, F64 d_mu, RF64 g_proportions, RF64 g_root_proportion,
                               RF64 g_origin, RF64 g_clock_rate, RF64 g_R, RF64 g_delta, RF64 g_s, RU8 ws) {
  const int32_t U = static_cast<int32_t>(unit_tree.element_count());
  const int32_t m = static_cast<int32_t>(R.dimensions()[1]);
  const vbsky_params in = Params(proportions, root_proportion, origin, origin_start, clock_rate, R, delta, s, rho);
  vbsky_param_grads out{};
  out.proportions = g_proportions->typed_data(), out.root_proportion = g_root_proportion->typed_data();
  out.origin = g_origin->typed_data(), out.clock_rate = g_clock_rate->typed_data();
  out.R = g_R->typed_data(), out.delta = g_delta->typed_data(), out.s = g_s->typed_data(), out.rho = nullptr;
  return Status(vbsky_heights_backward(Layout(layout), U, unit_tree.typed_data(), m, &in, heights.typed_data(),
                                       d_bl_scaled.typed_data(), d_xs.typed_data(), d_times.typed_data(), d_lam.typed_data(),
                                       d_psi.typed_data(), d_mu.typed_data(), &out, ws->untyped_data(), ws->size_bytes(), stream));
}

// --------------------------------------------------------------------------------------------------- fused loglik
ffi::Error LoglikImpl(cudaStream_t stream, int64_t layout, int64_t tips, int32_t flags, ffi::Span<const double> pi,
                      ffi::Span<const double> P, ffi::Span<const double> d, ffi::Span<const double> Pinv, S32 unit_tree,
                      F64 proportions, F64 root_proportion, F64 origin, F64 origin_start, F64 clock_rate, F64 R, F64 delta, F64 s,
                      F64 rho, RF64 ll, RF64 ll_tree, RF64 ll_data, RF64 g_proportions, RF64 g_root_proportion, RF64 g_origin,
                      RF64 g_clock_rate, RF64 g_R, RF64 g_delta, RF64 g_s, RF64 g_rho, RU8 ws) {
  const int32_t U = static_cast<int32_t>(unit_tree.element_count());
  const int32_t m = static_cast<int32_t>(R.dimensions()[1]);
  const vbsky_params in = Params(proportions, root_proportion, origin, origin_start, clock_rate, R, delta, s, rho);
  const vbsky_subst_model model = Model(pi, P, d, Pinv);
  vbsky_param_grads g{};
  g.proportions = g_proportions->typed_data(), g.root_proportion = g_root_proportion->typed_data();
  g.origin = g_origin->typed_data(), g.clock_rate = g_clock_rate->typed_data();
  g.R = g_R->typed_data(), g.delta = g_delta->typed_data(), g.s = g_s->typed_data(), g.rho = g_rho->typed_data();
  return Status(vbsky_loglik_value_and_grad(Layout(layout), Tips(tips), U, unit_tree.typed_data(), m, &in, &model, flags,
                                            ll->typed_data(), ll_tree->typed_data(), ll_data->typed_data(), &g,
                                            ws->untyped_data(), ws->size_bytes(), stream));
}

// ------------------------------------------------------------------------------------------------ ELBO reduction
ffi::Error ReduceElboImpl(cudaStream_t stream, int32_t T, int32_t n, double scale, S32 unit_tree, F64 ll, F64 g_proportions,
                          F64 g_root_proportion, F64 g_origin, F64 g_clock_rate, F64 g_R, F64 g_delta, F64 g_s, F64 g_rho,
                          RF64 out) {
  const int32_t U = static_cast<int32_t>(unit_tree.element_count());
  const int32_t m = static_cast<int32_t>(g_R.dimensions()[1]);
  vbsky_param_grads g{};
  g.proportions = g_proportions.typed_data(), g.root_proportion = g_root_proportion.typed_data();
  g.origin = g_origin.typed_data(), g.clock_rate = g_clock_rate.typed_data();
  g.R = g_R.typed_data(), g.delta = g_delta.typed_data(), g.s = g_s.typed_data(), g.rho = g_rho.typed_data();
  return Status(vbsky_reduce_elbo(U, T, n, m, unit_tree.typed_data(), ll.typed_data(), &g, scale, out->typed_data(), stream));
}

}  // namespace

#define VB_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define VB_MODEL_ATTRS .Attr<ffi::Span<const double>>("pi").Attr<ffi::Span<const double>>("P").Attr<ffi::Span<const double>>("d").Attr<ffi::Span<const double>>("Pinv")
#define VB_PARAM_ARGS .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyPrune, PruneImpl,
                              VB_STREAM.Attr<int64_t>("layout").Attr<int64_t>("tips") VB_MODEL_ATTRS
                                  .Arg<S32>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyBdsky, BdskyImpl,
                              VB_STREAM.Attr<int32_t>("n").Attr<int32_t>("condition_on_survival")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyHeightsForward, HeightsForwardImpl,
                              VB_STREAM.Attr<int64_t>("layout").Arg<S32>() VB_PARAM_ARGS
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyHeightsBackward, HeightsBackwardImpl,
                              VB_STREAM.Attr<int64_t>("layout").Arg<S32>() VB_PARAM_ARGS
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyLoglik, LoglikImpl,
                              VB_STREAM.Attr<int64_t>("layout").Attr<int64_t>("tips").Attr<int32_t>("flags") VB_MODEL_ATTRS
                                  .Arg<S32>() VB_PARAM_ARGS
                                  .Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(VbskyReduceElbo, ReduceElboImpl,
                              VB_STREAM.Attr<int32_t>("T").Attr<int32_t>("n").Attr<double>("scale")
                                  .Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>());
