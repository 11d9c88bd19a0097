// Felsenstein pruning log-likelihood and branch-length gradient for sm_100a.
//
// Replaces vmap(prune.prune_loglik) x counts (bdsky.py:331-343) and prune_loglik_jvp
// (prune.py:134-150) of the reference for every (tree, sample) unit at once.
//
// Design (DESIGN.md "Pruning kernel"):
//   * persistent, warp-specialised CTAs.  A work item is (unit, tile of kTile site patterns); four
//     consumer warps own two adjacent site patterns per thread for the whole tree walk (two
//     independent dependency chains per thread, 128-bit shared/global accesses), a fifth (producer)
//     warp streams everything the walk needs into shared-memory rings with cp.async.bulk (the
//     1-D TMA path) completing on mbarriers: per-step records {op, P(t) matrices}, the tip codes of
//     the step and -- in the down pass -- the children's messages from the CTA-private scratch.
//     The math loop therefore never waits on a global load;
//   * P(t) = U exp(Lambda t) U^-1 is built once per (unit, branch) by tables_kernel, in schedule
//     order, not per site as the reference does inside its vmap (substitution.py:64-70);
//   * the walk follows the layout's depth-first schedules, so the live "message" vectors
//     (clip(P(t_c) post_c), the quantity both passes actually consume) sit in a small
//     shared-memory stack; the up pass streams each internal node's message to scratch once and
//     the down pass streams it back once.  pre vectors never leave the SM.  Value-only evaluation
//     touches no scratch at all;
//   * the gradient uses the scale-free form of eq. (9),
//       g_c = (w^T Q m_c) / (w^T m_c),  w = pre_parent (.) m_sibling,
//     which equals prune.py:146-148 whenever the 1e-40 clips are inactive;
//   * two arithmetic variants, chosen per unit by tables_kernel:
//       CLIP    (some P(t) entry below 1e-20, i.e. a branch length of ~0, where the clips of
//               substitution.py:70 can be active): literal arithmetic -- clip to [1e-40, 1], divide by the
//               sum (prune.py:31-34, 70-74), scale factors accumulated as mantissa x 2^exponent;
//       no CLIP (every clipped quantity is provably >= 1e-40 and <= 1 + few ulp): the clips are
//               dropped and every rescaling is by an exact power of two (exponent arithmetic on the
//               integer pipe instead of an fp64 division), which changes no result beyond rounding;
//     either way one log per site instead of one per node (prune.py:37);
//   * branch gradients are reduced over the tile's sites through a transposed shared-memory
//     staging buffer, written as per-tile partials and summed in a fixed order by finalize_kernel
//     (bitwise reproducible).
#include <cmath>
#include <map>
#include <utility>
#include <type_traits>

#include "common.h"

namespace vbsky {

constexpr int TILE = kTile;           // site patterns per work item
constexpr int NS = 2;                 // site patterns per consumer thread (adjacent)
constexpr int CT = TILE / NS;         // consumer threads
constexpr int NCW = CT / 32;          // consumer warps
constexpr int THREADS = CT + 32;      // + producer warp
constexpr int CH = kChunkSteps;       // steps per ring stage
#ifndef VBSKY_NST
#define VBSKY_NST 3
#endif
constexpr int NST = VBSKY_NST;        // step-chunk ring stages
constexpr int NMS_MAX = 4;            // message ring slots available; prune_config uses the third when it costs no resident CTA
constexpr int CODE_ROW = TILE;        // bytes of tip codes per (step, child)
constexpr int GCH = 8;                // branch gradients staged per reduction round
constexpr int GROW = 33;              // per-warp staging row: one value per lane (its two sites pre-added) + pad
constexpr int OFF_STAGE = 128;
#ifdef VBSKY_GRAD_SHUFFLE
constexpr bool GRAD_SHUFFLE = true;   // branch gradients reduced with warp shuffles (no shared-memory staging block)
#else
constexpr bool GRAD_SHUFFLE = false;
#endif
#ifdef VBSKY_TIPQ_COMPUTE
constexpr bool TIPQ_COMPUTE = true;   // down pass: G m of a tip child computed instead of gathered from its Q table
#else
constexpr bool TIPQ_COMPUTE = false;
#endif
#ifdef VBSKY_NO_HOIST_CODES
constexpr bool HOIST_CODES = false;
#else
constexpr bool HOIST_CODES = true;   // tip codes of a step are loaded together with its op word
#endif
#ifdef VBSKY_OP_PREFETCH
constexpr bool OP_PREFETCH = true;   // the next step's op word is loaded one step ahead
#else
constexpr bool OP_PREFETCH = false;
#endif
#ifdef VBSKY_SHALLOW_MATVEC
constexpr bool SHALLOW_MATVEC = true;  // structured 4x4 product with one dependency level less and four more additions
#else
constexpr bool SHALLOW_MATVEC = false;
#endif
#ifndef VBSKY_MSG_PREFETCH
#define VBSKY_MSG_PREFETCH 0
#endif
constexpr int MSG_PREFETCH = VBSKY_MSG_PREFETCH;  // messages prefetched into L2 ahead of the ring (0 = off)
#ifndef VBSKY_MIN_CTAS
#define VBSKY_MIN_CTAS 3
#endif
constexpr int MIN_CTAS = VBSKY_MIN_CTAS;  // resident CTAs per SM the register allocation is sized for

__device__ __constant__ uint8_t kRankToMask[16] = {1, 2, 4, 8, 15, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 0};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Producer-side wait: the suspend-time hint lets the thread sleep in hardware until the phase completes
// instead of spinning in the issue slots of the consumer warp that shares its scheduler.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar),
      "r"(parity), "r"(1000000u)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Pull a range of global memory into L2 ahead of its use (no destination, no completion to wait for).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct RingState {
  uint32_t cst = 0, cph = 0;  // step-chunk ring stage / phase
  uint32_t mst = 0, mph = 0;  // message ring slot / phase
  __device__ __forceinline__ void adv_c() {
    if (++cst == NST) cst = 0, cph ^= 1;
  }
  __device__ __forceinline__ void adv_m(uint32_t nms) {
    if (++mst == nms) mst = 0, mph ^= 1;
  }
};

// Barrier slots (8 bytes each) at the start of dynamic shared memory.
__device__ __forceinline__ uint32_t bar_full_c(uint32_t base, uint32_t s) { return base + 8 * s; }
__device__ __forceinline__ uint32_t bar_empty_c(uint32_t base, uint32_t s) { return base + 8 * (NST + s); }
__device__ __forceinline__ uint32_t bar_full_m(uint32_t base, uint32_t s) { return base + 8 * (2 * NST + s); }
__device__ __forceinline__ uint32_t bar_empty_m(uint32_t base, uint32_t s) { return base + 8 * (2 * NST + NMS_MAX + s); }
__device__ __forceinline__ uint32_t bar_up_done(uint32_t base) { return base + 8 * (2 * NST + 2 * NMS_MAX); }

static thread_local int64_t g_stats[8] = {0};

// Optional CUDA-event timing of prune_kernel alone (bench.py's roofline leg): when enabled, every call
// brackets the kernel with an event pair on the launch stream; vbsky_prune_timing() drains them.
struct TimingState {
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};
static thread_local TimingState g_timing;

// ------------------------------------------------------------------------------------ fp64 path
namespace f64 {
using real = double;
using real2 = double2;
constexpr double kClipFreeMinP = 1e-20;  // every P(t) entry above this => no clip of substitution.py:70 can bind
constexpr int kPreRescaleExp = -400;
__device__ __forceinline__ real2 mk2(real a, real b) { return make_double2(a, b); }
__device__ __forceinline__ real clip_literal(real x) { return fmin(fmax(x, 1e-40), 1.0); }
// 1/x for a positive normal x: hardware seed + two Newton steps (~1 ulp), no range fix-up.
__device__ __forceinline__ real rcp_pos(real x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// Unbiased binary exponent of a positive normal value and multiplication by 2^-e on the integer pipe.
__device__ __forceinline__ int exponent_of(real x) { return (__double2hiint(x) >> 20) - 1023; }
__device__ __forceinline__ real scale_pow2(real x, int e) { return __hiloint2double(__double2hiint(x) - (e << 20), __double2loint(x)); }
// messages stay unnormalised until their magnitude has drifted below 2^-256 (a step shrinks it by <= ~2^-133)
// (the magnitude of a positive vector is judged by the largest high word of its entries: integer max, no FP64 adds)
__device__ __forceinline__ int mag_key(const real (&x)[4]) {
  return max(max(__double2hiint(x[0]), __double2hiint(x[1])), max(__double2hiint(x[2]), __double2hiint(x[3])));
}
__device__ __forceinline__ bool key_needs_rescale(int key) { return key < ((1023 - 256) << 20); }
__device__ __forceinline__ int key_exponent(int key) { return (key >> 20) - 1023; }
#include "prune_impl.inc"
}  // namespace f64

// ------------------------------------------------------------------------------------ fp32 path
// Same kernels with fp32 messages, tables, stack and scratch (half the HBM traffic, full-rate FP32 pipe);
// P(t) is still built in fp64 from expm1 and rounded once, the final log and all sums over patterns stay fp64.
// The narrow exponent range means every step rescales by a power of two (no lazy rescaling) and the clip-free
// variant requires every P(t) entry >= 1e-9.  Tolerance target: 1e-5 relative on log-likelihoods.
namespace f32 {
using real = float;
using real2 = float2;
constexpr double kClipFreeMinP = 1e-9;
constexpr int kPreRescaleExp = 1 << 30;  // always
__device__ __forceinline__ real2 mk2(real a, real b) { return make_float2(a, b); }
__device__ __forceinline__ real clip_literal(real x) { return fminf(fmaxf(x, 1e-40f), 1.0f); }
__device__ __forceinline__ real rcp_pos(real x) { return __frcp_rn(x); }
__device__ __forceinline__ int exponent_of(real x) { return (__float_as_int(x) >> 23) - 127; }
__device__ __forceinline__ real scale_pow2(real x, int e) { return __int_as_float(__float_as_int(x) - (e << 23)); }
__device__ __forceinline__ int mag_key(const real (&x)[4]) {
  return max(max(__float_as_int(x[0]), __float_as_int(x[1])), max(__float_as_int(x[2]), __float_as_int(x[3])));
}
__device__ __forceinline__ bool key_needs_rescale(int) { return true; }
__device__ __forceinline__ int key_exponent(int key) { return (key >> 23) - 127; }
#include "prune_impl.inc"
}  // namespace f32

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_prune_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t want_grad) {
  return f64::workspace_bytes(lay, tips, U, want_grad);
}
extern "C" int vbsky_prune_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                          const int32_t* unit_tree, const double* bl,
                                          const vbsky_subst_model* model, double* site_ll, double* ll,
                                          double* dll_dbl, void* ws, size_t ws_bytes, void* stream) {
  return f64::run(lay, tips, U, unit_tree, bl, model, site_ll, ll, dll_dbl, ws, ws_bytes, stream);
}
extern "C" int32_t vbsky_tips_tiles(const vbsky_tips* tips) {
  if (!tips) {
    set_error(VBSKY_EINVAL, "NULL tips");
    return -1;
  }
  return tips->tiles;
}
extern "C" int vbsky_prune_value_and_grad_tiles(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                                const int32_t* unit_tree, const double* bl, const vbsky_subst_model* model,
                                                int32_t tile_begin, int32_t tile_end, double* site_ll, double* ll,
                                                double* dll_dbl, void* ws, size_t ws_bytes, void* stream) {
  VB_CHECK_ARG(tile_begin >= 0 && tile_end > tile_begin, "bad tile range [%d, %d)", tile_begin, tile_end);
  return f64::run(lay, tips, U, unit_tree, bl, model, site_ll, ll, dll_dbl, ws, ws_bytes, stream, nullptr, tile_begin, tile_end);
}
extern "C" int vbsky_prune_spectral_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, const int32_t* unit_tree,
                                         const double* bl, const vbsky_subst_model* model, const double* g, double* ll,
                                         double* out, void* ws, size_t ws_bytes, void* stream) {
  VB_CHECK_ARG(g && out, "NULL argument");
  return f64::run(lay, tips, U, unit_tree, bl, model, nullptr, ll, out, ws, ws_bytes, stream, g);
}
extern "C" int vbsky_prune_param_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, const int32_t* unit_tree,
                                      const double* bl, const vbsky_subst_model* model, const double* dQ, double* ll,
                                      double* out, double* dpi, void* ws, size_t ws_bytes, void* stream) {
  VB_CHECK_ARG(dQ && out, "NULL argument");
  return f64::run(lay, tips, U, unit_tree, bl, model, nullptr, ll, out, ws, ws_bytes, stream, nullptr, -1, -1, dQ, dpi);
}
extern "C" size_t vbsky_prune_workspace_bytes_f32(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t want_grad) {
  return f32::workspace_bytes(lay, tips, U, want_grad);
}
extern "C" int vbsky_prune_value_and_grad_f32(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                              const int32_t* unit_tree, const double* bl,
                                              const vbsky_subst_model* model, double* site_ll, double* ll,
                                              double* dll_dbl, void* ws, size_t ws_bytes, void* stream) {
  return f32::run(lay, tips, U, unit_tree, bl, model, site_ll, ll, dll_dbl, ws, ws_bytes, stream);
}

extern "C" int vbsky_prune_last_stats(int64_t out[8]) {
  VB_CHECK_ARG(out, "NULL argument");
  for (int i = 0; i < 8; ++i) out[i] = g_stats[i];
  return VBSKY_OK;
}

extern "C" int vbsky_prune_timing(int32_t enable, double* ms_sum, int64_t* launches) {
  // Drain: waits for the recorded events, returns the summed prune_kernel time and launch count.
  double tot = 0.0;
  int64_t cnt = 0;
  for (auto& pr : g_timing.pending) {
    float ms = 0.f;
    VB_CUDA(cudaEventSynchronize(pr.second));
    VB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    tot += ms, ++cnt;
    cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
  }
  g_timing.pending.clear();
  g_timing.enabled = enable != 0;
  if (ms_sum) *ms_sum = tot;
  if (launches) *launches = cnt;
  return VBSKY_OK;
}

