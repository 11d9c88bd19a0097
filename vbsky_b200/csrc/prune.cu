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

#include "common.h"

namespace vbsky {

constexpr int TILE = kTile;           // site patterns per work item
constexpr int NS = 2;                 // site patterns per consumer thread (adjacent)
constexpr int CT = TILE / NS;         // consumer threads
constexpr int NCW = CT / 32;          // consumer warps
constexpr int THREADS = CT + 32;      // + producer warp
constexpr int CH = kChunkSteps;       // steps per ring stage
constexpr int NST = 2;                // step-chunk ring stages
constexpr int NMS_MAX = 2;            // message ring slots (a deeper ring measured slower: it only shrinks L1)
constexpr int TIPM = 20;              // doubles of a tip-branch table: columns of P(t) for A,C,G,T and their sum (N)
constexpr int UPREC = 2 + 2 * TIPM + 16;  // up-step record:   op(2) Ta(20) Tb(20) Pp(16)
constexpr int DNREC = 2 + 4 * TIPM;       // down-step record: op(2) Ta|Pa(20) Tb|Pb(20) QTa(20) QTb(20)
constexpr int UP_A = 2, UP_B = 2 + TIPM, UP_P = 2 + 2 * TIPM;
constexpr int DN_A = 2, DN_B = 2 + TIPM, DN_QA = 2 + 2 * TIPM, DN_QB = 2 + 3 * TIPM;
constexpr int CODE_ROW = TILE;        // bytes of tip codes per (step, child)
constexpr int STAGE_BYTES = CH * DNREC * 8 + 2 * CH * CODE_ROW;
constexpr int STAGE_CODES = CH * DNREC * 8;  // byte offset of the code rows inside a stage
constexpr int MSG_BYTES = 4 * TILE * 8;
constexpr int GCH = 8;                // branch gradients staged per reduction round
constexpr int GROW = 33;              // per-warp staging row: one value per lane (its two sites pre-added) + pad
constexpr int OFF_STAGE = 128;
constexpr int OFF_MSG = OFF_STAGE + ((NST * STAGE_BYTES + 127) / 128) * 128;

static_assert(NS == 2, "the consumer code is written for two sites per thread");
static_assert(STAGE_BYTES % 16 == 0 && (UPREC * 8) % 16 == 0 && (DNREC * 8) % 16 == 0, "bulk copies need 16-byte granularity");
static_assert(GCH == 8, "the per-warp gradient reduction assumes 8 staged branches x 4 lane groups");

struct PruneParams {
  const double* uptab;    // [U][n-1][UPREC]
  const double* dntab;    // [U][n-1][DNREC]
  const uint32_t* flags;  // [U] bit0: clips may be active
  const uint8_t* codes_up;
  const uint8_t* codes_dn;  // [T][tiles][2(n-1)][TILE]
  const double* counts;     // [T][S_pad]
  const int32_t* unit_tree;
  const int32_t* msg_rows;   // [T][max(n-2,1)]
  const int32_t* msg_start;  // [T][nchunks+1]
  double* scratch;           // [grid][n-1][4][TILE]
  double* site_ll;           // [U][S] or null
  double* ll_part;           // [U][tiles]
  double* g_part;            // [U][tiles][2n-2]
  double Q[16];
  double pi[4];
  int n, U, tiles, depth, nms;
  long long S, S_pad;
};

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

template <bool CLIP>
__device__ __forceinline__ double clipv(double x) {
  // jnp.clip(ret, 1e-40, 1.0) of substitution.py:70
  return CLIP ? fmin(fmax(x, 1e-40), 1.0) : x;
}

// 1/x for a positive normal x: hardware seed + two Newton steps (~1 ulp), no range fix-up.
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
// Unbiased binary exponent of a positive normal double and multiplication by 2^-e on the integer pipe.
__device__ __forceinline__ int exponent_of(double x) { return (__double2hiint(x) >> 20) - 1023; }
__device__ __forceinline__ double scale_pow2(double x, int e) {
  return __hiloint2double(__double2hiint(x) - (e << 20), __double2loint(x));
}

__device__ __forceinline__ void lds16(const double* p, double (&M)[16]) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double2 v = q[i];
    M[2 * i] = v.x, M[2 * i + 1] = v.y;
  }
}

// M e_mask for a tip's state set (substitution.py:10-35 partials).  Tip tables hold the columns of the
// matrix contiguously (T[r*4 + i] = M[i][r]) plus their sum as a fifth column, so ranks 0..4 (A, C, G, T and
// the fully ambiguous N) are one 32-byte gather; the other ambiguity codes add up their columns.
__device__ __forceinline__ void tip_columns(const double* T, int rank, double (&out)[4]) {
  if (rank <= 4) {
    const double2 a = *reinterpret_cast<const double2*>(T + rank * 4);
    const double2 b = *reinterpret_cast<const double2*>(T + rank * 4 + 2);
    out[0] = a.x, out[1] = a.y, out[2] = b.x, out[3] = b.y;
  } else {
    const int mask = kRankToMask[rank];
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((mask >> j) & 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] += T[j * 4 + i];
      }
  }
}

// rows [4][TILE] of doubles: this thread's two adjacent sites of row i
__device__ __forceinline__ void load_rows(const double* base, double (&x)[NS][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 v = *reinterpret_cast<const double2*>(base + i * TILE);
    x[0][i] = v.x, x[1][i] = v.y;
  }
}
__device__ __forceinline__ void store_rows(double* base, const double (&x)[NS][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<double2*>(base + i * TILE) = make_double2(x[0][i], x[1][i]);
}

__device__ __forceinline__ void copy_rows(const double (&src)[NS][4], double (&dst)[NS][4]) {
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[s][i] = src[s][i];
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

// Message of a tip child for this thread's two sites (+ optionally Q P(t) e_mask for the gradient).
template <bool CLIP, bool WITH_Q>
__device__ __forceinline__ void tip_messages(const double* P, const double* QP, const unsigned char* crow, double (&m)[NS][4],
                                             double (&q)[NS][4]) {
  const uchar2 rr = *reinterpret_cast<const uchar2*>(crow);
  tip_columns(P, rr.x, m[0]);
  tip_columns(P, rr.y, m[1]);
  if (WITH_Q) {
    tip_columns(QP, rr.x, q[0]);
    tip_columns(QP, rr.y, q[1]);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int i = 0; i < 4; ++i) m[s][i] = clipv<CLIP>(m[s][i]);
}

// One work item for the four consumer warps.
template <bool WANT_GRAD, bool CLIP>
__device__ __forceinline__ void consume_item(const PruneParams& p, unsigned char* smem, RingState& rs, int u, int tile,
                                             int t, double* myscr) {
  const uint32_t bars = smem_u32(smem);
  double* stk = reinterpret_cast<double*>(smem + OFF_MSG + p.nms * MSG_BYTES);  // [depth][4][TILE]
  double* gst = stk + (size_t)p.depth * 4 * TILE;           // [NCW][GCH][GROW]
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n = p.n;
  const int nsteps = n - 1;
  const long long s0 = (long long)tile * TILE + NS * tid;
  const double2 cnt2 = *reinterpret_cast<const double2*>(p.counts + (size_t)t * p.S_pad + s0);
  const double cnt[NS] = {cnt2.x, cnt2.y};

  // ------------------------------------------------------------------ up pass (prune.py:15-49)
  double sc[NS] = {1.0, 1.0};
  int ex[NS] = {0, 0};
  double ll_site[NS] = {0.0, 0.0};
  double carry[NS][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};  // result of the previous step
  for (int c0 = 0; c0 < nsteps; c0 += CH) {
    const unsigned char* stage = smem + OFF_STAGE + rs.cst * STAGE_BYTES;
    mbar_wait(bar_full_c(bars, rs.cst), rs.cph);
    const int cn = min(CH, nsteps - c0);
    for (int k = 0; k < cn; ++k) {
      const double* rec = reinterpret_cast<const double*>(stage) + k * UPREC;
      const int4 op = *reinterpret_cast<const int4*>(rec);
      const unsigned char* crow = stage + STAGE_CODES + 2 * k * CODE_ROW + NS * tid;
      double (&ma)[NS][4] = carry;  // child a: already there when it is the previous step's result
      double mb[NS][4], dummy[NS][4];
      if (op.w & 1)
        tip_messages<CLIP, false>(rec + UP_A, nullptr, crow, ma, dummy);
      else if (!(op.w & 8))
        load_rows(stk + (op.z & 0xff) * (4 * TILE) + NS * tid, ma);
      if (op.w & 2)
        tip_messages<CLIP, false>(rec + UP_B, nullptr, crow + CODE_ROW, mb, dummy);
      else
        load_rows(stk + ((op.z >> 8) & 0xff) * (4 * TILE) + NS * tid, mb);
      double pr[NS][4];
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) pr[s][i] = ma[s][i] * mb[s][i];
      if (op.w & 4) {
        // log(post[root] . pi) + const[root]  (prune.py:131) with post = Pprod / c folded in
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const double like = (pr[s][0] * p.pi[0] + pr[s][1] * p.pi[1]) + (pr[s][2] * p.pi[2] + pr[s][3] * p.pi[3]);
          ll_site[s] = log(CLIP ? like * sc[s] : like) + (double)ex[s] * 0.6931471805599453094;
        }
      } else {
        double post[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const double c = (pr[s][0] + pr[s][1]) + (pr[s][2] + pr[s][3]);
          if (CLIP) {
            const double inv = 1.0 / c;  // Pprod /= c (prune.py:34)
            sc[s] *= c;
            const int e = exponent_of(sc[s]);
            ex[s] += e;
            sc[s] = scale_pow2(sc[s], e);
#pragma unroll
            for (int i = 0; i < 4; ++i) post[s][i] = pr[s][i] * inv;
          } else {
            // messages stay unnormalised; only when the magnitude has drifted below 2^-256 is it pulled back
            // by an exact power of two (a step shrinks it by at most ~2^-133, so nothing underflows)
#pragma unroll
            for (int i = 0; i < 4; ++i) post[s][i] = pr[s][i];
            if (__double2hiint(c) < ((1023 - 256) << 20)) {
              const int e = exponent_of(c);
              ex[s] += e;
#pragma unroll
              for (int i = 0; i < 4; ++i) post[s][i] = scale_pow2(pr[s][i], e);
            }
          }
        }
        double M[16];
        lds16(rec + UP_P, M);
        double v[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            v[s][i] = clipv<CLIP>(M[i * 4 + 0] * post[s][0] + M[i * 4 + 1] * post[s][1] + M[i * 4 + 2] * post[s][2] + M[i * 4 + 3] * post[s][3]);
        copy_rows(v, carry);  // the next step's child a, if it consumes this result
        if (op.w & 32) store_rows(stk + ((op.z >> 16) & 0xff) * (4 * TILE) + NS * tid, v);  // parked for a later step
        if (WANT_GRAD) {
          double* gdst = myscr + (size_t)(c0 + k) * (4 * TILE);
#pragma unroll
          for (int i = 0; i < 4; ++i) __stcg(reinterpret_cast<double2*>(gdst + i * TILE), make_double2(v[0][i], v[1][i]));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty_c(bars, rs.cst));
    rs.adv_c();
  }
  if (WANT_GRAD) {
    // make this thread's scratch stores visible to the producer's async-proxy reads, then signal
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_up_done(bars));
  }
  if (p.site_ll != nullptr) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s0 + s < p.S) p.site_ll[(size_t)u * p.S + s0 + s] = ll_site[s];
  }
  {
    // sum_s counts[s] * ll_s over the tile (bdsky.py:341)
    double v = cnt[0] * ll_site[0] + cnt[1] * ll_site[1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) p.ll_part[((size_t)u * p.tiles + tile) * NCW + (tid >> 5)] = v;  // per-warp partial
  }
  if (!WANT_GRAD) return;

  // ------------------------------------------- down pass (prune.py:52-96) + eq. (9) (prune.py:146)
  double* gout = p.g_part + (((size_t)u * p.tiles + tile) * NCW + (tid >> 5)) * (2 * n - 2);  // this warp's partials
  double* gstw = gst + (tid >> 5) * (GCH * GROW);
  for (int c0 = 0; c0 < nsteps; c0 += CH) {
    const unsigned char* stage = smem + OFF_STAGE + rs.cst * STAGE_BYTES;
    mbar_wait(bar_full_c(bars, rs.cst), rs.cph);
    const int cn = min(CH, nsteps - c0);
    for (int k = 0; k < cn; ++k) {
      const int step = c0 + k;
      const double* rec = reinterpret_cast<const double*>(stage) + k * DNREC;
      const int4 op = *reinterpret_cast<const int4*>(rec);
      const unsigned char* crow = stage + STAGE_CODES + 2 * k * CODE_ROW + NS * tid;
      const int sp = op.z & 0xff;
      double (&pre)[NS][4] = carry;  // already there when the previous step computed it
      if (op.w & 8) {
      } else if (sp == 0xff) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i) pre[s][i] = p.pi[i];  // pre[root] = pi (prune.py:88)
      } else {
        load_rows(stk + sp * (4 * TILE) + NS * tid, pre);
      }
      double ma[NS][4], mb[NS][4], qa[NS][4], qb[NS][4];
      if (op.w & 1) {
        tip_messages<CLIP, true>(rec + DN_A, rec + DN_QA, crow, ma, qa);
      } else {
        mbar_wait(bar_full_m(bars, rs.mst), rs.mph);
        load_rows(reinterpret_cast<const double*>(smem + OFF_MSG + rs.mst * MSG_BYTES) + NS * tid, ma);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty_m(bars, rs.mst));
        rs.adv_m(p.nms);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            qa[s][i] = p.Q[i * 4 + 0] * ma[s][0] + p.Q[i * 4 + 1] * ma[s][1] + p.Q[i * 4 + 2] * ma[s][2] + p.Q[i * 4 + 3] * ma[s][3];
      }
      if (op.w & 2) {
        tip_messages<CLIP, true>(rec + DN_B, rec + DN_QB, crow + CODE_ROW, mb, qb);
      } else {
        mbar_wait(bar_full_m(bars, rs.mst), rs.mph);
        load_rows(reinterpret_cast<const double*>(smem + OFF_MSG + rs.mst * MSG_BYTES) + NS * tid, mb);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty_m(bars, rs.mst));
        rs.adv_m(p.nms);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            qb[s][i] = p.Q[i * 4 + 0] * mb[s][0] + p.Q[i * 4 + 1] * mb[s][1] + p.Q[i * 4 + 2] * mb[s][2] + p.Q[i * 4 + 3] * mb[s][3];
      }
      double wa[NS][4], wb[NS][4];
      int ed[NS];
      double gsum_a = 0.0, gsum_b = 0.0;
      const int gslot = (2 * step) & (GCH - 1);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wa[s][i] = pre[s][i] * mb[s][i], wb[s][i] = pre[s][i] * ma[s][i];
        const double den = (wa[s][0] * ma[s][0] + wa[s][1] * ma[s][1]) + (wa[s][2] * ma[s][2] + wa[s][3] * ma[s][3]);
        ed[s] = exponent_of(den);
        const double rden = cnt[s] * rcp_pos(den);
        const double ga = ((wa[s][0] * qa[s][0] + wa[s][1] * qa[s][1]) + (wa[s][2] * qa[s][2] + wa[s][3] * qa[s][3])) * rden;
        const double gb = ((wb[s][0] * qb[s][0] + wb[s][1] * qb[s][1]) + (wb[s][2] * qb[s][2] + wb[s][3] * qb[s][3])) * rden;
        gsum_a += ga, gsum_b += gb;
      }
      gstw[gslot * GROW + lane] = gsum_a;
      gstw[(gslot + 1) * GROW + lane] = gsum_b;

      if (!(op.w & 1)) {
        double M[16];
        lds16(rec + DN_A, M);
        double B[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            B[s][v] = clipv<CLIP>(wa[s][0] * M[0 * 4 + v] + wa[s][1] * M[1 * 4 + v] + wa[s][2] * M[2 * 4 + v] + wa[s][3] * M[3 * 4 + v]);
          if (CLIP) {
            const double inv = 1.0 / ((B[s][0] + B[s][1]) + (B[s][2] + B[s][3]));  // B /= c (prune.py:74)
#pragma unroll
            for (int v = 0; v < 4; ++v) B[s][v] *= inv;
          } else if (ed[s] < -400) {
            // any positive rescaling cancels in g = num/den: pull pre back by 2^-exponent(den) only when it has drifted far
#pragma unroll
            for (int v = 0; v < 4; ++v) B[s][v] = scale_pow2(B[s][v], ed[s]);
          }
        }
        copy_rows(B, carry);  // an internal child a is always expanded by the next step
      }
      if (!(op.w & 2)) {
        double M[16];
        lds16(rec + DN_B, M);
        double B[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            B[s][v] = clipv<CLIP>(wb[s][0] * M[0 * 4 + v] + wb[s][1] * M[1 * 4 + v] + wb[s][2] * M[2 * 4 + v] + wb[s][3] * M[3 * 4 + v]);
          if (CLIP) {
            const double inv = 1.0 / ((B[s][0] + B[s][1]) + (B[s][2] + B[s][3]));
#pragma unroll
            for (int v = 0; v < 4; ++v) B[s][v] *= inv;
          } else if (ed[s] < -400) {
#pragma unroll
            for (int v = 0; v < 4; ++v) B[s][v] = scale_pow2(B[s][v], ed[s]);
          }
        }
        store_rows(stk + ((op.z >> 16) & 0xff) * (4 * TILE) + NS * tid, B);  // an internal child b is parked until the walk returns to it
      }

      if (gslot + 2 == GCH || step == nsteps - 1) {
        // transposed reduction of the staged round over the warp's 64 sites: lane (r, q) adds 8 lanes' values
        // of branch r, two shuffles fold the four q groups; no CTA-wide barrier anywhere in the walk
        const int nval = gslot + 2;
        __syncwarp();
        const int r = lane & (GCH - 1), q = lane / GCH;
        const double* row = gstw + r * GROW + q * GCH;
        double acc = 0.0;
#pragma unroll
        for (int kk = 0; kk < GCH; ++kk) acc += row[kk];
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < nval) gout[2 * step + 2 - nval + lane] = acc;
        __syncwarp();
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty_c(bars, rs.cst));
    rs.adv_c();
  }
}

// Each launch handles the units whose clip flag equals CLIP and skips the others (two launches per
// evaluation; the literal-clip one normally finds nothing to do).  Keeping one arithmetic variant per
// kernel halves the code the instruction cache has to hold.
template <bool WANT_GRAD, bool CLIP>
__global__ void __launch_bounds__(THREADS, 3) prune_kernel(const PruneParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t bars = smem_u32(smem);
  const int tid = threadIdx.x;
  const int n = p.n;
  const int nsteps = n - 1;
  const long long items = (long long)p.U * p.tiles;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(bar_full_c(bars, i), 1), mbar_init(bar_empty_c(bars, i), NCW);
    for (int i = 0; i < NMS_MAX; ++i) mbar_init(bar_full_m(bars, i), 1), mbar_init(bar_empty_m(bars, i), NCW);
    mbar_init(bar_up_done(bars), NCW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  RingState rs;
  const unsigned char* scr_base = reinterpret_cast<const unsigned char*>(p.scratch) + (size_t)blockIdx.x * nsteps * MSG_BYTES;

  if (tid >= CT) {
    // =============================================================== producer warp (one lane)
    if (tid != CT) return;
    uint32_t up_phase = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const int u = (int)(item / p.tiles);
      if (((p.flags[u] & 1u) != 0) != CLIP) continue;
      const int tile = (int)(item - (long long)u * p.tiles);
      const int t = p.unit_tree[u];
      const unsigned char* utab = reinterpret_cast<const unsigned char*>(p.uptab) + (size_t)u * nsteps * (UPREC * 8);
      const unsigned char* cup = p.codes_up + ((size_t)t * p.tiles + tile) * (size_t)(2 * nsteps) * CODE_ROW;
      for (int c0 = 0; c0 < nsteps; c0 += CH) {
        const int cn = min(CH, nsteps - c0);
        mbar_wait_sleep(bar_empty_c(bars, rs.cst), rs.cph ^ 1);
        const uint32_t dst = smem_u32(smem + OFF_STAGE + rs.cst * STAGE_BYTES);
        const uint32_t full = bar_full_c(bars, rs.cst);
        mbar_expect_tx(full, cn * (UPREC * 8 + 2 * CODE_ROW));
        bulk_g2s(dst, utab + (size_t)c0 * (UPREC * 8), cn * UPREC * 8, full);
        bulk_g2s(dst + STAGE_CODES, cup + (size_t)c0 * 2 * CODE_ROW, cn * 2 * CODE_ROW, full);
        rs.adv_c();
      }
      if (!WANT_GRAD) continue;
      const unsigned char* dtab = reinterpret_cast<const unsigned char*>(p.dntab) + (size_t)u * nsteps * (DNREC * 8);
      const unsigned char* cdn = p.codes_dn + ((size_t)t * p.tiles + tile) * (size_t)(2 * nsteps) * CODE_ROW;
      const int nchunks = (nsteps + CH - 1) / CH;
      const int32_t* mrows = p.msg_rows + (size_t)t * max(n - 2, 1);
      const int32_t* mstart = p.msg_start + (size_t)t * (nchunks + 1);
      mbar_wait_sleep(bar_up_done(bars), up_phase);  // every consumer warp has finished (and fenced) its scratch stores
      up_phase ^= 1;
      int mi = 0;
      for (int c0 = 0, c = 0; c0 < nsteps; c0 += CH, ++c) {
        const int cn = min(CH, nsteps - c0);
        mbar_wait_sleep(bar_empty_c(bars, rs.cst), rs.cph ^ 1);
        const uint32_t dst = smem_u32(smem + OFF_STAGE + rs.cst * STAGE_BYTES);
        const uint32_t full = bar_full_c(bars, rs.cst);
        mbar_expect_tx(full, cn * (DNREC * 8 + 2 * CODE_ROW));
        bulk_g2s(dst, dtab + (size_t)c0 * (DNREC * 8), cn * DNREC * 8, full);
        bulk_g2s(dst + STAGE_CODES, cdn + (size_t)c0 * 2 * CODE_ROW, cn * 2 * CODE_ROW, full);
        rs.adv_c();
        const int mend = mstart[c + 1];
        for (; mi < mend; ++mi) {
          const int row = mrows[mi];
          mbar_wait_sleep(bar_empty_m(bars, rs.mst), rs.mph ^ 1);
          const uint32_t mfull = bar_full_m(bars, rs.mst);
          mbar_expect_tx(mfull, MSG_BYTES);
          bulk_g2s(smem_u32(smem + OFF_MSG + rs.mst * MSG_BYTES), scr_base + (size_t)row * MSG_BYTES, MSG_BYTES, mfull);
          rs.adv_m(p.nms);
        }
      }
    }
    return;
  }

  // ================================================================= consumer warps
  double* myscr = p.scratch + (size_t)blockIdx.x * nsteps * (4 * TILE) + NS * tid;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int u = (int)(item / p.tiles);
    const int tile = (int)(item - (long long)u * p.tiles);
    const int t = p.unit_tree[u];
    if (((p.flags[u] & 1u) != 0) != CLIP) continue;
    consume_item<WANT_GRAD, CLIP>(p, smem, rs, u, tile, t, myscr);
  }
}

// ------------------------------------------------------------------------------ transition tables
struct TablesParams {
  const double* bl;  // [U][2n-2]
  const int32_t* unit_tree;
  const Op* up;
  const Op* down;  // [T][n-1]
  double* uptab;
  double* dntab;
  uint32_t* flags;
  vbsky_subst_model mdl;
  int U, n, want_grad;
};

// P(t) = U diag(exp(d t)) Uinv (substitution.py:61-62); optionally Q P(t) = U diag(d exp(d t)) Uinv.
__device__ __forceinline__ bool expm16(const vbsky_subst_model& m, double t, bool tip, double* P, double* QP) {
  // Entries are accumulated as delta_ij + sum_k U_ik expm1(d_k t) Uinv_kj: for the short branches of
  // real data (t ~ 1e-6..1e-3) the off-diagonals of sum_k U_ik exp(d_k t) Uinv_kj are the remainder of
  // O(1) terms cancelling, i.e. carry ~1e-16/t relative rounding noise (as the reference's own
  // expm_action does); the expm1 form is exact to ~1 ulp.
  double e1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) e1[k] = expm1(t * m.d[k]);
  bool small = false;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = (i == j) ? 1.0 : 0.0, q = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double uu = m.P[i * 4 + k] * m.Pinv[k * 4 + j];
        s += uu * e1[k];
        q += (m.d[k] * uu) * (e1[k] + 1.0);
      }
      if (tip) {
        P[j * 4 + i] = s;  // column-major
        if (QP) QP[j * 4 + i] = q;
      } else {
        P[i * 4 + j] = s;
      }
      small |= !(s >= 1e-20) || !(s <= 2.0);  // also catches NaN / inf
    }
  if (tip) {
    // fifth column: the fully ambiguous state set (sum of the four columns, in the reference's order of addition)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      P[16 + i] = ((P[i] + P[4 + i]) + P[8 + i]) + P[12 + i];
      if (QP) QP[16 + i] = ((QP[i] + QP[4 + i]) + QP[8 + i]) + QP[12 + i];
    }
  }
  return small;
}

__global__ void tables_kernel(const TablesParams p) {
  const int nsteps = p.n - 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.U * nsteps) return;
  const int u = (int)(idx / nsteps), step = (int)(idx - (long long)u * nsteps);
  const int t = p.unit_tree[u];
  const double* bl = p.bl + (size_t)u * (2 * p.n - 2);
  bool small = false;
  {
    const Op op = p.up[(size_t)t * nsteps + step];
    double* rec = p.uptab + idx * UPREC;
    *reinterpret_cast<int4*>(rec) = make_int4(op.x, op.y, op.z, op.w);
    const int a = op.x & 0xffff, b = (op.x >> 16) & 0xffff;
    if (op.w & 1) small |= expm16(p.mdl, bl[a], true, rec + UP_A, nullptr);
    if (op.w & 2) small |= expm16(p.mdl, bl[b], true, rec + UP_B, nullptr);
    if (!(op.w & 4)) small |= expm16(p.mdl, bl[op.y], false, rec + UP_P, nullptr);
  }
  if (p.want_grad) {
    const Op op = p.down[(size_t)t * nsteps + step];
    double* rec = p.dntab + idx * DNREC;
    *reinterpret_cast<int4*>(rec) = make_int4(op.x, op.y, op.z, op.w);
    const int a = op.x & 0xffff, b = (op.x >> 16) & 0xffff;
    small |= expm16(p.mdl, bl[a], (op.w & 1) != 0, rec + DN_A, (op.w & 1) ? rec + DN_QA : nullptr);
    small |= expm16(p.mdl, bl[b], (op.w & 2) != 0, rec + DN_B, (op.w & 2) ? rec + DN_QB : nullptr);
  }
  if (small) atomicOr(p.flags + u, 1u);
}

// Fixed-order sums of the per-tile partials: ll[u] and dll_dbl[u][node].
__global__ void finalize_kernel(const double* __restrict__ ll_part, const double* __restrict__ g_part,
                                const int32_t* __restrict__ unit_tree, const int32_t* __restrict__ down_nodes,
                                int U, int n, int tiles, double* __restrict__ ll, double* __restrict__ dll) {
  const int nb = 2 * n - 2;
  const int u = blockIdx.x;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (j < nb && dll != nullptr) {
    const double* src = g_part + (size_t)u * tiles * nb + j;
    double acc = 0.0;
    for (int k = 0; k < tiles; ++k) acc += src[(size_t)k * nb];
    const int node = down_nodes[(size_t)unit_tree[u] * nb + j];
    dll[(size_t)u * nb + node] = acc;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    double acc = 0.0;
    for (int k = 0; k < tiles; ++k) acc += ll_part[(size_t)u * tiles + k];
    ll[u] = acc;
  }
}

struct PruneConfig {
  int grid = 0, depth = 0, tiles = 0, nms = 2;
  size_t smem = 0;
  size_t off_uptab = 0, off_dntab = 0, off_flags = 0, off_scratch = 0, off_llpart = 0, off_gpart = 0, total = 0;
};

static thread_local int64_t g_stats[8] = {0};

// Optional CUDA-event timing of prune_kernel alone (bench.py's roofline leg): when enabled, every call
// brackets the kernel with an event pair on the launch stream; vbsky_prune_timing() drains them.
struct TimingState {
  bool enabled = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};
static thread_local TimingState g_timing;

static int prune_config(const vbsky_layout* lay, const vbsky_tips* tips, int U, bool want_grad, PruneConfig* cfg) {
  const int n = lay->n;
  cfg->depth = want_grad ? std::max(lay->depth_up, lay->depth_down) : lay->depth_up;
  cfg->tiles = tips->tiles;
  int dev = 0, sms = 0, per_sm = 0, max_smem = 0;
  VB_CUDA(cudaGetDevice(&dev));
  VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  VB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  auto smem_for = [&](int nms) {
    return (size_t)OFF_MSG + (size_t)nms * MSG_BYTES + ((size_t)cfg->depth * 4 * TILE + (size_t)NCW * GCH * GROW) * sizeof(double);
  };
  if (smem_for(2) > (size_t)max_smem)
    return set_error(VBSKY_EINVAL, "tree needs %d parked vectors (%zu B shared memory) but the device offers %d B", cfg->depth, smem_for(2), max_smem);
  auto occupancy = [&](size_t smem, int* out) -> int {
    int a = 0, b = 0;
    if (want_grad) {
      VB_CUDA(cudaFuncSetAttribute(prune_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VB_CUDA(cudaFuncSetAttribute(prune_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, prune_kernel<true, false>, THREADS, smem));
      VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, prune_kernel<true, true>, THREADS, smem));
    } else {
      VB_CUDA(cudaFuncSetAttribute(prune_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VB_CUDA(cudaFuncSetAttribute(prune_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, prune_kernel<false, false>, THREADS, smem));
      VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, prune_kernel<false, true>, THREADS, smem));
    }
    *out = std::min(a, b);
    return VBSKY_OK;
  };
  // deepest message ring that does not cost a resident CTA
  int rc = occupancy(smem_for(2), &per_sm);
  if (rc != VBSKY_OK) return rc;
  cfg->nms = 2;
  for (int nms = 3; nms <= NMS_MAX; ++nms) {
    int occ = 0;
    if (smem_for(nms) > (size_t)max_smem) break;
    rc = occupancy(smem_for(nms), &occ);
    if (rc != VBSKY_OK) return rc;
    if (occ < per_sm) break;
    cfg->nms = nms;
  }
  cfg->smem = smem_for(cfg->nms);
  rc = occupancy(cfg->smem, &per_sm);  // leaves the function attributes at the chosen size
  if (rc != VBSKY_OK) return rc;
  if (per_sm < 1) return set_error(VBSKY_ECUDA, "pruning kernel cannot be resident (smem %zu B)", cfg->smem);
  const long long items = (long long)U * cfg->tiles;
  const long long slots = (long long)sms * per_sm;
  cfg->grid = (int)std::min<long long>(slots, std::max<long long>(items, 1));
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off = 0;
  cfg->off_uptab = off, off = align(off + (size_t)U * (n - 1) * UPREC * sizeof(double));
  cfg->off_dntab = off, off = align(off + (want_grad ? (size_t)U * (n - 1) * DNREC * sizeof(double) : 0));
  cfg->off_flags = off, off = align(off + (size_t)U * sizeof(uint32_t));
  cfg->off_scratch = off, off = align(off + (want_grad ? (size_t)slots * (n - 1) * MSG_BYTES : 0));
  cfg->off_llpart = off, off = align(off + (size_t)U * cfg->tiles * NCW * sizeof(double));
  cfg->off_gpart = off, off = align(off + (want_grad ? (size_t)U * cfg->tiles * NCW * (2 * n - 2) * sizeof(double) : 0));
  cfg->total = off;
  return VBSKY_OK;
}

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_prune_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t want_grad) {
  if (!lay || !tips || U < 1) {
    set_error(VBSKY_EINVAL, "bad argument to vbsky_prune_workspace_bytes");
    return 0;
  }
  PruneConfig cfg;
  if (prune_config(lay, tips, U, want_grad != 0, &cfg) != VBSKY_OK) return 0;
  return cfg.total;
}

extern "C" int vbsky_prune_value_and_grad(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U,
                                          const int32_t* unit_tree, const double* bl,
                                          const vbsky_subst_model* model, double* site_ll, double* ll,
                                          double* dll_dbl, void* ws, size_t ws_bytes, void* stream_) {
  VB_CHECK_ARG(lay && tips && unit_tree && bl && model && ll && ws, "NULL argument");
  VB_CHECK_ARG(U >= 1, "U must be >= 1");
  VB_CHECK_ARG(lay->device >= 0, "layout was created without a CUDA device; there is no CPU fallback");
  VB_CHECK_ARG(lay->T == tips->T && lay->n == tips->n, "layout (T=%d,n=%d) and tips (T=%d,n=%d) do not match", lay->T, lay->n, tips->T, tips->n);
  // prune.py:120-126 asserts a square rate matrix of the tips' alphabet size; A = 4 is fixed here.
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool want_grad = dll_dbl != nullptr;
  PruneConfig cfg;
  int rc = prune_config(lay, tips, U, want_grad, &cfg);
  if (rc != VBSKY_OK) return rc;
  if (ws_bytes < cfg.total) return set_error(VBSKY_ENOMEM, "workspace too small: %zu < %zu bytes", ws_bytes, cfg.total);
  const int n = lay->n;
  char* base = (char*)ws;

  TablesParams tp;
  tp.bl = bl, tp.unit_tree = unit_tree, tp.up = lay->d_up, tp.down = lay->d_down;
  tp.uptab = (double*)(base + cfg.off_uptab), tp.dntab = (double*)(base + cfg.off_dntab);
  tp.flags = (uint32_t*)(base + cfg.off_flags);
  tp.mdl = *model, tp.U = U, tp.n = n, tp.want_grad = want_grad;
  VB_CUDA(cudaMemsetAsync(tp.flags, 0, (size_t)U * sizeof(uint32_t), stream));
  {
    const long long tot = (long long)U * (n - 1);
    tables_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, stream>>>(tp);
    VB_CUDA(cudaGetLastError());
  }

  PruneParams p;
  p.uptab = tp.uptab, p.dntab = tp.dntab, p.flags = tp.flags;
  p.codes_up = tips->d_codes_up, p.codes_dn = tips->d_codes_dn, p.counts = tips->d_counts;
  p.unit_tree = unit_tree;
  p.msg_rows = lay->d_msg_rows, p.msg_start = lay->d_msg_start;
  p.scratch = (double*)(base + cfg.off_scratch);
  p.site_ll = site_ll;
  p.ll_part = (double*)(base + cfg.off_llpart);
  p.g_part = (double*)(base + cfg.off_gpart);
  vbsky_subst_rate_matrix(model, p.Q);
  for (int i = 0; i < 4; ++i) p.pi[i] = model->pi[i];
  p.n = n, p.U = U, p.tiles = cfg.tiles, p.depth = cfg.depth, p.nms = cfg.nms;
  p.S = tips->S, p.S_pad = tips->S_pad;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (g_timing.enabled) {
    VB_CUDA(cudaEventCreate(&ev0));
    VB_CUDA(cudaEventCreate(&ev1));
    VB_CUDA(cudaEventRecord(ev0, stream));
  }
  if (want_grad)
    prune_kernel<true, false><<<cfg.grid, THREADS, cfg.smem, stream>>>(p);
  else
    prune_kernel<false, false><<<cfg.grid, THREADS, cfg.smem, stream>>>(p);
  VB_CUDA(cudaGetLastError());
  if (g_timing.enabled) {
    VB_CUDA(cudaEventRecord(ev1, stream));
    g_timing.pending.push_back({ev0, ev1});
  }
  // units with a (near-)zero branch length: literal clip arithmetic
  if (want_grad)
    prune_kernel<true, true><<<cfg.grid, THREADS, cfg.smem, stream>>>(p);
  else
    prune_kernel<false, true><<<cfg.grid, THREADS, cfg.smem, stream>>>(p);
  VB_CUDA(cudaGetLastError());
  {
    dim3 grid((unsigned)U, (unsigned)((2 * n - 2 + 127) / 128));
    finalize_kernel<<<grid, 128, 0, stream>>>(p.ll_part, p.g_part, unit_tree, lay->d_down_nodes, U, n, cfg.tiles * NCW, ll, dll_dbl);
    VB_CUDA(cudaGetLastError());
  }
  g_stats[0] = 4, g_stats[1] = cfg.grid, g_stats[2] = TILE, g_stats[3] = (int64_t)cfg.smem;
  {
    // Algorithmic HBM bytes of prune_kernel: every non-root internal message written once and read
    // once (32 B per site), tip codes 1 B per tip-site and pass, counts, per-site ll out, per-tile gradient
    // partials out.  The per-unit step tables are L2-resident across a unit's tiles and not counted.
    const double S = (double)tips->S;
    double bytes = (double)U * S * n * (want_grad ? 2.0 : 1.0) + (double)U * S * 8.0 + (site_ll ? (double)U * S * 8.0 : 0.0);
    if (want_grad) bytes += (double)U * S * (n - 2) * 32.0 * 2.0 + (double)U * cfg.tiles * NCW * (2.0 * n - 2) * 8.0;
    g_stats[4] = (int64_t)bytes;
  }
  return VBSKY_OK;
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

