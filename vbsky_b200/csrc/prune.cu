// Felsenstein pruning log-likelihood and branch-length gradient for sm_100a.
//
// Replaces vmap(prune.prune_loglik) x counts (bdsky.py:331-343) and prune_loglik_jvp
// (prune.py:134-150) of the reference for every (tree, sample) unit at once.
//
// Design (DESIGN.md "Pruning kernel"):
//   * one persistent CTA per resident slot; a work item is (unit, tile of TILE site patterns);
//     one thread owns one site pattern for the whole tree walk;
//   * P(t) = U exp(Lambda t) U^-1 is built once per (unit, branch) by transition_kernel, not per
//     site as the reference does inside its vmap (substitution.py:64-70);
//   * the walk follows the layout's depth-first schedules, so the live "message" vectors
//     (clip(P(t_c) post_c), the quantity both passes actually consume) sit in a small
//     shared-memory stack; the up pass streams each internal node's message to a CTA-private
//     HBM scratch row once and the down pass streams it back once.  pre vectors never leave
//     the SM.  Value-only evaluation touches no scratch at all;
//   * per-site scale factors are accumulated as mantissa x 2^exponent (one log per site instead
//     of one per node, prune.py:37), and the gradient uses the scale-free form of eq. (9),
//       g_c = (w^T Q m_c) / (w^T m_c),  w = pre_parent (.) m_sibling,
//     which equals prune.py:146-148 whenever the 1e-40 clips are inactive;
//   * branch gradients are reduced over the tile's sites through a transposed shared-memory
//     staging buffer (1 add per value instead of a 5-step shuffle tree), written as per-tile
//     partials and summed in a fixed order by finalize_kernel (bitwise reproducible).
#include <cmath>

#include "common.h"

namespace vbsky {

constexpr int TILE = 128;  // site patterns per CTA tile == threads per CTA
constexpr int GCH = 16;    // branch gradients staged per reduction round
constexpr int GROW = TILE + 1;

struct PruneParams {
  const Op* up;
  const Op* down;
  const uint8_t* codes;
  const double* counts;
  const int32_t* unit_tree;
  const double* ptab;   // [U][2n-2][16]
  const double* qptab;  // [U][n][16]
  double* scratch;      // [grid][n-1][4][TILE]
  double* site_ll;      // [U][S] or null
  double* ll_part;      // [U][tiles]
  double* g_part;       // [U][tiles][2n-2]
  double Q[16];
  double pi[4];
  int n, U, tiles, depth;
  long long S, S_pad;
};

__device__ __constant__ uint8_t kRankToMask[16] = {1, 2, 4, 8, 15, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 0};

__device__ __forceinline__ double clip01(double x) {
  // jnp.clip(ret, 1e-40, 1.0) of substitution.py:70
  return fmin(fmax(x, 1e-40), 1.0);
}

__device__ __forceinline__ void load16(const double* __restrict__ p, double (&M)[16]) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double2 v = __ldg(q + i);
    M[2 * i] = v.x, M[2 * i + 1] = v.y;
  }
}

// Message of a tip: clip(P(t) e_mask) where e_mask is the 0/1 tip partial (substitution.py:10-35).
// rank < 4 is a single column of P.
__device__ __forceinline__ void tip_columns(const double* __restrict__ M, int rank, bool all_simple, double (&out)[4]) {
  if (all_simple) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __ldg(M + i * 4 + rank);
  } else {
    const int mask = kRankToMask[rank];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double v = __ldg(M + i * 4 + j);
        s += ((mask >> j) & 1) ? v : 0.0;
      }
      out[i] = s;
    }
  }
}

template <bool WANT_GRAD>
__global__ void __launch_bounds__(TILE, 4) prune_kernel(const PruneParams p) {
  extern __shared__ double smem[];
  double* stk = smem;                              // [depth][4][TILE]
  double* gst = stk + (size_t)p.depth * 4 * TILE;  // [GCH][GROW]
  double* red = gst + GCH * GROW;                  // [TILE/GCH][GCH] and block reductions
  const int tid = threadIdx.x;
  const int n = p.n, root = 2 * n - 2;
  const long long items = (long long)p.U * p.tiles;
  double* myscr = p.scratch + (size_t)blockIdx.x * (n - 1) * 4 * TILE + tid;

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int u = (int)(item / p.tiles);
    const int tile = (int)(item - (long long)u * p.tiles);
    const int t = p.unit_tree[u];
    const long long s = (long long)tile * TILE + tid;
    const Op* up = p.up + (size_t)t * (n - 1);
    const uint8_t* codes = p.codes + (size_t)t * n * p.S_pad + s;
    const double* ptab = p.ptab + (size_t)u * (2 * n - 2) * 16;
    const double cnt = p.counts[(size_t)t * p.S_pad + s];

    // ------------------------------------------------------------------ up pass (prune.py:15-49)
    double sc = 1.0;
    int ex = 0;
    double ll_site = 0.0;
    for (int step = 0; step < n - 1; ++step) {
      const Op op = up[step];
      const int a = op.x, b = op.y, par = op.z;
      double ma[4], mb[4];
      if (a < n) {
        const int r = codes[(size_t)a * p.S_pad];
        tip_columns(ptab + a * 16, r, __all_sync(0xffffffffu, r < 4), ma);
#pragma unroll
        for (int i = 0; i < 4; ++i) ma[i] = clip01(ma[i]);
      } else {
        const double* src = stk + (size_t)(op.w & 0xff) * 4 * TILE + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) ma[i] = src[i * TILE];
      }
      if (b < n) {
        const int r = codes[(size_t)b * p.S_pad];
        tip_columns(ptab + b * 16, r, __all_sync(0xffffffffu, r < 4), mb);
#pragma unroll
        for (int i = 0; i < 4; ++i) mb[i] = clip01(mb[i]);
      } else {
        const double* src = stk + (size_t)((op.w >> 8) & 0xff) * 4 * TILE + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) mb[i] = src[i * TILE];
      }
      double pr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pr[i] = ma[i] * mb[i];
      if (par == root) {
        // log(post[root] . pi) + const[root]  (prune.py:131) with post = Pprod / c folded in
        const double like = (pr[0] * p.pi[0] + pr[1] * p.pi[1]) + (pr[2] * p.pi[2] + pr[3] * p.pi[3]);
        ll_site = log(like * sc) + (double)ex * 0.6931471805599453094;
      } else {
        const double c = (pr[0] + pr[1]) + (pr[2] + pr[3]);
        const double inv = 1.0 / c;
        sc *= c;
        {
          const int hi = __double2hiint(sc);
          const int e = (hi >> 20) - 1023;
          ex += e;
          sc = __hiloint2double(hi - (e << 20), __double2loint(sc));
        }
        double post[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) post[i] = pr[i] * inv;
        double M[16];
        load16(ptab + par * 16, M);
        double* dst = stk + (size_t)((op.w >> 16) & 0xff) * 4 * TILE + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double v = clip01(M[i * 4 + 0] * post[0] + M[i * 4 + 1] * post[1] + M[i * 4 + 2] * post[2] + M[i * 4 + 3] * post[3]);
          dst[i * TILE] = v;
          if (WANT_GRAD) __stcg(myscr + ((size_t)step * 4 + i) * TILE, v);
        }
      }
    }
    if (p.site_ll != nullptr && s < p.S) p.site_ll[(size_t)u * p.S + s] = ll_site;
    {
      // sum_s counts[s] * ll_s over the tile (bdsky.py:341)
      double v = cnt * ll_site;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __syncthreads();
      if ((tid & 31) == 0) red[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < TILE / 32; ++w) tot += red[w];
        p.ll_part[(size_t)u * p.tiles + tile] = tot;
      }
    }
    if (!WANT_GRAD) continue;

    // ------------------------------------------- down pass (prune.py:52-96) + eq. (9) (prune.py:146)
    const Op* down = p.down + (size_t)t * (n - 1);
    const double* qptab = p.qptab + (size_t)u * n * 16;
    double* gout = p.g_part + ((size_t)u * p.tiles + tile) * (2 * n - 2);
    for (int step = 0; step < n - 1; ++step) {
      const Op op = down[step];
      const int a = op.x & 0xffff, b = (op.x >> 16) & 0xffff;
      const int sp = op.z & 0xff;
      double pre[4];
      if (sp == 0xff) {
#pragma unroll
        for (int i = 0; i < 4; ++i) pre[i] = p.pi[i];  // pre[root] = pi (prune.py:88)
      } else {
        const double* src = stk + (size_t)sp * 4 * TILE + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) pre[i] = src[i * TILE];
      }
      double ma[4], mb[4], qa[4], qb[4];
      if (a < n) {
        const int r = codes[(size_t)a * p.S_pad];
        const bool simple = __all_sync(0xffffffffu, r < 4);
        tip_columns(ptab + a * 16, r, simple, ma);
        tip_columns(qptab + a * 16, r, simple, qa);
#pragma unroll
        for (int i = 0; i < 4; ++i) ma[i] = clip01(ma[i]);
      } else {
        const double* src = myscr + (size_t)(op.y & 0xffff) * 4 * TILE;
#pragma unroll
        for (int i = 0; i < 4; ++i) ma[i] = __ldcg(src + i * TILE);
#pragma unroll
        for (int i = 0; i < 4; ++i) qa[i] = p.Q[i * 4 + 0] * ma[0] + p.Q[i * 4 + 1] * ma[1] + p.Q[i * 4 + 2] * ma[2] + p.Q[i * 4 + 3] * ma[3];
      }
      if (b < n) {
        const int r = codes[(size_t)b * p.S_pad];
        const bool simple = __all_sync(0xffffffffu, r < 4);
        tip_columns(ptab + b * 16, r, simple, mb);
        tip_columns(qptab + b * 16, r, simple, qb);
#pragma unroll
        for (int i = 0; i < 4; ++i) mb[i] = clip01(mb[i]);
      } else {
        const double* src = myscr + (size_t)((op.y >> 16) & 0xffff) * 4 * TILE;
#pragma unroll
        for (int i = 0; i < 4; ++i) mb[i] = __ldcg(src + i * TILE);
#pragma unroll
        for (int i = 0; i < 4; ++i) qb[i] = p.Q[i * 4 + 0] * mb[0] + p.Q[i * 4 + 1] * mb[1] + p.Q[i * 4 + 2] * mb[2] + p.Q[i * 4 + 3] * mb[3];
      }
      double wa[4], wb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wa[i] = pre[i] * mb[i], wb[i] = pre[i] * ma[i];
      const double den = (wa[0] * ma[0] + wa[1] * ma[1]) + (wa[2] * ma[2] + wa[3] * ma[3]);
      const double rden = cnt / den;
      const double ga = ((wa[0] * qa[0] + wa[1] * qa[1]) + (wa[2] * qa[2] + wa[3] * qa[3])) * rden;
      const double gb = ((wb[0] * qb[0] + wb[1] * qb[1]) + (wb[2] * qb[2] + wb[3] * qb[3])) * rden;
      const int gslot = (2 * step) & (GCH - 1);
      gst[gslot * GROW + tid] = ga;
      gst[(gslot + 1) * GROW + tid] = gb;

      if (a >= n) {
        double M[16];
        load16(ptab + a * 16, M);
        double B[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) B[v] = clip01(wa[0] * M[0 * 4 + v] + wa[1] * M[1 * 4 + v] + wa[2] * M[2 * 4 + v] + wa[3] * M[3 * 4 + v]);
        const double inv = 1.0 / ((B[0] + B[1]) + (B[2] + B[3]));
        double* dst = stk + (size_t)((op.z >> 8) & 0xff) * 4 * TILE + tid;
#pragma unroll
        for (int v = 0; v < 4; ++v) dst[v * TILE] = B[v] * inv;
      }
      if (b >= n) {
        double M[16];
        load16(ptab + b * 16, M);
        double B[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) B[v] = clip01(wb[0] * M[0 * 4 + v] + wb[1] * M[1 * 4 + v] + wb[2] * M[2 * 4 + v] + wb[3] * M[3 * 4 + v]);
        const double inv = 1.0 / ((B[0] + B[1]) + (B[2] + B[3]));
        double* dst = stk + (size_t)((op.z >> 16) & 0xff) * 4 * TILE + tid;
#pragma unroll
        for (int v = 0; v < 4; ++v) dst[v * TILE] = B[v] * inv;
      }

      const bool last = step == n - 2;
      if (gslot + 2 == GCH || last) {
        // transposed reduction of the staged chunk over the tile's sites
        const int nval = gslot + 2;  // valid staged rows
        __syncthreads();
        const int r = tid & (GCH - 1), q = tid / GCH;
        double acc = 0.0;
        if (r < nval) {
          const double* row = gst + r * GROW + q * GCH;
#pragma unroll
          for (int k = 0; k < GCH; ++k) acc += row[k];
        }
        red[q * GCH + r] = acc;
        __syncthreads();
        if (tid < nval) {
          double tot = 0.0;
#pragma unroll
          for (int k = 0; k < TILE / GCH; ++k) tot += red[k * GCH + tid];
          gout[2 * step + 2 - nval + tid] = tot;
        }
      }
    }
    __syncthreads();  // red / gst are reused by the next item
  }
}

// P(t) = U diag(exp(d t)) Uinv per (unit, node) (substitution.py:61-62), and Q P(t) for tip branches.
__global__ void transition_kernel(const double* __restrict__ bl, int U, int n, vbsky_subst_model mdl,
                                  double* __restrict__ ptab, double* __restrict__ qptab) {
  const int nb = 2 * n - 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)U * nb) return;
  const int u = (int)(idx / nb), c = (int)(idx - (long long)u * nb);
  const double t = bl[idx];
  double e[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) e[k] = exp(t * mdl.d[k]);
  double* out = ptab + idx * 16;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += mdl.P[i * 4 + k] * (e[k] * mdl.Pinv[k * 4 + j]);
      out[i * 4 + j] = s;
    }
  if (c < n && qptab != nullptr) {
    double* qo = qptab + ((size_t)u * n + c) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += mdl.P[i * 4 + k] * (mdl.d[k] * e[k] * mdl.Pinv[k * 4 + j]);
        qo[i * 4 + j] = s;
      }
  }
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
  int grid = 0, depth = 0, tiles = 0;
  size_t smem = 0;
  size_t off_ptab = 0, off_qptab = 0, off_scratch = 0, off_llpart = 0, off_gpart = 0, total = 0;
};

static thread_local int64_t g_stats[8] = {0};

static int prune_config(const vbsky_layout* lay, const vbsky_tips* tips, int U, bool want_grad, PruneConfig* cfg) {
  const int n = lay->n;
  cfg->depth = want_grad ? std::max(lay->depth_up, lay->depth_down) : lay->depth_up;
  cfg->smem = ((size_t)cfg->depth * 4 * TILE + (size_t)GCH * GROW + TILE) * sizeof(double);
  cfg->tiles = (int)((tips->S + TILE - 1) / TILE);
  int dev = 0, sms = 0, per_sm = 0;
  VB_CUDA(cudaGetDevice(&dev));
  VB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  size_t max_smem = 0;
  {
    int v = 0;
    VB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_smem = (size_t)v;
  }
  if (cfg->smem > max_smem)
    return set_error(VBSKY_EINVAL, "tree needs %d live vectors (%zu B shared memory) but the device offers %zu B", cfg->depth, cfg->smem, max_smem);
  if (want_grad) {
    VB_CUDA(cudaFuncSetAttribute(prune_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prune_kernel<true>, TILE, cfg->smem));
  } else {
    VB_CUDA(cudaFuncSetAttribute(prune_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
    VB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prune_kernel<false>, TILE, cfg->smem));
  }
  if (per_sm < 1) return set_error(VBSKY_ECUDA, "pruning kernel cannot be resident (smem %zu B)", cfg->smem);
  const long long items = (long long)U * cfg->tiles;
  cfg->grid = (int)std::min<long long>((long long)sms * per_sm, std::max<long long>(items, 1));
  auto align = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off = 0;
  cfg->off_ptab = off, off = align(off + (size_t)U * (2 * n - 2) * 16 * sizeof(double));
  cfg->off_qptab = off, off = align(off + (want_grad ? (size_t)U * n * 16 * sizeof(double) : 0));
  cfg->off_scratch = off, off = align(off + (want_grad ? (size_t)sms * per_sm * (n - 1) * 4 * TILE * sizeof(double) : 0));
  cfg->off_llpart = off, off = align(off + (size_t)U * cfg->tiles * sizeof(double));
  cfg->off_gpart = off, off = align(off + (want_grad ? (size_t)U * cfg->tiles * (2 * n - 2) * sizeof(double) : 0));
  cfg->total = off;
  return VBSKY_OK;
}

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_prune_workspace_bytes(const vbsky_layout* lay, const vbsky_tips* tips, int32_t U, int32_t want_grad) {
  if (!lay || !tips || U < 1) return 0;
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
  PruneParams p;
  p.up = lay->d_up, p.down = lay->d_down;
  p.codes = tips->d_codes, p.counts = tips->d_counts;
  p.unit_tree = unit_tree;
  p.ptab = (double*)(base + cfg.off_ptab);
  p.qptab = want_grad ? (double*)(base + cfg.off_qptab) : nullptr;
  p.scratch = (double*)(base + cfg.off_scratch);
  p.site_ll = site_ll;
  p.ll_part = (double*)(base + cfg.off_llpart);
  p.g_part = (double*)(base + cfg.off_gpart);
  vbsky_subst_rate_matrix(model, p.Q);
  for (int i = 0; i < 4; ++i) p.pi[i] = model->pi[i];
  p.n = n, p.U = U, p.tiles = cfg.tiles, p.depth = cfg.depth;
  p.S = tips->S, p.S_pad = tips->S_pad;

  {
    const long long tot = (long long)U * (2 * n - 2);
    transition_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, stream>>>(bl, U, n, *model, (double*)p.ptab, (double*)p.qptab);
    VB_CUDA(cudaGetLastError());
  }
  if (want_grad)
    prune_kernel<true><<<cfg.grid, TILE, cfg.smem, stream>>>(p);
  else
    prune_kernel<false><<<cfg.grid, TILE, cfg.smem, stream>>>(p);
  VB_CUDA(cudaGetLastError());
  {
    dim3 grid((unsigned)U, (unsigned)((2 * n - 2 + 127) / 128));
    finalize_kernel<<<grid, 128, 0, stream>>>(p.ll_part, p.g_part, unit_tree, lay->d_down_nodes, U, n, cfg.tiles, ll, dll_dbl);
    VB_CUDA(cudaGetLastError());
  }
  g_stats[0] = 3, g_stats[1] = cfg.grid, g_stats[2] = TILE, g_stats[3] = (int64_t)cfg.smem;
  {
    // algorithmic HBM bytes of prune_kernel: every non-root internal message written once and
    // read once, tip codes read once per pass, counts, per-site ll out, per-tile gradient partials out.
    const double S = (double)tips->S;
    double bytes = (double)U * S * n * (want_grad ? 2.0 : 1.0) + (double)U * S * 8.0 + (site_ll ? (double)U * S * 8.0 : 0.0);
    if (want_grad) bytes += (double)U * S * (n - 2) * 32.0 * 2.0 + (double)U * cfg.tiles * (2.0 * n - 2) * 8.0;
    g_stats[4] = (int64_t)bytes;
  }
  return VBSKY_OK;
}

extern "C" int vbsky_prune_last_stats(int64_t out[8]) {
  VB_CHECK_ARG(out, "NULL argument");
  for (int i = 0; i < 8; ++i) out[i] = g_stats[i];
  return VBSKY_OK;
}
