// Per-unit "chain" kernels around the pruning kernel (one CTA per (tree, sample) unit):
//   chain_forward_kernel   params -> node heights (tree_data.py:287-367), branch lengths x clock rate
//                          (bdsky.py:298-307,334), xs / times (bdsky.py:312-322), lam/psi/mu (bdsky.py:219-227)
//   bdsky_kernel           the birth-death-skyline tree density of bdsky.py:18-216 (Stadler et al. 2013, Thm 1)
//                          and its hand-derived gradient w.r.t. lam, psi, mu, rho, xs, times
//   chain_backward_kernel  adjoint of chain_forward: d/d(proportions, root_proportion, origin, clock_rate,
//                          R, delta, s) from d/d(branch lengths), d/d(xs), d/d(times), d/d(lam, psi, mu)
// The reference gets these derivatives from JAX autodiff through lax.scan; here they are explicit.
// Work is O(n + m) per unit and negligible next to pruning; the kernels favour fixed summation orders
// (bitwise reproducible results) over peak speed.
#include <cmath>

#include "common.h"

namespace vbsky {

constexpr int CT = 128;  // threads per unit

__device__ __forceinline__ double block_sum(double v, double* red) {
  // fixed-shape tree reduction over CT threads; result broadcast to all threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < CT / 32; ++w) tot += red[w];
  __syncthreads();
  return tot;
}

struct ChainParams {
  // layout
  const int32_t* unit_tree;
  const int32_t* child_parent;   // [T][N]
  const int32_t* parent_child;   // [T][N][2]
  const int32_t* level_nodes;    // [T][n-1]
  const int32_t* level_start;    // [T][max_levels+1]
  const double* lst;             // [T][N]
  const double* sample_times;    // [T][n]
  const double* max_sample_time; // [T]
  int n, m, max_levels, U;
  // inputs [U][...]
  const double *proportions, *root_proportion, *origin, *origin_start, *clock_rate, *R, *delta, *s;
  const double* grid;  // [U][m+1] or null (equidistant skyline intervals)
  // forward outputs
  double *heights, *bl_scaled, *xs, *times, *lam, *psi, *mu;
  // backward inputs (any may be null => treated as zero)
  const double *d_bl, *d_xs, *d_times, *d_lam, *d_psi, *d_mu;
  // backward outputs
  double *g_proportions, *g_root_proportion, *g_origin, *g_clock_rate, *g_R, *g_delta, *g_s, *g_grid;
  double* dh;  // workspace [U][N]
};

__global__ void __launch_bounds__(CT) chain_forward_kernel(const ChainParams p) {
  const int u = blockIdx.x, tid = threadIdx.x;
  const int n = p.n, N = 2 * n - 1, m = p.m;
  const int t = p.unit_tree[u];
  const int32_t* cp = p.child_parent + (size_t)t * N;
  const double* lst = p.lst + (size_t)t * N;
  double* h = p.heights + (size_t)u * N;
  const double tm = p.origin[u] + p.origin_start[u];
  const double mst = p.max_sample_time[t];
  // bdsky.py:298-300: root_height = root_proportion * (origin + origin_start - max(sample_times))
  const double root_height = p.root_proportion[u] * (tm - mst);
  for (int i = tid; i < n; i += CT) h[i] = lst[i];
  if (tid == 0) h[N - 1] = root_height + mst;  // tree_data.py:355
  __syncthreads();
  const int32_t* ln = p.level_nodes + (size_t)t * (n - 1);
  const int32_t* ls = p.level_start + (size_t)t * (p.max_levels + 1);
  const double* prop = p.proportions + (size_t)u * (n - 2);
  for (int l = 1; l < p.max_levels; ++l) {
    const int b = ls[l], e = ls[l + 1];
    for (int j = b + tid; j < e; j += CT) {
      const int i = ln[j];
      const double pr = prop[i - n];
      h[i] = (1.0 - pr) * lst[i] + pr * h[cp[i]];  // tree_data.py:335
    }
    __syncthreads();
  }
  const double clock = p.clock_rate[u];
  double* bl = p.bl_scaled + (size_t)u * (N - 1);
  double* xs = p.xs ? p.xs + (size_t)u * N : nullptr;
  for (int c = tid; c < N; c += CT) {
    const double hc = h[c];
    if (c < N - 1) bl[c] = (h[cp[c]] - hc) * clock;  // bdsky.py:307,334
    if (xs) xs[c] = tm - hc;                         // bdsky.py:322
  }
  if (p.times) {
    double* times = p.times + (size_t)u * (m + 1);
    // jnp.linspace(0, tm, m+1): 0*(1-i/m) + tm*(i/m) for i < m, endpoint exactly tm (bdsky.py:316)
    if (p.grid) {
      // prespecified grid (bdsky.py:317-319): times = tm - grid, times[0] = 0
      const double* grid = p.grid + (size_t)u * (m + 1);
      for (int i = tid; i <= m; i += CT) times[i] = i == 0 ? 0.0 : tm - grid[i];
    } else {
      for (int i = tid; i <= m; i += CT) times[i] = i < m ? tm * ((double)i / (double)m) : tm;
    }
    for (int i = tid; i < m; i += CT) {
      const double d = p.delta[(size_t)u * m + i];
      const double psi = p.s[(size_t)u * m + i] * d;
      p.lam[(size_t)u * m + i] = p.R[(size_t)u * m + i] * d;  // bdsky.py:223-225
      p.psi[(size_t)u * m + i] = psi;
      p.mu[(size_t)u * m + i] = d - psi;
    }
  }
}

__global__ void __launch_bounds__(CT) chain_backward_kernel(const ChainParams p) {
  __shared__ double red[CT / 32];
  const int u = blockIdx.x, tid = threadIdx.x;
  const int n = p.n, N = 2 * n - 1, m = p.m;
  const int t = p.unit_tree[u];
  const int32_t* cp = p.child_parent + (size_t)t * N;
  const double* lst = p.lst + (size_t)t * N;
  const double* h = p.heights + (size_t)u * N;
  const double* dbl = p.d_bl ? p.d_bl + (size_t)u * (N - 1) : nullptr;
  const double* dxs = p.d_xs ? p.d_xs + (size_t)u * N : nullptr;
  double* dh = p.dh + (size_t)u * N;
  const double clock = p.clock_rate[u];
  const double tm = p.origin[u] + p.origin_start[u];
  const double mst = p.max_sample_time[t];

  // d/d clock_rate and the direct dependence on tm (xs = tm - h, times = tm * i/m)
  double acc_clock = 0.0, acc_tm = 0.0;
  for (int c = tid; c < N; c += CT) {
    if (dbl && c < N - 1) acc_clock += dbl[c] * (h[cp[c]] - h[c]);
    if (dxs) acc_tm += dxs[c];
  }
  if (p.d_times) {
    if (p.grid) {
      for (int i = tid; i <= m; i += CT) {
        const double d = i == 0 ? 0.0 : p.d_times[(size_t)u * (m + 1) + i];
        acc_tm += d;
        if (p.g_grid) p.g_grid[(size_t)u * (m + 1) + i] = -d;
      }
    } else {
      for (int i = tid; i <= m; i += CT) acc_tm += p.d_times[(size_t)u * (m + 1) + i] * (i < m ? (double)i / (double)m : 1.0);
    }
  } else if (p.grid && p.g_grid) {
    for (int i = tid; i <= m; i += CT) p.g_grid[(size_t)u * (m + 1) + i] = 0.0;
  }
  const double g_clock = block_sum(acc_clock, red);
  double g_tm = block_sum(acc_tm, red);

  // adjoint of the heights: total(v) = -dxs[v] - g_v + sum_children [ g_c + (c internal ? total(c) * p_c : 0) ],
  // g_c = d/d(bl_scaled[c]) * clock.  Children sit on deeper levels, so deepest level first.
  const int32_t* ln = p.level_nodes + (size_t)t * (n - 1);
  const int32_t* ls = p.level_start + (size_t)t * (p.max_levels + 1);
  const double* prop = p.proportions + (size_t)u * (n - 2);
  double* gprop = p.g_proportions + (size_t)u * (n - 2);
  const int32_t* kids = p.parent_child + (size_t)t * N * 2;  // [N][2]
  for (int c = tid; c < N; c += CT) dh[c] = (c >= n && dxs) ? -dxs[c] : 0.0;
  __syncthreads();
  for (int l = p.max_levels - 1; l >= 0; --l) {
    const int b = ls[l], e = ls[l + 1];
    for (int j = b + tid; j < e; j += CT) {
      const int v = ln[j];
      double tot = dh[v];
      if (dbl && v < N - 1) tot -= dbl[v] * clock;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = kids[2 * v + k];
        if (dbl) tot += dbl[c] * clock;
        if (c >= n) tot += dh[c] * prop[c - n];
      }
      dh[v] = tot;
      if (v < N - 1) gprop[v - n] = tot * (h[cp[v]] - lst[v]);  // h_v = (1-p) lst + p h_pa
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double d_rh = dh[N - 1];
    p.g_root_proportion[u] = d_rh * (tm - mst);
    g_tm += d_rh * p.root_proportion[u];
    p.g_origin[u] = g_tm;  // d/d origin == d/d origin_start
    p.g_clock_rate[u] = g_clock;
  }
  if (p.g_R) {
    for (int i = tid; i < m; i += CT) {
      const size_t k = (size_t)u * m + i;
      const double dl = p.d_lam ? p.d_lam[k] : 0.0, dp = p.d_psi ? p.d_psi[k] : 0.0, dm = p.d_mu ? p.d_mu[k] : 0.0;
      const double d = p.delta[k], s = p.s[k], R = p.R[k];
      p.g_R[k] = dl * d;
      p.g_delta[k] = dl * R + dp * s + dm * (1.0 - s);
      p.g_s[k] = (dp - dm) * d;
    }
  }
}

// ------------------------------------------------------------------------------------------ BDSKY
struct BdskyParams {
  const double *lam, *psi, *mu, *rho;  // [U][m]
  const double* xs;                    // [U][N]
  const double* times;                 // [U][m+1]
  double* lp;                          // [U]
  double *dlam, *dpsi, *dmu, *drho;    // [U][m] (null => value only)
  double* dxs;                         // [U][N]
  double* dtimes;                      // [U][m+1]
  double* node_ws;                     // [U][3][N] per-node adjoint pieces
  int32_t* node_iv;                    // [U][N] interval index of each node
  int n, m, U, condition_on_survival;
};

__device__ __forceinline__ bool isclose_d(double a, double b) { return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b); }

struct LogQ {
  double val, dz, dB;  // value, d/dz (z = A (t - t_i)), d/dB
};
__device__ __forceinline__ LogQ log_q(double A, double B, double t, double ti) {
  // bdsky.py:136-143
  const double z = A * (t - ti);
  const double E = exp(z);
  const double D = E * (1.0 - B) + (1.0 + B);
  LogQ r;
  r.val = 1.3862943611198906 + z - 2.0 * log(fabs(D));
  r.dz = 1.0 - 2.0 * E * (1.0 - B) / D;
  r.dB = -2.0 * (1.0 - E) / D;
  return r;
}

__global__ void __launch_bounds__(CT) bdsky_kernel(const BdskyParams p) {
  extern __shared__ double sm[];
  __shared__ double red[CT / 32];
  __shared__ double s_p1, s_logit_p1;
  const int u = blockIdx.x, tid = threadIdx.x;
  const int n = p.n, N = 2 * n - 1, m = p.m;
  const bool want_grad = p.dlam != nullptr;
  // shared arrays (doubles)
  double* times = sm;             // m+1
  double* lam = times + (m + 1);  // m each below
  double* psi = lam + m;
  double* mu = psi + m;
  double* rho = mu + m;
  double* A = rho + m;
  double* B = A + m;
  double* W = B + m;      // exp(-A dt)
  double* F = W + m;      // f
  double* PIN = F + m;    // p carried into step i
  double* PRAW = PIN + m; // unclipped p_i
  double* Abar = PRAW + m;
  double* Bbar = Abar + m;
  double* lbar = Bbar + m;
  double* pbar = lbar + m;
  double* mbar = pbar + m;
  double* rbar = mbar + m;
  double* tbar = rbar + m;       // m+1
  double* tbar2 = tbar + (m + 1);  // m+2 contributions written for index j+1 by thread j
  int* hist = reinterpret_cast<int*>(tbar2 + (m + 2));  // m+1
  int* Ncnt = hist + (m + 1);                           // m

  for (int i = tid; i <= m; i += CT) times[i] = p.times[(size_t)u * (m + 1) + i], hist[i] = 0, tbar[i] = 0.0;
  for (int i = tid; i < m + 2; i += CT) tbar2[i] = 0.0;
  for (int i = tid; i < m; i += CT) {
    const size_t k = (size_t)u * m + i;
    lam[i] = p.lam[k], psi[i] = p.psi[k], mu[i] = p.mu[k], rho[i] = p.rho[k];
    Ncnt[i] = 0;
  }
  __syncthreads();

  // ---- phase A: A_i and the reverse recursion for B_i, p_i (bdsky.py:48-120), sequential in m
  if (tid == 0) {
    double pc = 1.0, lgc = 0.0;
    for (int i = m - 1; i >= 0; --i) {
      const double c = lam[i] - mu[i] - psi[i];
      const double Ai = sqrt(c * c + 4.0 * lam[i] * psi[i]);
      const double Bi = ((1.0 - 2.0 * (1.0 - rho[i]) * pc) * lam[i] + mu[i] + psi[i]) / Ai;
      const double dt = times[i + 1] - times[i];
      const double w = exp(-(Ai * dt));
      const double f = ((1.0 + Bi) - w * (1.0 - Bi)) / ((1.0 + Bi) + w * (1.0 - Bi));
      const double praw = (lam[i] + mu[i] + psi[i] - Ai * f) / (2.0 * lam[i]);
      const double pn = fmin(praw, 1.0);
      const double q = dt * fmin(lam[i] - mu[i], 0.0);
      const double lp = log(pn / (1.0 - pn));
      double lg = lp;
      if (!isfinite(lp) && isclose_d(psi[i], 0.0) && isclose_d(rho[i], 0.0)) {
        lg = lam[i] < mu[i] ? lgc - log(pc) - q - (lam[i] - exp(q) * mu[i]) * (1.0 - pc) / (lam[i] - mu[i])
                            : log(mu[i] / fabs(lam[i] - mu[i]));
      }
      A[i] = Ai, B[i] = Bi, W[i] = w, F[i] = f, PIN[i] = pc, PRAW[i] = praw;
      pc = pn, lgc = lg;
    }
    s_p1 = pc, s_logit_p1 = lgc;
  }
  __syncthreads();

  // ---- phase B: per-node terms l2, l3 (bdsky.py:125-130,188-198)
  const double* xs = p.xs + (size_t)u * N;
  double* ws = want_grad ? p.node_ws + (size_t)u * 3 * N : nullptr;
  int32_t* iv = want_grad ? p.node_iv + (size_t)u * N : nullptr;
  double val = 0.0;
  for (int k = tid; k < N; k += CT) {
    const double x = xs[k];
    int lo = 0, hi = m + 1;  // searchsorted(times, x, 'right')
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (times[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int Ix = lo;
    const int ai = min(max(Ix - 1, 0), m - 1);  // jnp.take clip semantics
    const int ti = min(max(Ix, 0), m);
    const LogQ q = log_q(A[ai], B[ai], x, times[ti]);
    double sgn;
    if (k >= n) {
      val += q.val + log(lam[ai]);
      sgn = 1.0;
      atomicAdd(&hist[min(max(Ix - 1, 0), m)], 1);
    } else {
      bool near = false;
      for (int j = 0; j < m; ++j)
        if (rho[j] > 0.0 && isclose_d(x, times[j + 1])) near = true, atomicAdd(&Ncnt[j], 1);
      sgn = near ? 0.0 : -1.0;
      if (!near) val += log(psi[ai]) - q.val;
      atomicAdd(&hist[min(max(Ix - 1, 0), m)], -1);
    }
    if (want_grad) {
      p.dxs[(size_t)u * N + k] = sgn * q.dz * A[ai];
      ws[k] = sgn * q.dz * (x - times[ti]);  // -> Abar[ai]
      ws[N + k] = sgn * q.dB;                // -> Bbar[ai]
      ws[2 * N + k] = -sgn * q.dz * A[ai];   // -> tbar[ti]
      iv[k] = ai | (ti << 16) | (k >= n ? 0 : (sgn != 0.0 ? (1 << 30) : 0));
    }
  }
  __syncthreads();

  // ---- phase C: per-interval gathers in node order + grid terms l1, l4, l5 (bdsky.py:172-214)
  for (int j = tid; j <= m; j += CT) {
    double a_ = 0.0, b_ = 0.0, l_ = 0.0, ps_ = 0.0, t_ = 0.0, r_ = 0.0;
    if (want_grad) {
      for (int k = 0; k < N; ++k) {
        const int code = iv[k];
        const int ai = code & 0xffff, ti = (code >> 16) & 0x3fff;
        if (ai == j && j < m) {
          a_ += ws[k], b_ += ws[N + k];
          if (k >= n) l_ += 1.0 / lam[j];
          else if (code & (1 << 30)) ps_ += 1.0 / psi[j];
        }
        if (ti == j) t_ += ws[2 * N + k];
      }
    }
    if (j < m) {
      // lineage counts: n_i[j] = 1 + #{internal x < times[j+1]} - #{leaf y < times[j+1]}, n_i[m-1] = 0
      // (util.py:25-71 + bdsky.py:152-158 reduce to this count)
      double nj = 0.0;
      if (j < m - 1) {
        int c = 1;
        for (int e = 0; e <= j; ++e) c += hist[e];
        nj = (double)c;
      }
      if (nj != 0.0) {
        val += nj * log1p(-rho[j]);  // l41 = xlog1py(n_i, -rho)
        r_ += -nj / (1.0 - rho[j]);
      }
      if (Ncnt[j] > 0) {
        val += (double)Ncnt[j] * log(rho[j]);  // l5 = xlogy(N_i, rho)
        r_ += (double)Ncnt[j] / rho[j];
      }
      // l42 term owned by interval j (j >= 1): n_i[j-1] * log q_{j+1}(times[j])  (A[j], B[j], t_i = times[j+1])
      if (j >= 1) {
        int c = 1;
        for (int e = 0; e <= j - 1; ++e) c += hist[e];
        const double njm1 = (double)c;
        const LogQ q = log_q(A[j], B[j], times[j], times[j + 1]);
        val += njm1 * q.val;
        a_ += njm1 * q.dz * (times[j] - times[j + 1]);
        b_ += njm1 * q.dB;
        t_ += njm1 * q.dz * A[j];
        tbar2[j + 1] += -njm1 * q.dz * A[j];
      } else {
        // l1 = log q_1(t_0)  (A[0], B[0], t = times[0], t_i = times[1])
        const LogQ q = log_q(A[0], B[0], times[0], times[1]);
        val += q.val;
        a_ += q.dz * (times[0] - times[1]);
        b_ += q.dB;
        t_ += q.dz * A[0];
        tbar2[1] += -q.dz * A[0];
      }
      Abar[j] = a_, Bbar[j] = b_, lbar[j] = l_, pbar[j] = ps_, mbar[j] = 0.0, rbar[j] = r_;
    }
    tbar[j] = t_;
  }
  __syncthreads();
  double tot = block_sum(val, red);

  // ---- phase D: conditioning on survival and the adjoint of the recursion
  if (tid == 0) {
    double pb = 0.0;  // adjoint of p_1(t_0)
    if (p.condition_on_survival) {
      const double l1_0 = log1p(-s_p1);
      if (isfinite(l1_0)) {
        tot -= l1_0;
        pb = 1.0 / (1.0 - s_p1);
      } else {
        // p_1(t_0) == 1: the reference falls back to the logit recursion (bdsky.py:178-179); its autodiff
        // gives NaN -> nan_to_num -> 0 there, so no gradient is propagated through this term.
        const double lg = s_logit_p1;
        tot -= (lg > 10.0) ? -lg : -log1p(exp(lg));
      }
    }
    p.lp[u] = tot;
    if (want_grad) {
      for (int i = 0; i <= m; ++i) tbar[i] += tbar2[i];
      for (int i = 0; i < m; ++i) {
        const double Ai = A[i], Bi = B[i], w = W[i], f = F[i], pin = PIN[i];
        const double dt = times[i + 1] - times[i];
        double ab = Abar[i], bb = Bbar[i], lb = lbar[i], sb = pbar[i], mb = mbar[i], rb = rbar[i];
        double fb = 0.0;
        if (!(PRAW[i] > 1.0)) {  // min(praw, 1) passes the gradient only when not clipped
          const double il = 1.0 / (2.0 * lam[i]);
          fb += pb * (-Ai * il);
          ab += pb * (-f * il);
          lb += pb * (il - PRAW[i] / lam[i]);
          mb += pb * il;
          sb += pb * il;
        }
        const double a1 = 1.0 + Bi, b1 = 1.0 - Bi, den = a1 + w * b1;
        bb += fb * 4.0 * w / (den * den);
        const double wb = fb * (-2.0 * a1 * b1 / (den * den));
        ab += wb * (-dt * w);
        const double dtb = wb * (-Ai * w);
        tbar[i + 1] += dtb;
        tbar[i] -= dtb;
        pb = bb * (-2.0 * (1.0 - rho[i]) * lam[i] / Ai);
        rb += bb * (2.0 * pin * lam[i] / Ai);
        lb += bb * (1.0 - 2.0 * (1.0 - rho[i]) * pin) / Ai;
        mb += bb / Ai;
        sb += bb / Ai;
        ab += bb * (-Bi / Ai);
        const double c = lam[i] - mu[i] - psi[i];
        lb += ab * (c + 2.0 * psi[i]) / Ai;
        mb += ab * (-c) / Ai;
        sb += ab * (-c + 2.0 * lam[i]) / Ai;
        const size_t k = (size_t)u * m + i;
        p.dlam[k] = lb, p.dpsi[k] = sb, p.dmu[k] = mb, p.drho[k] = rb;
      }
      for (int i = 0; i <= m; ++i) p.dtimes[(size_t)u * (m + 1) + i] = tbar[i];
    }
  }
}

static size_t bdsky_smem(int m) { return (size_t)(18 * m + (m + 1) * 2 + (m + 2)) * sizeof(double) + (size_t)(2 * m + 2) * sizeof(int); }

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_bdsky_workspace_bytes(int32_t U, int32_t n) {
  const size_t N = 2 * (size_t)n - 1;
  return ((size_t)U * 3 * N * sizeof(double) + 255) / 256 * 256 + (size_t)U * N * sizeof(int32_t);
}

extern "C" int vbsky_bdsky_value_and_grad(int32_t U, int32_t n, int32_t m, const double* lam, const double* psi,
                                          const double* mu, const double* rho, const double* xs, const double* times,
                                          int32_t condition_on_survival, double* lp, double* dlam, double* dpsi,
                                          double* dmu, double* drho, double* dxs, double* dtimes, void* ws,
                                          size_t ws_bytes, void* stream) {
  VB_CHECK_ARG(lam && psi && mu && rho && xs && times && lp, "NULL argument");
  VB_CHECK_ARG(U >= 1 && n >= 2 && m >= 1 && m <= 8191, "bad shape U=%d n=%d m=%d", U, n, m);
  const bool want_grad = dlam != nullptr;
  if (want_grad) {
    VB_CHECK_ARG(dpsi && dmu && drho && dxs && dtimes && ws, "gradient outputs and workspace must all be given");
    if (ws_bytes < vbsky_bdsky_workspace_bytes(U, n)) return set_error(VBSKY_ENOMEM, "bdsky workspace too small");
  }
  const size_t N = 2 * (size_t)n - 1;
  BdskyParams p;
  p.lam = lam, p.psi = psi, p.mu = mu, p.rho = rho, p.xs = xs, p.times = times, p.lp = lp;
  p.dlam = dlam, p.dpsi = dpsi, p.dmu = dmu, p.drho = drho, p.dxs = dxs, p.dtimes = dtimes;
  p.node_ws = (double*)ws;
  p.node_iv = ws ? (int32_t*)((char*)ws + ((size_t)U * 3 * N * sizeof(double) + 255) / 256 * 256) : nullptr;
  p.n = n, p.m = m, p.U = U, p.condition_on_survival = condition_on_survival;
  const size_t smem = bdsky_smem(m);
  if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(bdsky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bdsky_kernel<<<U, CT, smem, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

static int fill_chain(const vbsky_layout* lay, int32_t U, const int32_t* unit_tree, int32_t m, const vbsky_params* in,
                      ChainParams* p) {
  VB_CHECK_ARG(lay && unit_tree && in, "NULL argument");
  VB_CHECK_DEVICE(lay);
  VB_CHECK_ARG(U >= 1 && m >= 0, "bad shape U=%d m=%d", U, m);
  VB_CHECK_ARG((in->proportions || lay->n == 2) && in->root_proportion && in->origin && in->origin_start && in->clock_rate, "NULL parameter array");
  *p = ChainParams{};
  p->unit_tree = unit_tree, p->child_parent = lay->d_child_parent, p->parent_child = lay->d_parent_child, p->level_nodes = lay->d_level_nodes;
  p->level_start = lay->d_level_start, p->lst = lay->d_lst, p->sample_times = lay->d_sample_times;
  p->max_sample_time = lay->d_max_sample_time;
  p->n = lay->n, p->m = m, p->max_levels = lay->max_levels, p->U = U;
  p->proportions = in->proportions, p->root_proportion = in->root_proportion, p->origin = in->origin;
  p->origin_start = in->origin_start, p->clock_rate = in->clock_rate, p->R = in->R, p->delta = in->delta, p->s = in->s;
  p->grid = m >= 1 ? in->grid : nullptr;
  return VBSKY_OK;
}

extern "C" int vbsky_heights_forward(const vbsky_layout* lay, int32_t U, const int32_t* unit_tree, int32_t m,
                                     const vbsky_params* in, double* heights, double* bl_scaled, double* xs,
                                     double* times, double* lam, double* psi, double* mu, void* stream) {
  ChainParams p;
  int rc = fill_chain(lay, U, unit_tree, m, in, &p);
  if (rc != VBSKY_OK) return rc;
  VB_CHECK_ARG(heights && bl_scaled, "heights and bl_scaled outputs are required");
  if (times) VB_CHECK_ARG(m >= 1 && in->R && in->delta && in->s && lam && psi && mu, "skyline outputs need m >= 1 and R, delta, s");
  p.heights = heights, p.bl_scaled = bl_scaled, p.xs = xs, p.times = times, p.lam = lam, p.psi = psi, p.mu = mu;
  chain_forward_kernel<<<U, CT, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_heights_backward(const vbsky_layout* lay, int32_t U, const int32_t* unit_tree, int32_t m,
                                      const vbsky_params* in, const double* heights, const double* d_bl_scaled,
                                      const double* d_xs, const double* d_times, const double* d_lam,
                                      const double* d_psi, const double* d_mu, vbsky_param_grads* out, void* ws,
                                      size_t ws_bytes, void* stream) {
  ChainParams p;
  int rc = fill_chain(lay, U, unit_tree, m, in, &p);
  if (rc != VBSKY_OK) return rc;
  VB_CHECK_ARG(heights && out && ws, "NULL argument");
  VB_CHECK_ARG((out->proportions || lay->n == 2) && out->root_proportion && out->origin && out->clock_rate, "NULL gradient output");
  const size_t N = 2 * (size_t)lay->n - 1;
  if (ws_bytes < (size_t)U * N * sizeof(double)) return set_error(VBSKY_ENOMEM, "heights workspace too small");
  p.heights = const_cast<double*>(heights);
  p.d_bl = d_bl_scaled, p.d_xs = d_xs, p.d_times = d_times, p.d_lam = d_lam, p.d_psi = d_psi, p.d_mu = d_mu;
  p.g_proportions = out->proportions, p.g_root_proportion = out->root_proportion, p.g_origin = out->origin;
  p.g_clock_rate = out->clock_rate, p.g_R = out->R, p.g_delta = out->delta, p.g_s = out->s;
  p.g_grid = p.grid ? out->grid : nullptr;
  if (p.g_R) VB_CHECK_ARG(m >= 1 && in->R && in->delta && in->s && out->delta && out->s, "skyline gradients need m >= 1 and R, delta, s");
  p.dh = (double*)ws;
  chain_backward_kernel<<<U, CT, 0, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}
