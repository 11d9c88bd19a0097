// Starting trees on the device: what SeqData.sample_trees does per subsample
// (/root/reference/vbsky/fasta.py:255-284) -- pairwise SNP distances (the external `snp-dists -b`), the
// serial-sample correction of get_distance_matrix (upgma.py:25-62) and Biopython's
// DistanceTreeConstructor.upgma (fasta.py:283) -- batched over the T subsamples of an ensemble.
//
//   snp_pack_kernel     chars [N][L] -> four bit planes (A, C, G, T) per sequence, once per alignment
//   snp_dists_kernel    one warp per (subsample, pair): popcount of (valid_a & valid_b & ~same), coalesced words
//   supgma_dist_kernel  one CTA per subsample: omega by the within-class centred normal equations, then
//                       D[r][c] = y + omega * (t_r + t_c - 2 min t)
//   upgma_kernel        one CTA per subsample: n-1 rounds of (argmin with Biopython's tie rule, average, retire)
//
// Integer outputs are bit-exact with the restated algorithms; omega agrees to rounding (it is a least-squares
// solution computed by a different, closed-form route).
#include "common.h"

namespace vbsky {

__global__ void snp_pack_kernel(const char* __restrict__ seqs, long long N, long long L, int W, uint32_t* __restrict__ planes) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * W) return;
  const long long row = idx / W;
  const int w = (int)(idx % W);
  const char* s = seqs + row * L + (long long)w * 32;
  const int len = (int)min(32LL, L - (long long)w * 32);
  uint32_t a = 0, c = 0, g = 0, t = 0;
  for (int b = 0; b < len; ++b) {
    char ch = s[b];
    if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);  // snp-dists upper-cases unless -k
    a |= (uint32_t)(ch == 'A') << b;
    c |= (uint32_t)(ch == 'C') << b;
    g |= (uint32_t)(ch == 'G') << b;
    t |= (uint32_t)(ch == 'T') << b;
  }
  uint32_t* o = planes + row * 4 * W + w;
  o[0] = a, o[W] = c, o[2 * W] = g, o[3 * W] = t;
}

__global__ void __launch_bounds__(256) snp_dists_kernel(const uint32_t* __restrict__ planes, int W, const int32_t* __restrict__ rows,
                                                        int n, long long npairs, int32_t* __restrict__ dist) {
  const int t = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= npairs) return;
  // p -> (r, c), r > c, in np.tril_indices(n, -1) order: p = r (r-1)/2 + c
  int r = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
  while ((long long)r * (r - 1) / 2 > p) --r;
  while ((long long)(r + 1) * r / 2 <= p) ++r;
  const int c = (int)(p - (long long)r * (r - 1) / 2);
  const uint32_t* A = planes + (size_t)rows[(size_t)t * n + r] * 4 * W;
  const uint32_t* B = planes + (size_t)rows[(size_t)t * n + c] * 4 * W;
  int cnt = 0;
  for (int w = lane; w < W; w += 32) {
    const uint32_t a0 = A[w], a1 = A[W + w], a2 = A[2 * W + w], a3 = A[3 * W + w];
    const uint32_t b0 = B[w], b1 = B[W + w], b2 = B[2 * W + w], b3 = B[3 * W + w];
    const uint32_t valid = (a0 | a1 | a2 | a3) & (b0 | b1 | b2 | b3);
    const uint32_t same = (a0 & b0) | (a1 & b1) | (a2 & b2) | (a3 & b3);
    cnt += __popc(valid & ~same);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) {
    int32_t* d = dist + (size_t)t * n * n;
    d[(size_t)r * n + c] = cnt;
    d[(size_t)c * n + r] = cnt;
    if (c == 0) d[(size_t)r * n + r] = 0;
    if (p == 0) d[0] = 0;
  }
}

constexpr int TB = 1024;
constexpr int kUpgmaPosPerThread = 4;  // upgma_kernel handles n <= 4 * TB tips

// Sum of v over the CTA in a fixed order (thread-id tree), result broadcast.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double s = 0;
  for (int i = 0; i < TB / 32; ++i) s += red[i];
  return s;
}

struct SupgmaParams {
  const int32_t* dist;  // [T][n][n]
  const double* times;  // [T][n]
  double* D;            // [T][n][n] out
  double* omega;        // [T] out
  int n, single_theta;
};

// get_distance_matrix (upgma.py:25-62).  With the design [one-hot(class of the later sample) | dt] the least-squares
// slope is the within-class centred ratio Sxy / Sxx (one global class when single_theta); a pair (r, c), r > c, belongs
// to r when t_r >= t_c and to c otherwise (upgma.py:46), and its class is that sample's time value.
__global__ void __launch_bounds__(TB) supgma_dist_kernel(const SupgmaParams p) {
  extern __shared__ double sh[];
  const int n = p.n, tid = threadIdx.x, t = blockIdx.x;
  double* tm = sh;              // [n]
  double* o_cnt = tm + n;       // per owner: pairs, sum dt, sum y, then class means
  double* o_sdt = o_cnt + n;
  double* o_sy = o_sdt + n;
  double* m_dt = o_sy + n;
  double* m_y = m_dt + n;
  __shared__ double red[TB / 32];
  const int32_t* dist = p.dist + (size_t)t * n * n;
  double* D = p.D + (size_t)t * n * n;
  for (int i = tid; i < n; i += TB) tm[i] = p.times[(size_t)t * n + i];
  __syncthreads();
  double tmin = tm[0];
  for (int i = 1; i < n; ++i) tmin = fmin(tmin, tm[i]);
  auto owns = [&](int i, int j) { return tm[i] > tm[j] || (tm[i] == tm[j] && i > j); };
  for (int i = tid; i < n; i += TB) {
    double cnt = 0, sdt = 0, sy = 0;
    for (int j = 0; j < n; ++j)
      if (j != i && owns(i, j)) cnt += 1, sdt += fabs(tm[i] - tm[j]), sy += (double)dist[(size_t)max(i, j) * n + min(i, j)];
    o_cnt[i] = cnt, o_sdt[i] = sdt, o_sy[i] = sy;
  }
  __syncthreads();
  for (int i = tid; i < n; i += TB) {
    double cnt = 0, sdt = 0, sy = 0;
    for (int k = 0; k < n; ++k)
      if (p.single_theta || tm[k] == tm[i]) cnt += o_cnt[k], sdt += o_sdt[k], sy += o_sy[k];
    m_dt[i] = cnt > 0 ? sdt / cnt : 0.0;
    m_y[i] = cnt > 0 ? sy / cnt : 0.0;
  }
  __syncthreads();
  double sxy = 0, sxx = 0, sx2 = 0;
  for (int i = tid; i < n; i += TB)
    for (int j = 0; j < n; ++j)
      if (j != i && owns(i, j)) {
        const double dt = fabs(tm[i] - tm[j]);
        const double x = dt - m_dt[i];
        const double y = (double)dist[(size_t)max(i, j) * n + min(i, j)] - m_y[i];
        sxy += x * y, sxx += x * x, sx2 += dt * dt;
      }
  sxy = block_sum(sxy, red);
  sxx = block_sum(sxx, red);
  sx2 = block_sum(sx2, red);
  double om;
  if (sxx > 1e-24 * sx2) {
    om = sxy / sxx;
  } else if (p.single_theta || sx2 == 0.0) {
    om = 0.0;  // constant regressor: the minimum-norm solution puts nothing on it
  } else {
    // dt is constant within every class: minimum-norm solution of beta_c + omega a_c = ybar_c
    double num = 0, den = 1;
    for (int i = 0; i < n; ++i) {
      bool first = true;  // visit every distinct time value once
      for (int k = 0; k < i && first; ++k) first = tm[k] != tm[i];
      if (!first) continue;
      double cnt = 0;
      for (int k = 0; k < n; ++k)
        if (tm[k] == tm[i]) cnt += o_cnt[k];
      if (cnt > 0) num += m_dt[i] * m_y[i], den += m_dt[i] * m_dt[i];
    }
    om = num / den;
  }
  if (tid == 0) p.omega[t] = om;
  for (int idx = tid; idx < n * n; idx += TB) {
    const int r = idx / n, c = idx % n;
    double v = 0.0;
    if (r != c) {
      const int hi = max(r, c), lo = min(r, c);
      const double shift = __dsub_rn(__dadd_rn(tm[hi], tm[lo]), __dmul_rn(2.0, tmin));
      v = __dadd_rn((double)dist[(size_t)hi * n + lo], __dmul_rn(om, shift));
    }
    D[idx] = v;
  }
}

struct UpgmaParams {
  double* D;         // [T][n][n] symmetric working copy (destroyed)
  int32_t* merges;   // [T][n-1][2]
  double* heights;   // [T][n-1]
  int n;
};

// DistanceTreeConstructor.upgma.  Biopython deletes row/column min_i and keeps the new clade at position min_j, so the
// positions of the survivors keep their relative order: scanning "for i: for j < i" with `min_dist >= dm[i, j]` picks,
// among equal minima, the largest i and then the largest j in slot order.  Slots are retired instead of deleted.
__global__ void __launch_bounds__(TB) upgma_kernel(const UpgmaParams p) {
  extern __shared__ int shi[];
  const int n = p.n, tid = threadIdx.x, t = blockIdx.x, lane = tid & 31, w = tid >> 5;
  int* alive = shi;      // [n] slot ids still in the matrix, ascending
  int* cid = alive + n;  // [n] clade id held by a slot
  __shared__ double red_d[TB / 32];
  __shared__ int red_k[TB / 32];
  __shared__ int best_k;
  double* D = p.D + (size_t)t * n * n;
  for (int i = tid; i < n; i += TB) alive[i] = i, cid[i] = i;
  __syncthreads();
  int m = n;
  for (int round = 0; round < n - 1; ++round, --m) {
    // candidates: positions (a, b), b < a < m  ->  slots (alive[a], alive[b]); key = alive[a] * n + alive[b]
    double bd = 0.0;
    int bk = -1;
    const int total = m * m;
    for (int idx = tid; idx < total; idx += TB) {
      const int a = idx / m, b = idx - a * m;
      if (b >= a) continue;
      const int i = alive[a], j = alive[b];
      const double d = D[(size_t)i * n + j];
      const int k = i * n + j;
      if (bk < 0 || d < bd || (d == bd && k > bk)) bd = d, bk = k;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (ok >= 0 && (bk < 0 || od < bd || (od == bd && ok > bk))) bd = od, bk = ok;
    }
    if (lane == 0) red_d[w] = bd, red_k[w] = bk;
    __syncthreads();
    if (tid == 0) {
      for (int x = 1; x < TB / 32; ++x) {
        const double od = red_d[x];
        const int ok = red_k[x];
        if (ok >= 0 && (bk < 0 || od < bd || (od == bd && ok > bk))) bd = od, bk = ok;
      }
      best_k = bk;
      const int i = bk / n, j = bk % n;
      p.merges[((size_t)t * (n - 1) + round) * 2 + 0] = cid[i];
      p.merges[((size_t)t * (n - 1) + round) * 2 + 1] = cid[j];
      p.heights[(size_t)t * (n - 1) + round] = bd * 1.0 / 2;
    }
    __syncthreads();
    const int i = best_k / n, j = best_k % n;
    for (int a = tid; a < m; a += TB) {
      const int k = alive[a];
      if (k != i && k != j) {
        const double v = __dadd_rn(D[(size_t)i * n + k], D[(size_t)j * n + k]) * 1.0 / 2;
        D[(size_t)j * n + k] = v;
        D[(size_t)k * n + j] = v;
      }
    }
    __syncthreads();
    // retire slot i: shift the tail of the (ascending) position list down by one (read everything, then write)
    int nv[kUpgmaPosPerThread];
#pragma unroll
    for (int k = 0; k < kUpgmaPosPerThread; ++k) {
      const int a = tid + k * TB;
      nv[k] = (a < m - 1 && alive[a] >= i) ? alive[a + 1] : -1;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kUpgmaPosPerThread; ++k)
      if (nv[k] >= 0) alive[tid + k * TB] = nv[k];
    if (tid == 0) cid[j] = n + round;
    __syncthreads();
  }
}

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_snp_planes_bytes(int64_t N, int64_t L) {
  return (size_t)N * 4 * (size_t)((L + 31) / 32) * sizeof(uint32_t);
}

extern "C" int vbsky_snp_pack_device(const char* seqs, int64_t N, int64_t L, uint32_t* planes, void* stream) {
  VB_CHECK_ARG(seqs && planes, "NULL argument");
  VB_CHECK_ARG(N >= 1 && L >= 1, "bad shape N=%lld L=%lld", (long long)N, (long long)L);
  const int W = (int)((L + 31) / 32);
  const long long total = N * W;
  snp_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seqs, N, L, W, planes);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_snp_dists_device(const uint32_t* planes, int64_t N, int64_t L, const int32_t* rows, int32_t T, int32_t n,
                                      int32_t* dist, void* stream) {
  VB_CHECK_ARG(planes && rows && dist, "NULL argument");
  VB_CHECK_ARG(N >= 1 && L >= 1 && T >= 1 && n >= 1 && T <= 65535, "bad shape N=%lld L=%lld T=%d n=%d", (long long)N, (long long)L, T, n);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 1) {
    VB_CUDA(cudaMemsetAsync(dist, 0, (size_t)T * sizeof(int32_t), st));
    return VBSKY_OK;
  }
  const long long npairs = (long long)n * (n - 1) / 2;
  dim3 grid((unsigned)((npairs + 7) / 8), (unsigned)T);
  snp_dists_kernel<<<grid, 256, 0, st>>>(planes, (int)((L + 31) / 32), rows, n, npairs, dist);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_supgma_distances_device(int32_t T, int32_t n, const int32_t* dist, const double* times, int32_t single_theta,
                                             double* D, double* omega, void* stream) {
  VB_CHECK_ARG(dist && times && D && omega, "NULL argument");
  VB_CHECK_ARG(T >= 1 && n >= 2 && n <= 4096, "bad shape T=%d n=%d", T, n);
  SupgmaParams p{dist, times, D, omega, n, single_theta ? 1 : 0};
  const size_t smem = (size_t)6 * n * sizeof(double);
  VB_CUDA(cudaFuncSetAttribute(supgma_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  supgma_dist_kernel<<<T, TB, smem, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_upgma_device(int32_t T, int32_t n, double* D, int32_t* merges, double* heights, void* stream) {
  VB_CHECK_ARG(D && merges && heights, "NULL argument");
  VB_CHECK_ARG(T >= 1 && n >= 2 && n <= kUpgmaPosPerThread * TB, "bad shape T=%d n=%d (n <= %d)", T, n, kUpgmaPosPerThread * TB);
  UpgmaParams p{D, merges, heights, n};
  upgma_kernel<<<T, TB, (size_t)2 * n * sizeof(int), (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}
