// Tip data: character encoding, site-pattern compression (host, bit-exact with the reference's
// NumPy calls) and the device-resident packed layout the pruning kernel reads.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.h"

using namespace vbsky;

namespace {

// substitution.py:10-30 (BEAST Nucleotide state sets) packed as bit0=A bit1=C bit2=G bit3=T.
// 0xff = not in the table (reference raises KeyError; lower case is not accepted either).
struct EncodeTable {
  uint8_t t[256];
  EncodeTable() {
    std::memset(t, 0xff, sizeof t);
    t['A'] = 1, t['C'] = 2, t['G'] = 4, t['T'] = 8, t['U'] = 8;
    t['R'] = 1 | 4, t['Y'] = 2 | 8, t['M'] = 1 | 2, t['W'] = 1 | 8, t['S'] = 2 | 4, t['K'] = 4 | 8;
    t['B'] = 2 | 4 | 8, t['D'] = 1 | 4 | 8, t['H'] = 1 | 2 | 8, t['V'] = 1 | 2 | 4;
    t['N'] = 15, t['X'] = 15, t['-'] = 15, t['?'] = 15;
  }
};
const EncodeTable kEnc;

// NumPy's lexicographic row order on the unpacked [n,4] 0/1 rows compares state A first, so the
// per-tip sort key is the bit-reversed mask (A most significant).
inline uint8_t sort_key(uint8_t m) { return (uint8_t)(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3)); }

}  // namespace

namespace vbsky {
// mask -> rank used on the device: the four unambiguous states come first so that a rank < 4
// is directly the column of P(t) to read; rank 4 is the fully ambiguous 'N'.
__host__ __device__ inline int mask_to_rank(int m) {
  switch (m & 15) {
    case 1: return 0;
    case 2: return 1;
    case 4: return 2;
    case 8: return 3;
    case 15: return 4;
    case 3: return 5;
    case 5: return 6;
    case 6: return 7;
    case 7: return 8;
    case 9: return 9;
    case 10: return 10;
    case 11: return 11;
    case 12: return 12;
    case 13: return 13;
    case 14: return 14;
    default: return 15;  // empty mask
  }
}

// [S][n] masks -> [n][S_pad] ranks (32x32 shared-memory transpose); rows >= S become 'N'.
__global__ void pack_tips_kernel(const uint8_t* __restrict__ patterns, uint8_t* __restrict__ codes,
                                 int n, long long S, long long S_pad) {
  __shared__ uint8_t tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const long long t = blockIdx.z;
  const uint8_t* src = patterns + t * S * n;
  uint8_t* dst = codes + t * (long long)n * S_pad;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long s = s0 + r;
    const int c = c0 + threadIdx.x;
    uint8_t v = 15;
    if (s < S && c < n) v = src[s * n + c];
    tile[r][threadIdx.x] = (uint8_t)mask_to_rank(v);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long s = s0 + threadIdx.x;
    if (c < n && s < S_pad) dst[(long long)c * S_pad + s] = tile[threadIdx.x][r];
  }
}

// tip-major ranks [n][S_pad] -> schedule-ordered, tile-major rows [tiles][2(n-1)][kTile]:
// row 2s / 2s+1 holds the codes of child a / b of step s when that child is a tip.
__global__ void order_tips_kernel(const uint8_t* __restrict__ tipmajor, const Op* __restrict__ ops, int n,
                                  long long S_pad, uint8_t* __restrict__ out) {
  const int rows = 2 * (n - 1);
  const int tile = blockIdx.x;
  for (int row = blockIdx.y; row < rows; row += gridDim.y) {
    const Op op = ops[row >> 1];
    const int node = (row & 1) ? ((op.x >> 16) & 0xffff) : (op.x & 0xffff);
    uint8_t v = 4;  // 'N' filler for internal children (never read)
    if (node < n) v = tipmajor[(long long)node * S_pad + (long long)tile * kTile + threadIdx.x];
    out[((long long)tile * rows + row) * kTile + threadIdx.x] = v;
  }
}
}  // namespace vbsky

extern "C" int vbsky_encode_tips(const char* seqs, int64_t n, int64_t L, uint8_t* codes) {
  VB_CHECK_ARG(seqs && codes, "NULL argument");
  VB_CHECK_ARG(n >= 1 && L >= 0, "bad shape n=%lld L=%lld", (long long)n, (long long)L);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t l = 0; l < L; ++l) {
      const unsigned char ch = (unsigned char)seqs[i * L + l];
      const uint8_t m = kEnc.t[ch];
      if (m == 0xff) return set_error(VBSKY_EKEY, "KeyError: character '%c' (0x%02x) at sequence %lld, site %lld is not in the BEAST nucleotide table", ch >= 32 && ch < 127 ? ch : '?', ch, (long long)i, (long long)l);
      codes[l * n + i] = m;
    }
  return VBSKY_OK;
}

extern "C" int vbsky_compress_patterns(const uint8_t* codes, int64_t L, int64_t n, const int64_t* perm,
                                       uint8_t* patterns, int64_t* counts, int64_t* S_out) {
  VB_CHECK_ARG(codes && patterns && counts && S_out, "NULL argument");
  VB_CHECK_ARG(L >= 0 && n >= 1, "bad shape L=%lld n=%lld", (long long)L, (long long)n);
  if (perm) {
    std::vector<char> seen(n, 0);
    for (int64_t j = 0; j < n; ++j) {
      VB_CHECK_ARG(perm[j] >= 0 && perm[j] < n && !seen[perm[j]], "perm is not a permutation of 0..n-1");
      seen[perm[j]] = 1;
    }
  }
  // keys[l][j] = sort key of column perm[j]
  std::vector<uint8_t> keys((size_t)L * n);
  for (int64_t l = 0; l < L; ++l)
    for (int64_t j = 0; j < n; ++j) {
      const uint8_t m = codes[l * n + (perm ? perm[j] : j)];
      VB_CHECK_ARG(m <= 15, "code %d at site %lld is not a 4-bit mask", (int)m, (long long)l);
      keys[l * n + j] = sort_key(m);
    }
  std::vector<int64_t> order(L);
  std::iota(order.begin(), order.end(), 0);
  const uint8_t* kp = keys.data();
  std::sort(order.begin(), order.end(), [kp, n](int64_t a, int64_t b) {
    return std::memcmp(kp + a * n, kp + b * n, (size_t)n) < 0;
  });
  int64_t S = 0;
  for (int64_t i = 0; i < L; ++i) {
    const uint8_t* row = kp + order[i] * n;
    if (i > 0 && std::memcmp(row, kp + order[i - 1] * n, (size_t)n) == 0) {
      counts[S - 1]++;
    } else {
      for (int64_t j = 0; j < n; ++j) patterns[S * n + j] = sort_key(row[j]);  // key -> mask (involution)
      counts[S] = 1;
      ++S;
    }
  }
  *S_out = S;
  return VBSKY_OK;
}

extern "C" void vbsky_tips_destroy(vbsky_tips* tips) {
  if (!tips) return;
  cudaFree(tips->d_codes_up);
  cudaFree(tips->d_codes_dn);
  cudaFree(tips->d_counts);
  delete tips;
}

extern "C" int vbsky_tips_create(const vbsky_layout* lay, int64_t S, const uint8_t* patterns,
                                 const double* counts, vbsky_tips** out) {
  VB_CHECK_ARG(out, "out is NULL");
  *out = nullptr;
  VB_CHECK_ARG(lay, "layout is NULL");
  VB_CHECK_DEVICE(lay);
  const int32_t T = lay->T, n = lay->n;
  VB_CHECK_ARG(S >= 1, "S must be >= 1 (got %lld)", (long long)S);
  VB_CHECK_ARG(patterns && counts, "NULL argument");
  const int64_t total = (int64_t)T * S * n;
  for (int64_t i = 0; i < total; ++i) VB_CHECK_ARG(patterns[i] <= 15, "pattern code %d is not a 4-bit mask", (int)patterns[i]);

  auto* tp = new vbsky_tips();
  tp->T = T, tp->n = n, tp->S = S;
  tp->S_pad = (S + kTile - 1) / kTile * kTile;
  tp->tiles = (int32_t)(tp->S_pad / kTile);
  const size_t rows = 2 * (size_t)(n - 1);
  const size_t per_tree = rows * tp->S_pad;
  uint8_t *d_raw = nullptr, *d_tipmajor = nullptr;
  auto fail = [&](int code) {
    cudaFree(d_raw);
    cudaFree(d_tipmajor);
    vbsky_tips_destroy(tp);
    return code;
  };
#define VB_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return fail(set_error(VBSKY_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e))); \
  } while (0)
  VB_TRY(cudaGetDevice(&tp->device));
  VB_TRY(cudaMalloc((void**)&tp->d_codes_up, (size_t)T * per_tree));
  VB_TRY(cudaMalloc((void**)&tp->d_codes_dn, (size_t)T * per_tree));
  VB_TRY(cudaMalloc((void**)&tp->d_counts, (size_t)T * tp->S_pad * sizeof(double)));
  VB_TRY(cudaMemset(tp->d_counts, 0, (size_t)T * tp->S_pad * sizeof(double)));
  VB_TRY(cudaMemcpy2D(tp->d_counts, tp->S_pad * sizeof(double), counts, S * sizeof(double), S * sizeof(double), T, cudaMemcpyHostToDevice));
  // Upload the masks tree by tree (bounded staging buffers); transpose/remap, then gather rows in schedule order.
  VB_TRY(cudaMalloc((void**)&d_raw, (size_t)S * n));
  VB_TRY(cudaMalloc((void**)&d_tipmajor, (size_t)n * tp->S_pad));
  for (int t = 0; t < T; ++t) {
    VB_TRY(cudaMemcpy(d_raw, patterns + (size_t)t * S * n, (size_t)S * n, cudaMemcpyHostToDevice));
    dim3 grid((unsigned)((tp->S_pad + 31) / 32), (unsigned)((n + 31) / 32), 1), block(32, 8);
    pack_tips_kernel<<<grid, block>>>(d_raw, d_tipmajor, n, S, tp->S_pad);
    VB_TRY(cudaGetLastError());
    dim3 g2((unsigned)tp->tiles, (unsigned)std::min<size_t>(rows, 64));
    order_tips_kernel<<<g2, kTile>>>(d_tipmajor, lay->d_up + (size_t)t * (n - 1), n, tp->S_pad, tp->d_codes_up + (size_t)t * per_tree);
    VB_TRY(cudaGetLastError());
    order_tips_kernel<<<g2, kTile>>>(d_tipmajor, lay->d_down + (size_t)t * (n - 1), n, tp->S_pad, tp->d_codes_dn + (size_t)t * per_tree);
    VB_TRY(cudaGetLastError());
  }
  VB_TRY(cudaDeviceSynchronize());
#undef VB_TRY
  cudaFree(d_raw);
  cudaFree(d_tipmajor);
  *out = tp;
  return VBSKY_OK;
}

extern "C" int vbsky_tips_set_counts(vbsky_tips* tips, const double* counts) {
  VB_CHECK_ARG(tips && counts, "NULL argument");
  VB_CUDA(cudaMemcpy2D(tips->d_counts, tips->S_pad * sizeof(double), counts, tips->S * sizeof(double), tips->S * sizeof(double),
                       tips->T, cudaMemcpyHostToDevice));
  return VBSKY_OK;
}
