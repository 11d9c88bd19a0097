// Tip encoding and site-pattern compression on the device, bit-exact with the
// host path (vbsky_encode_tips / vbsky_compress_patterns) and hence with the reference's
// encode_partials (substitution.py:38-48) + np.unique(axis=0, return_counts=True) (fasta.py:539-548).
//
// One CTA per alignment.  The rows (site columns of the alignment, n tips each) are sorted
// lexicographically with a least-significant-column-first radix sort on the 4-bit per-tip keys (stable
// counting sort per column, 16 buckets), then equal neighbours are merged.  NumPy orders the unpacked
// [n,4] 0/1 rows with state A most significant, so the key of a tip is its bit-reversed mask.
#include "common.h"

namespace vbsky {

constexpr int ST = 1024;  // threads per alignment

__device__ __forceinline__ int key_of(uint8_t m) { return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3); }

// Exclusive scan of flat[0..16*ST) in place; every thread owns 16 consecutive entries.
__device__ __forceinline__ void scan16(int* flat, int* warp_sums) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int loc[16], sum = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) loc[k] = flat[tid * 16 + k], sum += loc[k];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int v = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += x;
    }
    warp_sums[lane] = v;
  }
  __syncthreads();
  int run = incl - sum + (w > 0 ? warp_sums[w - 1] : 0);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int c = loc[k];
    flat[tid * 16 + k] = run;
    run += c;
  }
  __syncthreads();
}

struct CompressParams {
  const uint8_t* codes;  // [T][L][n] masks
  const int32_t* perm;   // [T][n] column permutation or null
  int32_t* order_a;      // [T][L]
  int32_t* order_b;      // [T][L]
  uint8_t* patterns;     // [T][L][n] out (first S rows valid)
  long long* counts;     // [T][L] out
  int32_t* S_out;        // [T]
  int L, n;
};

__global__ void __launch_bounds__(ST) compress_kernel(const CompressParams p) {
  extern __shared__ int sm[];
  int* flat = sm;  // [16][ST] bucket-major counters / offsets
  __shared__ int warp_sums[32];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int L = p.L, n = p.n;
  const uint8_t* codes = p.codes + (size_t)t * L * n;
  const int32_t* perm = p.perm ? p.perm + (size_t)t * n : nullptr;
  int32_t* oa = p.order_a + (size_t)t * L;
  int32_t* ob = p.order_b + (size_t)t * L;
  const int per = (L + ST - 1) / ST;
  const int i0 = min(tid * per, L), i1 = min(i0 + per, L);
  for (int i = tid; i < L; i += ST) oa[i] = i;
  __syncthreads();
  for (int col = n - 1; col >= 0; --col) {
    const int src = perm ? perm[col] : col;
    int c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = 0;
    for (int i = i0; i < i1; ++i) {
      const int k = key_of(codes[(size_t)oa[i] * n + src]);
#pragma unroll
      for (int b = 0; b < 16; ++b) c[b] += (k == b);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) flat[k * ST + tid] = c[k];
    __syncthreads();
    scan16(flat, warp_sums);
#pragma unroll
    for (int k = 0; k < 16; ++k) c[k] = flat[k * ST + tid];
    for (int i = i0; i < i1; ++i) {
      const int row = oa[i];
      const int k = key_of(codes[(size_t)row * n + src]);
      int pos = 0;
#pragma unroll
      for (int b = 0; b < 16; ++b)
        if (k == b) pos = c[b]++;
      ob[pos] = row;
    }
    __syncthreads();
    int32_t* tmp = oa;
    oa = ob, ob = tmp;
  }
  // oa holds the sorted row order.  Mark group starts, scan, emit patterns and counts.
  int nstart = 0;
  for (int i = i0; i < i1; ++i) {
    bool start = (i == 0);
    if (!start) {
      const uint8_t* a = codes + (size_t)oa[i] * n;
      const uint8_t* b = codes + (size_t)oa[i - 1] * n;
      for (int j = 0; j < n && !start; ++j) start = a[j] != b[j];  // equality does not depend on the column order
    }
    ob[i] = start ? 1 : 0;
    nstart += start;
  }
  // exclusive scan of the per-thread start counts (reuse flat: one entry per thread, rest zero)
#pragma unroll
  for (int k = 0; k < 16; ++k) flat[tid * 16 + k] = (k == 0) ? nstart : 0;
  __syncthreads();
  scan16(flat, warp_sums);
  int g = flat[tid * 16];  // groups started before this thread's chunk
  uint8_t* pout = p.patterns + (size_t)t * L * n;
  long long* cout = p.counts + (size_t)t * L;
  for (int i = i0; i < i1; ++i) {
    if (ob[i]) {
      const uint8_t* a = codes + (size_t)oa[i] * n;
      for (int j = 0; j < n; ++j) pout[(size_t)g * n + j] = a[perm ? perm[j] : j];
      cout[g] = i;  // start position; turned into a count below
      ++g;
    }
  }
  if (i1 == L && i0 < L) p.S_out[t] = g;
  if (L == 0 && tid == 0) p.S_out[t] = 0;
  __syncthreads();
  __threadfence_block();
  const int S = p.S_out[t];
  // counts[g] = start[g+1] - start[g]; read all starts first, then overwrite
  for (int base = 0; base < S; base += ST) {
    const int gi = base + tid;
    long long a = 0, b = 0;
    if (gi < S) a = cout[gi], b = (gi + 1 < S) ? cout[gi + 1] : (long long)L;
    __syncthreads();
    if (gi < S) cout[gi] = b - a;
    __syncthreads();
  }
}

__global__ void encode_kernel(const char* __restrict__ seqs, const uint8_t* __restrict__ table, long long n, long long L,
                              uint8_t* __restrict__ codes, int* __restrict__ bad) {
  // seqs [n][L] characters -> codes [L][n] masks (32x32 tile transpose through shared memory)
  __shared__ uint8_t tile[32][33];
  const long long l0 = (long long)blockIdx.x * 32, i0 = (long long)blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long i = i0 + r, l = l0 + threadIdx.x;
    uint8_t m = 15;
    if (i < n && l < L) {
      m = table[(unsigned char)seqs[i * L + l]];
      if (m == 0xff) atomicMin(bad, (int)min((long long)0x7fffffff, i * L + l)), m = 15;
    }
    tile[r][threadIdx.x] = m;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long l = l0 + r, i = i0 + threadIdx.x;
    if (l < L && i < n) codes[l * n + i] = tile[threadIdx.x][r];
  }
}

}  // namespace vbsky

using namespace vbsky;

extern "C" size_t vbsky_compress_workspace_bytes(int32_t T, int64_t L) { return (size_t)T * (size_t)L * 2 * sizeof(int32_t) + 256; }

extern "C" int vbsky_compress_patterns_device(int32_t T, int64_t L, int32_t n, const uint8_t* codes, const int32_t* perm,
                                              uint8_t* patterns, int64_t* counts, int32_t* S_out, void* ws, size_t ws_bytes,
                                              void* stream) {
  VB_CHECK_ARG(codes && patterns && counts && S_out && ws, "NULL argument");
  VB_CHECK_ARG(T >= 1 && L >= 1 && L < (1LL << 31) && n >= 1, "bad shape T=%d L=%lld n=%d", T, (long long)L, n);
  if (ws_bytes < vbsky_compress_workspace_bytes(T, L)) return set_error(VBSKY_ENOMEM, "compress workspace too small");
  CompressParams p;
  p.codes = codes, p.perm = perm;
  p.order_a = (int32_t*)ws, p.order_b = p.order_a + (size_t)T * L;
  p.patterns = patterns, p.counts = (long long*)counts, p.S_out = S_out, p.L = (int)L, p.n = n;
  const size_t smem = (size_t)16 * ST * sizeof(int);
  VB_CUDA(cudaFuncSetAttribute(compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  compress_kernel<<<T, ST, smem, (cudaStream_t)stream>>>(p);
  VB_CUDA(cudaGetLastError());
  return VBSKY_OK;
}

extern "C" int vbsky_encode_tips_device(const char* seqs, int64_t n, int64_t L, uint8_t* codes, int64_t* bad_index, void* stream) {
  VB_CHECK_ARG(seqs && codes, "NULL argument");
  VB_CHECK_ARG(n >= 1 && L >= 1 && n * L < (1LL << 31), "bad shape n=%lld L=%lld", (long long)n, (long long)L);
  // BEAST nucleotide table (substitution.py:10-30) as masks; 0xff = KeyError in the reference
  uint8_t t[256];
  for (int i = 0; i < 256; ++i) t[i] = 0xff;
  t['A'] = 1, t['C'] = 2, t['G'] = 4, t['T'] = 8, t['U'] = 8;
  t['R'] = 5, t['Y'] = 10, t['M'] = 3, t['W'] = 9, t['S'] = 6, t['K'] = 12;
  t['B'] = 14, t['D'] = 13, t['H'] = 11, t['V'] = 7;
  t['N'] = 15, t['X'] = 15, t['-'] = 15, t['?'] = 15;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* d_table = nullptr;
  int* d_bad = nullptr;
  VB_CUDA(cudaMalloc((void**)&d_table, 256));
  VB_CUDA(cudaMalloc((void**)&d_bad, sizeof(int)));
  int init = 0x7fffffff, bad = 0;
  VB_CUDA(cudaMemcpyAsync(d_table, t, 256, cudaMemcpyHostToDevice, st));
  VB_CUDA(cudaMemcpyAsync(d_bad, &init, sizeof(int), cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((L + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
  encode_kernel<<<grid, block, 0, st>>>(seqs, d_table, n, L, codes, d_bad);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_table), cudaFree(d_bad);
  if (e != cudaSuccess) return set_error(VBSKY_ECUDA, "encode failed: %s", cudaGetErrorString(e));
  if (bad_index) *bad_index = bad == 0x7fffffff ? -1 : bad;
  if (bad != 0x7fffffff)
    return set_error(VBSKY_EKEY, "KeyError: character at sequence %lld, site %lld is not in the BEAST nucleotide table",
                     (long long)(bad / L), (long long)(bad % L));
  return VBSKY_OK;
}
