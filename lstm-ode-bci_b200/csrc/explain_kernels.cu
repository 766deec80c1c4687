// Permutation-importance inputs built on the device (SURVEY.md §8 f rank 1; 07_explainability.py:287-361).
//
// The reference copies the whole test subset on the host for every (channel, repetition) pair, overwrites one channel with the
// same channel of a permuted sample order (07:336-339), uploads the copy in batches of 128 and runs the model: 61 x 5 + 1 sweeps of
// 1 000 windows.  Here the subset stays resident in HBM and every variant row is materialised by one gather straight into the
// forward's input layout -- V variants share one launch, so the forward runs at its large-batch rate instead of at batch 128.
//
// HBM-bound byte work: algorithmic traffic = the rows written (16 B per four fp32 elements, 8 B in bf16) + the subset read once;
// the roofline is the write stream.  Grid: a multiple of the SM count (8 CTAs of 256 threads per SM), tasks grid-strided.
#include "common.cuh"
#include <cuda_bf16.h>

namespace bci {

// out row r (global variant row g = row0 + r): variant v = g / n, sample i = g % n;
//   out[r][t][c] = x[c == channel[v] ? perm[g] : i][t][c]        (channel[v] < 0: an unpermuted copy, the baseline sweep)
//
// Vector kernel (wlen = T*C a multiple of 4, C >= 4 so a 4-vector holds the permuted channel at most once, 16-byte aligned
// pointers).  A task is (sample i, group of 1 024 float4 of its window): the CTA loads that piece ONCE -- four independent 16-byte
// loads per thread -- and writes it to every variant of the row range, patching the one element in C that comes from the permuted
// sample.  History (ncu, profiles/r5_permute_gather.md): row-per-CTA with one load in flight per thread ran at 38 % of DRAM peak;
// four loads in flight 46 %, but every variant re-read its sample from DRAM (L2 hit rate 27 %: the streamed output evicts the
// subset); reading each piece once per task leaves the write stream as the only HBM traffic.
template <bool BF16_OUT>
__global__ void __launch_bounds__(256, 4) permute_channels_vec_kernel(const float* __restrict__ x, int n, int wlen, int C,
                                                                   const int32_t* __restrict__ perm,
                                                                   const int32_t* __restrict__ channel, long long row0,
                                                                   long long rows, void* __restrict__ out) {
  __shared__ int s_perm[256], s_ch[256];
  const int wq = wlen >> 2;
  const int groups = (wq + 1023) >> 10;
  const long long tasks = (long long)n * groups;
  for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int i = (int)(task / groups), q_base = (int)(task - (long long)i * groups) << 10;
    // variants v with row0 <= v*n + i < row0 + rows
    const long long lo = row0 - i, hi = row0 + rows - 1 - i;
    if (hi < 0) continue;
    const long long v_lo = lo <= 0 ? 0 : (lo + n - 1) / n, v_hi = hi / n;
    if (v_hi < v_lo) continue;
    const float* base = x + (size_t)i * wlen;
    float4 val[4];
    int c0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = q_base + k * 256 + threadIdx.x;
      if (q < wq) val[k] = __ldg(reinterpret_cast<const float4*>(base) + q);
      c0[k] = (4 * q) % C;
    }
    for (long long vb = v_lo; vb <= v_hi; vb += 256) {
      const int nv = (int)((v_hi - vb + 1) < 256 ? (v_hi - vb + 1) : 256);
      __syncthreads();
      if ((int)threadIdx.x < nv) {
        s_ch[threadIdx.x] = __ldg(channel + vb + threadIdx.x);
        s_perm[threadIdx.x] = __ldg(perm + (vb + threadIdx.x) * n + i);
      }
      __syncthreads();
      for (int vv = 0; vv < nv; ++vv) {
        const int ch = s_ch[vv];
        const float* other = x + (size_t)s_perm[vv] * wlen;
        const size_t r = (size_t)((vb + vv) * n + i - row0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = q_base + k * 256 + threadIdx.x;
          if (q >= wq) continue;
          float4 o = val[k];
          int j = ch - c0[k];
          if (j < 0) j += C;
          if (ch >= 0 && j < 4) {
            const float f = __ldg(other + 4 * q + j);
            if (j == 0) o.x = f; else if (j == 1) o.y = f; else if (j == 2) o.z = f; else o.w = f;
          }
          if (BF16_OUT) {
            __nv_bfloat162 l2 = __floats2bfloat162_rn(o.x, o.y), h2 = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&l2);
            pk.y = *reinterpret_cast<uint32_t*>(&h2);
            __stcs(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * wlen) + q, pk);
          } else {
            __stcs(reinterpret_cast<float4*>(static_cast<float*>(out) + r * wlen) + q, o);
          }
        }
      }
    }
  }
}

// Scalar kernel for every other shape: one CTA per output row.
template <bool BF16_OUT>
__global__ void __launch_bounds__(256) permute_channels_kernel(const float* __restrict__ x, int n, int wlen, int C,
                                                               const int32_t* __restrict__ perm,
                                                               const int32_t* __restrict__ channel, long long row0,
                                                               long long rows, void* __restrict__ out) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const long long g = row0 + r;
    const int v = (int)(g / n), i = (int)(g - (long long)v * n);
    const int ch = __ldg(channel + v), p = __ldg(perm + g);
    const float* base = x + (size_t)i * wlen;
    const float* other = x + (size_t)p * wlen;
    for (int e = threadIdx.x; e < wlen; e += blockDim.x) {
      const float val = (e % C == ch) ? __ldg(other + e) : __ldg(base + e);
      if (BF16_OUT) static_cast<__nv_bfloat16*>(out)[(size_t)r * wlen + e] = __float2bfloat16_rn(val);
      else static_cast<float*>(out)[(size_t)r * wlen + e] = val;
    }
  }
}

}  // namespace bci

extern "C" int bci_permute_channels(const float* x, int32_t n, int32_t seq_len, int32_t channels, const int32_t* perm,
                                    const int32_t* channel, int64_t row0, int64_t rows, int32_t out_dtype, void* out,
                                    void* stream) {
  using namespace bci;
  BCI_REQUIRE(n >= 1 && seq_len >= 1 && channels >= 1 && row0 >= 0 && rows >= 0, BCI_EINVAL,
              "bci_permute_channels: n, seq_len, channels must be >= 1 and row0, rows >= 0");
  BCI_REQUIRE(out_dtype == BCI_IN_F32 || out_dtype == BCI_IN_BF16, BCI_EINVAL, "bci_permute_channels: out_dtype must be BCI_IN_F32 or BCI_IN_BF16");
  BCI_REQUIRE((long long)seq_len * channels < (1ll << 31), BCI_EINVAL, "bci_permute_channels: window too large");
  if (rows == 0) return BCI_OK;
  BCI_REQUIRE(x && perm && channel && out, BCI_EINVAL, "bci_permute_channels: NULL pointer");
  NvtxRange nv("bci_permute_channels");
  const int wlen = seq_len * channels;
  const bool vec = (wlen % 4 == 0) && channels >= 4 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const long long cap = 8ll * sm_count();  // eight 256-thread CTAs per SM: every SM holds its full thread complement
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = out_dtype == BCI_IN_BF16;
  if (vec) {
    const long long tasks = (long long)n * (((wlen >> 2) + 1023) >> 10);
    const long long cap4 = 4ll * sm_count();  // 50 registers: four resident CTAs per SM
    const unsigned grid = (unsigned)(tasks < cap4 ? tasks : cap4);
    if (bf) permute_channels_vec_kernel<true><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
    else permute_channels_vec_kernel<false><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
  } else {
    const unsigned grid = (unsigned)(rows < cap ? rows : cap);
    if (bf) permute_channels_kernel<true><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
    else permute_channels_kernel<false><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}
