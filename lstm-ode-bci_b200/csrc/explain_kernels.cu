// Permutation-importance inputs built on the device (SURVEY.md §8 f rank 1; 07_explainability.py:287-361).
//
// The reference copies the whole test subset on the host for every (channel, repetition) pair, overwrites one channel with the
// same channel of a permuted sample order (07:336-339), uploads the copy in batches of 128 and runs the model: 61 x 5 + 1 sweeps of
// 1 000 windows.  Here the subset stays resident in HBM and every variant row is materialised by one gather straight into the
// forward's input layout -- V variants share one launch, so the forward runs at its large-batch rate instead of at batch 128.
//
// HBM-bound byte work: one 16-byte load of the unpermuted window + (once every C elements) a 4-byte load of the permuted sample's
// channel value per 16 (fp32) or 8 (bf16) bytes written.  The subset (n x T x C fp32 = 62 MB at n = 1 000) is L2-resident, so the
// roofline is the write stream.  Grid: a multiple of the SM count, rows grid-strided.
#include "common.cuh"
#include <cuda_bf16.h>

namespace bci {

// out row r (global variant row g = row0 + r): variant v = g / n, sample i = g % n;
//   out[r][t][c] = x[c == channel[v] ? perm[g] : i][t][c]        (channel[v] < 0: an unpermuted copy, the baseline sweep)
// VEC: wlen (= T*C) is a multiple of 4 and C >= 4, so a 4-vector holds the permuted channel at most once.
template <bool BF16_OUT, bool VEC>
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
    if (VEC) {
      const int wq = wlen >> 2;
      for (int q = threadIdx.x; q < wq; q += blockDim.x) {
        float4 val = __ldg(reinterpret_cast<const float4*>(base) + q);
        int j = ch - (4 * q) % C;
        if (j < 0) j += C;
        if (ch >= 0 && j < 4) {
          const float o = __ldg(other + 4 * q + j);
          if (j == 0) val.x = o; else if (j == 1) val.y = o; else if (j == 2) val.z = o; else val.w = o;
        }
        if (BF16_OUT) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(val.x, val.y), hi = __floats2bfloat162_rn(val.z, val.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          __stcs(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + (size_t)r * wlen) + q, pk);
        } else {
          __stcs(reinterpret_cast<float4*>(static_cast<float*>(out) + (size_t)r * wlen) + q, val);
        }
      }
    } else {
      for (int e = threadIdx.x; e < wlen; e += blockDim.x) {
        const float val = (e % C == ch) ? __ldg(other + e) : __ldg(base + e);
        if (BF16_OUT) static_cast<__nv_bfloat16*>(out)[(size_t)r * wlen + e] = __float2bfloat16_rn(val);
        else static_cast<float*>(out)[(size_t)r * wlen + e] = val;
      }
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
  const unsigned grid = (unsigned)(rows < cap ? rows : cap);
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = out_dtype == BCI_IN_BF16;
  if (vec) {
    if (bf) permute_channels_kernel<true, true><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
    else permute_channels_kernel<false, true><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
  } else {
    if (bf) permute_channels_kernel<true, false><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
    else permute_channels_kernel<false, false><<<grid, 256, 0, st>>>(x, n, wlen, channels, perm, channel, row0, rows, out);
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}
