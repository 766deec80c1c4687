// Host-side staging copy of the drop-in callers (06_lstm_ode_integration.py:346 copies each pageable numpy batch to the device
// synchronously).  Here the pageable -> pinned leg is the bound of that call (the GPU and the PCIe link both wait on it), so it is
// native and does as little memory traffic as the destination format allows:
//   * several host threads, each a contiguous slice;
//   * non-temporal (streaming) stores: the pinned destination is only read by the DMA engine afterwards, so it is written without
//     the read-for-ownership a cached store would add (memcpy: read 4 + RFO 4 + write 4 bytes per value; here read 4 + write 4 or 2);
//   * optional fp32 -> bf16 narrowing on the way (round to nearest even on the bit pattern -- exactly what the input projection's
//     converter warps do on load with cvt.rn.bf16.f32, denormals included, so the bf16 engine's result bits do not change while
//     the PCIe bytes halve).
// No device code in this file: it is compiled by nvcc only so that the library stays a single link.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include "common.cuh"

namespace {

inline uint16_t bf16_rne(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fffu;      // NaN: cvt.rn.bf16.f32's canonical NaN
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

void narrow_scalar(uint16_t* dst, const float* src, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, src + i, 4);
    dst[i] = bf16_rne(u);
  }
}

__attribute__((target("avx2"))) inline __m256i bf16_rne8(__m256i u) {   // eight fp32 bit patterns -> eight 32-bit lanes holding bf16
  const __m256i lsb = _mm256_and_si256(_mm256_srli_epi32(u, 16), _mm256_set1_epi32(1));
  const __m256i r = _mm256_srli_epi32(_mm256_add_epi32(u, _mm256_add_epi32(lsb, _mm256_set1_epi32(0x7fff))), 16);
  const __m256i mag = _mm256_and_si256(u, _mm256_set1_epi32(0x7fffffff));
  const __m256i nan = _mm256_cmpgt_epi32(mag, _mm256_set1_epi32(0x7f800000));
  return _mm256_blendv_epi8(r, _mm256_set1_epi32(0x7fff), nan);
}

__attribute__((target("avx2"))) void narrow_avx2(uint16_t* dst, const float* src, int64_t n) {
  int64_t i = 0;
  const int64_t head = (int64_t)(((32 - ((uintptr_t)dst & 31)) & 31) / 2);   // bf16 values up to a 32-byte boundary of dst
  const int64_t h = head < n ? head : n;
  narrow_scalar(dst, src, h);
  i = h;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = bf16_rne8(_mm256_loadu_si256((const __m256i*)(src + i)));
    const __m256i b = bf16_rne8(_mm256_loadu_si256((const __m256i*)(src + i + 8)));
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xd8);   // packus interleaves the 128-bit halves
    _mm256_stream_si256((__m256i*)(dst + i), p);
  }
  narrow_scalar(dst + i, src + i, n - i);
  _mm_sfence();
}

__attribute__((target("avx2"))) void copy_avx2(float* dst, const float* src, int64_t n) {
  int64_t i = 0;
  const int64_t head = (int64_t)(((32 - ((uintptr_t)dst & 31)) & 31) / 4);
  const int64_t h = head < n ? head : n;
  memcpy(dst, src, (size_t)h * 4);
  i = h;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    _mm256_stream_si256((__m256i*)(dst + i), a);
    _mm256_stream_si256((__m256i*)(dst + i + 8), b);
  }
  memcpy(dst + i, src + i, (size_t)(n - i) * 4);
  _mm_sfence();
}

void stage_slice(void* dst, const float* src, int64_t n, int to_bf16, bool avx2) {
  if (to_bf16) {
    if (avx2) narrow_avx2((uint16_t*)dst, src, n);
    else narrow_scalar((uint16_t*)dst, src, n);
  } else {
    if (avx2) copy_avx2((float*)dst, src, n);
    else memcpy(dst, src, (size_t)n * 4);
  }
}

}  // namespace

extern "C" int bci_host_stage(void* dst, const float* src, int64_t n, int32_t to_bf16, int32_t threads) {
  if (n < 0 || (n > 0 && (!dst || !src)) || (to_bf16 ? ((uintptr_t)dst & 1) : ((uintptr_t)dst & 3))) {
    ::bci::set_error("bci_host_stage: null / misaligned buffer or negative count");
    return BCI_EINVAL;
  }
  static const bool avx2 = __builtin_cpu_supports("avx2");
  int nt = threads < 1 ? 1 : threads;
  const int64_t min_per_thread = 1 << 18;                   // below 1 MB of input a second thread costs more than it moves
  if ((int64_t)nt > (n + min_per_thread - 1) / min_per_thread) nt = (int)((n + min_per_thread - 1) / min_per_thread);
  if (nt <= 1) {
    stage_slice(dst, src, n, to_bf16, avx2);
    return BCI_OK;
  }
  // slices start on multiples of 16 values: every thread's destination keeps the alignment of dst
  const int64_t per = (((n + nt - 1) / nt) + 15) & ~(int64_t)15;
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  const size_t esz = to_bf16 ? 2 : 4;
  for (int t = 1; t < nt; ++t) {
    const int64_t a = per * t;
    if (a >= n) break;
    const int64_t m = (a + per <= n) ? per : n - a;
    pool.emplace_back(stage_slice, (char*)dst + (size_t)a * esz, src + a, m, (int)to_bf16, avx2);
  }
  stage_slice(dst, src, per < n ? per : n, to_bf16, avx2);
  for (auto& th : pool) th.join();
  return BCI_OK;
}
