// Shared helpers for the bci_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <nvtx3/nvToolsExt.h>

#include "../../include/bci_b200.h"

namespace bci {

void set_error(const char* fmt, ...);
void note_launch();  // counts kernel launches (bci_launch_count)

#define BCI_CUDA_OK(expr)                                                             \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::bci::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BCI_ECUDA;                                                               \
    }                                                                                 \
  } while (0)

#define BCI_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::bci::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define BCI_LAUNCH_OK()                                                                \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::bci::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BCI_ECUDA;                                                                \
    }                                                                                  \
    ::bci::note_launch();                                                              \
  } while (0)

// Per-device state: one process may drive several GPUs (the tests and torch allow it), and kernel attributes
// (cudaFuncSetAttribute), occupancy results and the SM count belong to the device that is current at the call.
// BCI_NVTX=1: NVTX ranges around the library's entry points (no-ops without an attached tool; nvtx3 is header-only)
inline bool nvtx_on() {
  static const bool on = [] { const char* e = getenv("BCI_NVTX"); return e && e[0] == '1'; }();
  return on;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_on()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

constexpr int BCI_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & (BCI_MAX_DEVICES - 1);
}
struct PerDeviceFlag {
  bool v[BCI_MAX_DEVICES] = {};
  bool& cur() { return v[current_device()]; }
};
struct PerDeviceInt {
  int v[BCI_MAX_DEVICES] = {};
  int& cur() { return v[current_device()]; }
};
inline int sm_count() {
  static PerDeviceInt n_pd;
  int& n = n_pd.cur();
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
    if (n <= 0) n = 148;
  }
  return n;
}

__host__ __device__ constexpr int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU: nn.GELU() default, 04_lstm_model.py:176
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Gate activations of the fp32 recurrences (lstm_rec_f32, lstm_bptt_f32): ex2.approx-based forms (MUFU.EX2 + one fast division, ~8 instructions) instead of libm's
// expf / tanhf (~25).  The five activations per (unit, window, step) were 20 % of the kernel's instructions at inference batch sizes:
// fp32 forward 91.5 k -> 99.7 k windows/s, training step 16.3 -> 15.3 ms.  Absolute error ~1e-7 per gate; measured against the
// reference's CPU path at logit gain 12: logits 7.2e-7 (H = 128) / 2.1e-6 (H = 256), attention <= 8e-9 -- inside the 1e-5 / 1e-6
// parity tolerances with 5-14x to spare (3e-7 / 1e-6 with libm).  -DBCI_REC_ACCURATE_ACT (BCI_NVCC_DEFINES, build.py) restores libm.
// fast_expf / fast_divf: the flush-to-zero forms `ex2.approx.ftz.f32` / `rcp.approx.ftz.f32` = ONE MUFU each.  CUDA's __expf / __fdividef
// are the non-ftz approximations, which ptxas wraps in denormal range handling -- FSETP + two predicated FMULs around every MUFU.EX2,
// FSETP + scaling FMULs around every MUFU.RCP: in the SASS of lstm_rec_f16x3_pipe that was 7 FSETP and ~14 of the 25 FMULs per hidden
// unit (64 instructions per unit in all).  Inside the fast path both forms are the same MUFU instruction, so results only differ where a
// value is below 2^-126 (flushed to 0) or a divisor above 2^126 (quotient flushed to 0): for the gate forms below that is sigma -> 0 / 1
// and tanh -> -1 / 1 at |x| > 87, their correctly rounded limits.
__device__ __forceinline__ float fast_ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_expf(float x) { return fast_ex2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_divf(float a, float b) { return a * fast_rcpf(b); }

#ifndef BCI_REC_ACCURATE_ACT
__device__ __forceinline__ float rec_sigmoid(float x) { return fast_rcpf(1.0f + fast_expf(-x)); }
__device__ __forceinline__ float rec_tanh(float x) { return 1.0f - 2.0f * fast_rcpf(1.0f + fast_expf(2.0f * x)); }
#else
__device__ __forceinline__ float rec_sigmoid(float x) { return sigmoid_acc(x); }
__device__ __forceinline__ float rec_tanh(float x) { return tanhf(x); }
#endif


// stateless dropout of the training step: the mask is a hash of (seed, site, element index) and is regenerated instead of stored
__device__ __forceinline__ float drop_scale(uint64_t seed, uint32_t site, uint64_t idx, float p) {
  if (p <= 0.f) return 1.f;
  // 32-bit avalanche hash (two multiply-xorshift rounds) of the element index keyed by (seed, site): ~10 integer instructions per
  // element -- the masks are evaluated inside the recurrence epilogues, where the 64-bit splitmix of the first version (six 64-bit
  // multiplies' worth of 32-bit IMADs) cost 50 us per layer.  The key part is loop-invariant.
  const uint32_t key = ((uint32_t)seed * 0x9E3779B1u) ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u) ^ (site * 0xC2B2AE3Du) ^ 0x27D4EB2Fu;
  uint32_t h = ((uint32_t)idx * 0x9E3779B1u) ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ key;
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
  return u < p ? 0.f : 1.0f / (1.0f - p);
}

}  // namespace bci
