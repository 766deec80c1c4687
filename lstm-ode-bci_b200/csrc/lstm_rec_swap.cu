// Training-size recurrences on the tensor cores: the "swapped" form  pre^T = W_hh . h^T  with W_hh RESIDENT IN TENSOR MEMORY.
//
// A 512-window training batch is 4 tiles of 128 windows: the tile-per-CTA tensor-core recurrences of the inference paths
// (lstm_bf16_fused.cu, lstm_fp32_tc.cu) would keep 8 SMs busy, which is why the training step of round 1 ran its recurrences on
// the CUDA cores (thread = (unit, 4 windows), 128 CTAs: 4.9 us per forward step, 7 us per BPTT step; 62 % of the step).  Swapping
// the operands makes the WINDOWS the MMA's N dimension, which may be as small as 16:
//
//   forward   D_g[u][n] = sum_k W_hh[g*128 + u][k] . h_{t-1}[n][k]      four M128 x N16 x K128 products (one per gate) -> 64 TMEM columns
//   BPTT      D[j][n]   = sum_k W_hh^T[j][k] . dG_t[n][k],  k = (gate, unit)    one M128 x N16 x K512 product            -> 16 TMEM columns
//
//   A = the weights: 128 KB of 16-bit values per direction = 256 of the SM's 512 tensor-memory columns (lane = row, column c of a
//       K = 16 slice = elements 2c | 2c+1 << 16), written once per CTA with tcgen05.st and read by tcgen05.mma as its A operand:
//       an M128 x N16 x K16 product then costs 8 tensor cycles, against 32 for a 4 KB A tile on the shared-memory port
//   B = h_{t-1} (4 KB) / dG_t (16 KB) in shared memory: rows = windows, written by the epilogue threads of the previous step
//   D = TMEM lane u (hidden unit) x column n (window): a thread owns unit u for two of the CTA's 8 windows and reads its four gate
//       pre-activations with four tcgen05.ld.32x32b.x2 -- the whole cell update is thread-local, no exchange of any kind.
//
// One CTA = 8 windows x one direction (N = 16 is the smallest legal N at M = 128; rows 8-15 of B stay zero), one CTA per SM:
// 512 windows x 2 directions = 128 CTAs.  16 epilogue warps + one MMA warp, coupled by two mbarriers only (acc_full: tcgen05.commit;
// op_ready: one arrive per epilogue warp).  Three things took the forward step from 2.44 to 0.93 us (clock64 timelines, CTA 0):
//   * the MMAs are issued by ONE ELECTED LANE OF A CONVERGED WARP whose index the compiler knows to be warp-uniform
//     (__shfl_sync(tid / 32) + elect.sync): ptxas then emits the 32 UTCHMMA back to back (260 cycles); issued under `if (tid == 0)`
//     every one of them sat in its own election loop (ELECT / BRA.U.ANY) and cost 47 cycles -- 1 500 per step, more than the math
//   * 16 epilogue warps instead of 4: with one warp per scheduler the per-window dependent chain (tcgen05.ld -> 5 activations ->
//     cell update -> stores) ran at its full latency, 2 400 cycles for 8 windows; four warps per scheduler overlap it: 650
//   * next step's G_t / saved activations requested into registers before the accumulator wait, and the rows of four steps
//     ahead pulled into L2 (a single step of lookahead is shorter than a DRAM round trip under load: step times jittered 1.5-2.5 k)
//
// Tried and rejected: two independent window groups per CTA (4 windows each, own B tile / accumulator columns / barrier pair) so that
// one group's product runs under the other group's epilogue.  Correct, but slower (forward 0.95 -> 1.03 us, BPTT 1.00 -> 1.22 us per
// step, the split forms 25 % slower): after the fixes above a step is no longer "MMA latency + epilogue" but the epilogue warps' own
// instruction stream (~170 instructions per thread and step on 16 warps: issue slots 59 % busy) -- there is no idle time left for a
// second group to fill, and the N = 16 minimum doubles the tensor work.
//
// Two precisions:
//   mixed (BCI_TRAIN_MIXED; the analogue of the reference's autocast training, 04_lstm_model.py:486-490): one product chain, forward
//       operands fp16 (h in [-1, 1], 11 significant bits), BPTT operands bf16 (gradients need the exponent range, not the bits),
//       gates from tanh.approx (one MUFU each); accumulation, gates, cell state, dG in fp32.
//   split (the fp32-parity step): every operand as an fp16 (hi, lo) pair and three chains  lo.hi + hi.lo + hi.hi  (small terms
//       first; the arithmetic of lstm_fp32_tc.cu): W_hi in tensor memory, W_lo (128 KB) in shared memory; BPTT scales each step's
//       dG tile by a power of two that follows the gradient's magnitude (see lstm_bptt_swap).  3e-7 from float64 on h, 2e-7 of
//       autograd on dG; forward 0.50 ms and BPTT 0.67 ms per layer against 1.25 / 1.8 ms on the CUDA cores.
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int SW_NW = 8;                     // windows per CTA
constexpr uint32_t SW_BATOM = 16 * 128;      // B atom: 16 rows x 64 halves (rows 8-15 zero)
constexpr uint32_t SW_FWD_B = 2 * SW_BATOM, SW_BWD_B = 8 * SW_BATOM;
constexpr uint32_t SW_AATOM = 128 * 128;     // A atom in shared memory (split mode: the weights' lo parts): 128 rows x 64 halves
constexpr uint32_t SW_ALO_BYTES = 2 * SW_AATOM;   // split mode: the fourth quarter of W_lo (forward: gate o, BPTT: K 384-511), two atoms
constexpr float SW_WSCALE = 16.0f;           // weights are stored x 16 (keeps the lo parts out of fp16's subnormals)
// SPLIT (fp32-parity step): weights hi in tensor memory, weights lo in shared memory, B tile as an fp16 (hi, lo) pair
constexpr size_t sw_smem_bytes(uint32_t b_bytes, bool split) { return 1024 + (split ? SW_ALO_BYTES + 2 * b_bytes : b_bytes) + 64; }

__host__ __device__ constexpr uint32_t sw_idesc(int M, int N, bool bf16) {
  return (1u << 4) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns, registers -> tensor memory
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[TENSOR MEMORY] * B[smem]^T: A is 128 lanes (rows) x 8 columns per K = 16 slice, column c of a slice = the 16-bit
// K elements (2c | 2c+1 << 16) (tests/test_gpu_rec_swap.py::test_tmem_a_operand_layout_probe).  Read at TMEM bandwidth: an
// M128 x N16 x K16 product costs ~8 tensor cycles instead of the 32 its 4 KB A tile needs on the shared-memory port.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
constexpr uint32_t SW_WCOL = 64;   // first TMEM column of the resident weights (accumulators sit in columns [0, 64))

// ---- operand packing --------------------------------------------------------------------------------------------------------
// w_hh (4H, H) fp32 -> [part hi/lo][4H][H] fp16 of 16 w, same row order
__global__ void pack_whh_swap_fwd_kernel(const float* __restrict__ w, __half* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i] * SW_WSCALE;
  const __half hi = __float2half_rn(v);
  dst[i] = hi;
  dst[n + i] = __float2half_rn(v - __half2float(hi));
}
// w_hh (4H, H) fp32 -> the transpose [j][k = gate*H + unit] = w_hh[k][j]: bf16 (mixed BPTT) and an fp16 (hi, lo) pair of 16 w
__global__ void pack_whh_swap_bwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, __half* __restrict__ dst16, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H * H) return;
  const int j = i / (4 * H), k = i - j * 4 * H;
  const float v = w[(size_t)k * H + j];
  dst[i] = __float2bfloat16_rn(v);
  const __half hi = __float2half_rn(v * SW_WSCALE);
  dst16[i] = hi;
  dst16[4 * H * H + i] = __float2half_rn(v * SW_WSCALE - __half2float(hi));
}
// fwd [2][4H][H] fp16, bwd [H][4H] bf16, bwd16 [2][H][4H] fp16
int pack_whh_swap(const float* w_hh, __half* fwd, __nv_bfloat16* bwd, __half* bwd16, int H, cudaStream_t st) {
  pack_whh_swap_fwd_kernel<<<ceil_div(4 * H * H, 256), 256, 0, st>>>(w_hh, fwd, 4 * H * H);
  BCI_LAUNCH_OK();
  pack_whh_swap_bwd_kernel<<<ceil_div(4 * H * H, 256), 256, 0, st>>>(w_hh, bwd, bwd16, H);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// one elected lane of a converged warp (the branch on it is warp-uniform for the compiler: tcgen05 instructions inside are issued
// back to back instead of inside a per-instruction election loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sw_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gate activations of the mixed mode: one MUFU each (sigma(x) = 1/2 + 1/2 tanh(x/2)); ~5e-4 absolute, the size of the fp16
// operand rounding.  The saved activations are the values actually used, so BPTT differentiates the computed function.
__device__ __forceinline__ float sw_sigmoid(float x) { return fmaf(tanh_mufu(0.5f * x), 0.5f, 0.5f); }

// Block layout: 16 epilogue warps (thread = hidden unit u = tid % 128, window pair tid / 128: with ONE warp per scheduler the
// per-step elementwise chain ran at its full dependent latency -- 2 400 cycles for 8 windows -- instead of its issue rate)
// + one MMA warp.  The two sides meet on two mbarriers only: acc_full (tcgen05.commit) and op_ready (one arrive per epilogue warp).
constexpr int SW_EPI = 512;
constexpr int SW_BLOCK = SW_EPI + 32;
constexpr int SW_WPT = SW_NW / (SW_EPI / 128);   // windows per thread: 2

struct SwCtx {
  uint32_t sA, sB, acc_full, op_ready, tmem;
  uint8_t* genB;
};
// prologue: barriers, TMEM (512 columns: accumulators at [0, 64), weights hi at [SW_WCOL, SW_WCOL + 256)), zeroed B tile, and the
// weights: thread (u, q) stores 64 columns = 128 sixteen-bit K elements of row u at `wrow` (forward: q = gate block, row q*128 + u
// of [512][128]; BPTT: q = K quarter of row u of [128][512]).
// SPLIT: the lo parts follow the same way for q < 3 at columns [SW_WLO, SW_WLO + 192) -- with the 64 accumulator columns that is
// all 512 -- and only the fourth quarter (`wsm`: 128 rows x 16 chunks of 8 halves, rows `wsm_stride` chunks apart) goes to shared
// memory as two K-major SWIZZLE_128B atoms in front of the B tiles (hi tile, then lo tile): 8 of the 32 W_lo MMAs of a step read
// their A operand through the shared-memory port (32 cycles each) instead of all 32
constexpr uint32_t SW_WLO = SW_WCOL + 256;
template <uint32_t B_BYTES, bool SPLIT>
__device__ __forceinline__ SwCtx sw_prologue(uint8_t* raw_ptr, const uint4* __restrict__ wrow, const uint4* __restrict__ wrow_lo,
                                             const uint4* __restrict__ wsm, int wsm_stride) {
  constexpr uint32_t B_ALL = SPLIT ? 2 * B_BYTES : B_BYTES;
  SwCtx c;
  const uint32_t raw = smem_u32(raw_ptr);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* genA = raw_ptr + (base - raw);
  c.sA = base;
  c.genB = genA + (SPLIT ? SW_ALO_BYTES : 0u);
  c.sB = base + (SPLIT ? SW_ALO_BYTES : 0u);
  uint8_t* ctl = c.genB + B_ALL;
  if (SPLIT) {
    for (int i = threadIdx.x; i < 128 * 16; i += SW_BLOCK) {
      const int row = i >> 4, cc = i & 15;
      *reinterpret_cast<uint4*>(genA + (uint32_t)(cc >> 3) * SW_AATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) =
          __ldg(wsm + (size_t)row * wsm_stride + cc);
    }
  }
  c.acc_full = smem_u32(ctl);
  c.op_ready = c.acc_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 16);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(c.acc_full, 1);
    mbar_init(c.op_ready, SW_EPI / 32);
    fence_mbar_init();
  }
  if (tid >= SW_EPI) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = tid; i < (int)(B_ALL / 16); i += SW_BLOCK) reinterpret_cast<uint4*>(c.genB)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;
  if (tid < SW_EPI) {
    const int u = tid & 127, q = tid >> 7;
    const uint32_t lane_base = c.tmem + ((uint32_t)((u >> 5) * 32) << 16);
#pragma unroll 1
    for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
      if (part == 1 && q == 3) break;   // the fourth lo quarter lives in shared memory
      const uint4* src = part ? wrow_lo : wrow;
      const uint32_t dst = lane_base + (part ? SW_WLO : SW_WCOL) + (uint32_t)q * 64u;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 v = __ldg(src + k * 4 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tmem_st16(dst + (uint32_t)k * 16u, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return c;
}

// dropout between LSTM layers (nn.LSTM(dropout=p), 04:181-188) folded into the recurrence kernels: the forward writes the dropped
// copy of h_t next to h_t (it was a separate pass over the layer output), BPTT multiplies the incoming gradient by the same mask
struct SwDrop {
  float* outd;       // forward: dropped copy of out (nullptr: none)
  float* outd_lo;    // forward: its tf32 remainder for the split-precision GEMM that reads it next (nullptr: not needed)
  float p;           // BPTT: p > 0 applies the mask of `site` to dout
  uint64_t seed;
  uint32_t site;
};
__device__ __forceinline__ float tf32_lo(float x) {   // x - trunc_tf32(x), rounded to tf32 (what split_tf32_kernel produces)
  const float rem = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
  return __uint_as_float(r);
}

// BCI_FUSED_JITTER (tests only; a power of two = the longest sleep in ns): every role sleeps a pseudo-random time at its
// synchronisation points, to shake out ordering assumptions that only hold at the natural timing (compute-sanitizer is not available
// on the GPU pool; lstm_bf16_fused.cu's first protocol bug only showed under such perturbation)
struct SwJitter {
  uint32_t state;
  int max_ns;
  __device__ __forceinline__ SwJitter(int m) : state((uint32_t)(blockIdx.x * 7919u + blockIdx.y * 131u + threadIdx.x * 104729u + 12345u)), max_ns(m) {}
  __device__ __forceinline__ void operator()() {
    if (max_ns) {
      state = state * 1664525u + 1013904223u;
      __nanosleep((state >> 20) & (uint32_t)(max_ns - 1));
    }
  }
};
static int sw_jitter() {
  static const int v = [] { const char* e = getenv("BCI_FUSED_JITTER"); int j = e ? atoi(e) : 0; return (j > 0 && (j & (j - 1)) == 0) ? j : 0; }();
  return v;
}

#define SW_STAMP(cond, i) do { if (dbg && st >= 100 && st < 104 && (cond) && blockIdx.x == 0 && blockIdx.y == 0) dbg[(st - 100) * 8 + (i)] = clock64(); } while (0)

// ---- forward --------------------------------------------------------------------------------------------------------------------
// grid = (ceil(Bc / 8), ND).  SPLIT = the fp32-parity form: h . W_hh^T as three fp16 product chains  lo.hi + hi.lo + hi.hi  (the small
// terms first: the tensor core adds into TMEM with truncation), accurate (ex2-based) gate activations -- the arithmetic of
// lstm_fp32_tc.cu's pair kernel; W_hi lives in tensor memory, W_lo (128 KB) in shared memory.
template <bool SPLIT>
__global__ void __launch_bounds__(SW_BLOCK, 1)
lstm_rec_swap_fwd(const float* __restrict__ G,        // [T*Bc][ldg] fp32: column dir*512 + unit*4 + gate, bias included
                  int ldg,
                  const __half* __restrict__ whh,     // [ND][2 parts][512][128] fp16 of 16 w, PyTorch row order
                  float* __restrict__ out,            // [T][Bc][D]: h_t at column dir*128 + unit
                  float* __restrict__ gates,          // optional [T*Bc][ldg] gate ACTIVATIONS, same layout as G
                  float* __restrict__ csave,          // optional [T*Bc][D] cell states
                  SwDrop dr,                          // optional dropped copy of h_t (the next layer's input), written here
                  int D, int Bc, int T, long long* __restrict__ dbg, int jitter, int g_half) {   // g_half: G is fp16 (mixed step)
  SwJitter jit(jitter);
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, u = tid & 127, wq = (tid >> 7) & 3;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int dir = blockIdx.y, b0 = blockIdx.x * SW_NW;
  const __half* wdir = whh + (size_t)dir * 2 * 512 * 128;
  const __half* wlo = wdir + 512 * 128;
  const SwCtx cx = sw_prologue<SW_FWD_B, SPLIT>(sw_smem_raw, reinterpret_cast<const uint4*>(wdir + ((size_t)wq * 128 + u) * 128),
                                                reinterpret_cast<const uint4*>(wlo + ((size_t)wq * 128 + u) * 128),
                                                reinterpret_cast<const uint4*>(wlo + (size_t)3 * 128 * 128), 16);

  if (warp_u == SW_EPI / 32) {
    // ---- MMA warp: 4 gate blocks x 8 K slices per step, A = resident weights in tensor memory, B = h_{t-1}
    for (int st = 0; st < T; ++st) {
      jit();
      if (st > 0) mbar_wait(cx.op_ready, (uint32_t)((st - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        SW_STAMP(true, 6);
        constexpr uint32_t idesc = sw_idesc(128, 16, false);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (SPLIT) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // W_lo . h_hi: gates i, f, g~ from tensor memory, o from shared memory
              const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
              if (g < 3) {
                umma_f16_ts(cx.tmem + g * 16, cx.tmem + SW_WLO + g * 64 + k * 8, db, idesc, k != 0 ? 1u : 0u);
              } else {
                const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)(k >> 2) * SW_AATOM + (uint32_t)(k & 3) * 32u);
                umma_bf16(cx.tmem + g * 16, da, db, idesc, k != 0 ? 1u : 0u);
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // W_hi (tensor memory) . h_lo
              const uint64_t db = umma_desc_sw128(cx.sB + SW_FWD_B + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
              umma_f16_ts(cx.tmem + g * 16, cx.tmem + SW_WCOL + g * 64 + k * 8, db, idesc, 1u);
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
            umma_f16_ts(cx.tmem + g * 16, cx.tmem + SW_WCOL + g * 64 + k * 8, db, idesc, (SPLIT || k != 0) ? 1u : 0u);
          }
        }
        umma_commit(cx.acc_full);
        SW_STAMP(true, 7);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue warps
    const uint32_t taddr = cx.tmem + ((uint32_t)((u >> 5) * 32) << 16) + (uint32_t)(wq * SW_WPT);
    uint32_t hoff[SW_WPT];
    int brow[SW_WPT];
#pragma unroll
    for (int i = 0; i < SW_WPT; ++i) {
      const int n = wq * SW_WPT + i;
      hoff[i] = (uint32_t)(u >> 6) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)n) << 4) + (uint32_t)(u & 7) * 2u;
      brow[i] = b0 + n < Bc ? b0 + n : Bc - 1;   // dead windows read a valid row and store nothing
    }
    const int colg = dir * 512 + u * 4, colh = dir * 128 + u;
    float c[SW_WPT];
    uint4 gq[SW_WPT];
    // the next step's G row is requested RAW into registers (fp32: 16 bytes, fp16: 8) and converted when it is used a step later:
    // converting at the load would stall on the DRAM round trip right here, in front of the accumulator wait
    auto load_g = [&](int tt, int i) -> uint4 {
      const long long e = ((long long)tt * Bc + brow[i]) * ldg + colg;
      if (!g_half) return __ldg(reinterpret_cast<const uint4*>(G + e));
      const uint2 gv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(G) + e));
      return make_uint4(gv.x, gv.y, 0u, 0u);
    };
    auto cvt_g = [&](const uint4& r) -> float4 {
      if (!g_half) return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
      return make_float4(a.x, a.y, b.x, b.y);
    };
    for (int i = 0; i < SW_WPT; ++i) {
      c[i] = 0.f;
      gq[i] = load_g(dir ? T - 1 : 0, i);
    }
    for (int st = 0; st < T; ++st) {
      const int t = dir ? (T - 1 - st) : st;
      SW_STAMP(tid == 0, 0);
      float4 gc[SW_WPT];
#pragma unroll
      for (int i = 0; i < SW_WPT; ++i) gc[i] = cvt_g(gq[i]);
      if (st + 1 < T) {
        const int tn = dir ? (T - 2 - st) : st + 1;
#pragma unroll
        for (int i = 0; i < SW_WPT; ++i) gq[i] = load_g(tn, i);
      }
      // the register prefetch above covers one step (~1.2 us): not always a DRAM round trip under load.  Pull the rows of
      // step st + 4 into L2 now (one request per 128-byte line)
      if (st + 4 < T && (tid & (g_half ? 15 : 7)) == 0) {
        const int t4 = dir ? (T - 5 - st) : st + 4;
#pragma unroll
        for (int i = 0; i < SW_WPT; ++i) {
          const long long e = ((long long)t4 * Bc + brow[i]) * ldg + colg;
          if (g_half) sw_prefetch_l2(reinterpret_cast<const __half*>(G) + e); else sw_prefetch_l2(G + e);
        }
      }
      float dsc[SW_WPT];   // dropout factors of this step's outputs: index arithmetic only, done while the product runs
      if (dr.outd) {
#pragma unroll
        for (int i = 0; i < SW_WPT; ++i) dsc[i] = drop_scale(dr.seed, dr.site, (uint64_t)(((long long)t * Bc + brow[i]) * D + colh), dr.p);
      }
      SW_STAMP(tid == 0, 1);
      jit();
      mbar_wait(cx.acc_full, (uint32_t)(st & 1));
      tc_fence_after();
      SW_STAMP(tid == 0, 2);
      uint32_t a[4][SW_WPT];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld2(taddr + g * 16, a[g]);
      tmem_ld_wait();
      SW_STAMP(tid == 0, 3);
#pragma unroll
      for (int i = 0; i < SW_WPT; ++i) {
        const float pi = fmaf(__uint_as_float(a[0][i]), 1.0f / SW_WSCALE, gc[i].x), pf = fmaf(__uint_as_float(a[1][i]), 1.0f / SW_WSCALE, gc[i].y);
        const float pg = fmaf(__uint_as_float(a[2][i]), 1.0f / SW_WSCALE, gc[i].z), po = fmaf(__uint_as_float(a[3][i]), 1.0f / SW_WSCALE, gc[i].w);
        const float ig = SPLIT ? rec_sigmoid(pi) : sw_sigmoid(pi);
        const float fg = SPLIT ? rec_sigmoid(pf) : sw_sigmoid(pf);
        const float gg = SPLIT ? rec_tanh(pg) : tanh_mufu(pg);
        const float og = SPLIT ? rec_sigmoid(po) : sw_sigmoid(po);
        c[i] = fmaf(fg, c[i], ig * gg);
        const float hv = og * (SPLIT ? rec_tanh(c[i]) : tanh_mufu(c[i]));
        const __half hh = __float2half_rn(hv);
        *reinterpret_cast<__half*>(cx.genB + hoff[i]) = hh;
        if (SPLIT) *reinterpret_cast<__half*>(cx.genB + SW_FWD_B + hoff[i]) = __float2half_rn(hv - __half2float(hh));
        if (b0 + wq * SW_WPT + i < Bc) {
          const long long row = (long long)t * Bc + b0 + wq * SW_WPT + i;
          out[row * D + colh] = hv;
          if (gates) {
            if (SPLIT) {
              *reinterpret_cast<float4*>(gates + row * ldg + colg) = make_float4(ig, fg, gg, og);
            } else {   // mixed mode: the saved activations are fp16 (values in [-1, 1]; half the bytes of an HBM-bound kernel)
              const __half2 lo2 = __floats2half2_rn(ig, fg), hi2 = __floats2half2_rn(gg, og);
              *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(gates) + row * ldg + colg) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
            }
          }
          if (csave) csave[row * D + colh] = c[i];
          if (dr.outd) {
            const float hd = hv * dsc[i];
            dr.outd[row * D + colh] = hd;
            if (dr.outd_lo) dr.outd_lo[row * D + colh] = tf32_lo(hd);
          }
        }
      }
      SW_STAMP(tid == 0, 4);
      fence_proxy_async_smem();  // h_t (generic-proxy stores) -> visible to the next step's tcgen05.mma
      tc_fence_before();         // this thread's TMEM reads are ordered before the arrive
      __syncwarp();
      if ((tid & 31) == 0) { jit(); mbar_arrive(cx.op_ready); }
      SW_STAMP(tid == 0, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_u == SW_EPI / 32) {
    tc_fence_after();
    tmem_dealloc(cx.tmem, 512);
  }
}

// ---- BPTT -----------------------------------------------------------------------------------------------------------------------
// The mirror of lstm_bptt_f32 (lstm_train.cu): walks the direction's time order backwards, dG_t to global (fp32, + optional tf32
// remainder for the split-precision GEMMs) and, as bf16, into the B tile of  dh_{t-1}[j] = sum_k dG_t[k] W_hh[k][j].
//
// SPLIT = the fp32-parity form.  Gradients do not fit fp16's range as they come, so every step's dG tile is multiplied by a power
// of two S before it is split into an fp16 (hi, lo) pair: S puts the PREVIOUS step's largest |dG| of this CTA at 2^8 (the maximum
// is collected with one shared-memory atomicMax per warp and read a step later -- no extra barrier; three rotating slots), which
// leaves a factor 128 of growth per step before the saturating conversion clips, 22 significant bits for every element within 2^-10
// of the tile's maximum and an absolute error of 2^-33 of that maximum below.  dh = D / (16 S) is exact.  Three product chains as in
// the forward: W_lo (shared memory) . dG_hi, W_hi (tensor memory) . dG_lo, W_hi . dG_hi.
__device__ __forceinline__ uint16_t f2h_sat(float x) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
template <bool SPLIT>
__global__ void __launch_bounds__(SW_BLOCK, 1)
lstm_bptt_swap(const float* __restrict__ dout,           // [T][Bc][D]
               const float* __restrict__ gates,          // [T*Bc][ldg]: i,f,g,o of (dir, unit) at column dir*512 + unit*4
               const float* __restrict__ csave,          // [T*Bc][D]
               const void* __restrict__ whhT_v,          // mixed: [ND][128 j][512 k = gate*128 + unit] bf16; SPLIT: [ND][2][128][512] fp16 of 16 w
               float* __restrict__ dG,                   // [T*Bc][ldg]
               float* __restrict__ dG_lo,                // optional
               float* __restrict__ dbias,                // optional [ldg]: += sum over rows of dG (the bias gradient, b_ih = b_hh)
               SwDrop dr,                                // dr.p > 0: dout is the gradient wrt the DROPPED layer output (mask of dr.site)
               int ldg, int D, int Bc, int T, int jitter) {
  SwJitter jit(jitter);
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, u = tid & 127, wq = (tid >> 7) & 3;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int dir = blockIdx.y, b0 = blockIdx.x * SW_NW;
  const uint16_t* whhT = reinterpret_cast<const uint16_t*>(whhT_v) + (size_t)dir * (SPLIT ? 2 : 1) * 128 * 512;
  const SwCtx cx = sw_prologue<SW_BWD_B, SPLIT>(sw_smem_raw, reinterpret_cast<const uint4*>(whhT + (size_t)u * 512 + wq * 128),
                                                reinterpret_cast<const uint4*>(whhT + 128 * 512 + (size_t)u * 512 + wq * 128),
                                                reinterpret_cast<const uint4*>(whhT + 128 * 512 + 384), 64);
  uint32_t* mx = reinterpret_cast<uint32_t*>(cx.genB + (SPLIT ? 2 : 1) * SW_BWD_B + 32);   // three slots behind the barriers / TMEM slot
  if (SPLIT && tid < 3) mx[tid] = 0u;
  if (SPLIT) __syncthreads();

  if (warp_u == SW_EPI / 32) {
    for (int s = T - 1; s > 0; --s) {
      jit();
      mbar_wait(cx.op_ready, (uint32_t)((T - 1 - s) & 1));
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = sw_idesc(128, 16, !SPLIT);
        if (SPLIT) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {   // W_lo . dG_hi: K quarters 0-2 from tensor memory, quarter 3 from shared memory
            const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
            if (k < 24) {
              umma_f16_ts(cx.tmem, cx.tmem + SW_WLO + k * 8, db, idesc, k != 0 ? 1u : 0u);
            } else {
              const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)((k - 24) >> 2) * SW_AATOM + (uint32_t)(k & 3) * 32u);
              umma_bf16(cx.tmem, da, db, idesc, 1u);
            }
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) {   // W_hi (tensor memory) . dG_lo
            const uint64_t db = umma_desc_sw128(cx.sB + SW_BWD_B + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
            umma_f16_ts(cx.tmem, cx.tmem + SW_WCOL + k * 8, db, idesc, 1u);
          }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
          umma_f16_ts(cx.tmem, cx.tmem + SW_WCOL + k * 8, db, idesc, (SPLIT || k != 0) ? 1u : 0u);
        }
        umma_commit(cx.acc_full);
      }
      __syncwarp();
    }
  } else {
    const uint32_t taddr = cx.tmem + ((uint32_t)((u >> 5) * 32) << 16) + (uint32_t)(wq * SW_WPT);
    // element (row n, k = gate*128 + u) of the B tile: atom gate*2 + u/64
    uint32_t boff[SW_WPT];
    int brow[SW_WPT];
#pragma unroll
    for (int i = 0; i < SW_WPT; ++i) {
      const int n = wq * SW_WPT + i;
      boff[i] = (uint32_t)(u >> 6) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)n) << 4) + (uint32_t)(u & 7) * 2u;
      brow[i] = b0 + n < Bc ? b0 + n : Bc - 1;
    }
    const int colg = dir * 512 + u * 4, colh = dir * 128 + u;
    float dh_rec[SW_WPT], dc[SW_WPT];
    float4 pg[SW_WPT];
    float pc[SW_WPT], pcp[SW_WPT], pdo[SW_WPT];
#pragma unroll
    for (int i = 0; i < SW_WPT; ++i) { dh_rec[i] = 0.f; dc[i] = 0.f; }
    // everything fetch() requests is left RAW in registers (fp16 gate bits unconverted, dout unscaled: the dropout factor goes to
    // sc[]): any arithmetic on a loaded value here would stall on its DRAM round trip in front of the accumulator wait
    float pds[SW_WPT];
    auto fetch = [&](int s, float4* g4, float* cc, float* cp, float* dd, float* sc) {
      const int t = dir ? (T - 1 - s) : s;
      const int tp = dir ? (t + 1) : (t - 1);
#pragma unroll
      for (int i = 0; i < SW_WPT; ++i) {
        const long long row = (long long)t * Bc + brow[i];
        if (SPLIT) {
          g4[i] = __ldg(reinterpret_cast<const float4*>(gates + row * ldg + colg));
        } else {   // mixed mode: fp16 activations (bits in .x / .y)
          const uint2 gv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(gates) + row * ldg + colg));
          g4[i] = make_float4(__uint_as_float(gv.x), __uint_as_float(gv.y), 0.f, 0.f);
        }
        if (cc) cc[i] = __ldg(csave + row * D + colh);
        cp[i] = (s > 0) ? __ldg(csave + ((long long)tp * Bc + brow[i]) * D + colh) : 0.f;
        dd[i] = __ldg(dout + row * D + colh);
        sc[i] = dr.p > 0.f ? drop_scale(dr.seed, dr.site, (uint64_t)(row * D + colh), dr.p) : 1.0f;
      }
    };
    auto gate_vals = [&](const float4& r) -> float4 {
      if (SPLIT) return r;
      const uint32_t x = __float_as_uint(r.x), y = __float_as_uint(r.y);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&x)), b = __half22float2(*reinterpret_cast<const __half2*>(&y));
      return make_float4(a.x, a.y, b.x, b.y);
    };
    fetch(T - 1, pg, pc, pcp, pdo, pds);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);   // this thread's share of the bias gradient: its windows, all steps
    for (int s = T - 1; s >= 0; --s) {
      const int t = dir ? (T - 1 - s) : s;
      const int it = T - 1 - s;
      float4 dgs[SW_WPT];
      float lmax = 0.f;
#pragma unroll
      for (int i = 0; i < SW_WPT; ++i) {
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + wq * SW_WPT + i < Bc) {
          const float4 g = gate_vals(pg[i]);
          const float dh = fmaf(pdo[i], pds[i], dh_rec[i]);
          const float tc = SPLIT ? rec_tanh(pc[i]) : tanh_mufu(pc[i]);  // the forward's own tanh(c)
          const float dct = fmaf(dh * g.w, 1.0f - tc * tc, dc[i]);
          dg.x = dct * g.z * g.x * (1.0f - g.x);
          dg.y = dct * pcp[i] * g.y * (1.0f - g.y);
          dg.z = dct * g.x * (1.0f - g.z * g.z);
          dg.w = dh * tc * g.w * (1.0f - g.w);
          dc[i] = dct * g.y;
          bsum.x += dg.x; bsum.y += dg.y; bsum.z += dg.z; bsum.w += dg.w;
          const long long row = (long long)t * Bc + b0 + wq * SW_WPT + i;
          *reinterpret_cast<float4*>(dG + row * ldg + colg) = dg;
          if (dG_lo) {
            auto lo = [](float x) {
              const float rem = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
              uint32_t r;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
              return __uint_as_float(r);
            };
            *reinterpret_cast<float4*>(dG_lo + row * ldg + colg) = make_float4(lo(dg.x), lo(dg.y), lo(dg.z), lo(dg.w));
          }
        }
        if (!SPLIT && s > 0) {
          *reinterpret_cast<__nv_bfloat16*>(cx.genB + 0 * 2 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.x);
          *reinterpret_cast<__nv_bfloat16*>(cx.genB + 1 * 2 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.y);
          *reinterpret_cast<__nv_bfloat16*>(cx.genB + 2 * 2 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.z);
          *reinterpret_cast<__nv_bfloat16*>(cx.genB + 3 * 2 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.w);
        }
        if (SPLIT) {
          dgs[i] = dg;
          lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(dg.x), fabsf(dg.y)), fmaxf(fabsf(dg.z), fabsf(dg.w))));
        }
      }
      if (s == 0) {
        if (dbias) {
          atomicAdd(dbias + colg + 0, bsum.x); atomicAdd(dbias + colg + 1, bsum.y);
          atomicAdd(dbias + colg + 2, bsum.z); atomicAdd(dbias + colg + 3, bsum.w);
        }
        break;
      }
      float inv_scale = 1.0f;
      if (SPLIT) {
        // this CTA's largest |dG| of this step -> slot it % 3 (read by step it + 1); slot (it + 1) % 3 was last read in step it - 1
        lmax = warp_max(lmax);
        if ((tid & 31) == 0) atomicMax(&mx[it % 3], __float_as_uint(lmax));
        if (tid == 0) mx[(it + 1) % 3] = 0u;
        if (it == 0) asm volatile("bar.sync 1, %0;" ::"n"(SW_EPI) : "memory");   // first step: its own maximum
        const uint32_t pm = *reinterpret_cast<volatile uint32_t*>(&mx[it == 0 ? 0 : (it - 1) % 3]);
        int se = 262 - (int)((pm >> 23) & 0xffu);     // biased exponent of S: largest |dG| . S in [2^8, 2^9)
        se = se > 187 ? 187 : (se < 40 ? 40 : se);    // an all-zero tile (exponent field 0) or absurd magnitudes: clamp S to [2^-87, 2^60]
        const float S = __uint_as_float((uint32_t)se << 23);
        inv_scale = __uint_as_float((uint32_t)(254 - se) << 23) * (1.0f / SW_WSCALE);
#pragma unroll
        for (int i = 0; i < SW_WPT; ++i) {
          const float v[4] = {dgs[i].x * S, dgs[i].y * S, dgs[i].z * S, dgs[i].w * S};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint16_t hi = f2h_sat(v[g]);
            const uint16_t lo = f2h_sat(v[g] - __half2float(__ushort_as_half(hi)));
            *reinterpret_cast<uint16_t*>(cx.genB + g * 2 * SW_BATOM + boff[i]) = hi;
            *reinterpret_cast<uint16_t*>(cx.genB + SW_BWD_B + g * 2 * SW_BATOM + boff[i]) = lo;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) { jit(); mbar_arrive(cx.op_ready); }
      // step s-1's saved activations are requested while the product runs; c(s-1) is this step's cprev
      float4 ng[SW_WPT];
      float ncp[SW_WPT], ndo[SW_WPT], nds[SW_WPT];
      fetch(s - 1, ng, nullptr, ncp, ndo, nds);
      if (s >= 4) {   // rows of step s - 4 -> L2
        const int t4 = dir ? (T - 1 - (s - 4)) : s - 4;
#pragma unroll
        for (int i = 0; i < SW_WPT; ++i) {
          const long long row = (long long)t4 * Bc + brow[i];
          if (SPLIT) { if ((tid & 7) == 0) sw_prefetch_l2(gates + row * ldg + colg); }
          else if ((tid & 15) == 0) sw_prefetch_l2(reinterpret_cast<const __half*>(gates) + row * ldg + colg);
          if ((tid & 31) == 0) { sw_prefetch_l2(csave + row * D + colh); sw_prefetch_l2(dout + row * D + colh); }
        }
      }
      mbar_wait(cx.acc_full, (uint32_t)((T - 1 - s) & 1));
      tc_fence_after();
      uint32_t a[SW_WPT];
      tmem_ld2(taddr, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < SW_WPT; ++i) {
        dh_rec[i] = SPLIT ? __uint_as_float(a[i]) * inv_scale : __uint_as_float(a[i]);
        pc[i] = pcp[i]; pg[i] = ng[i]; pcp[i] = ncp[i]; pdo[i] = ndo[i]; pds[i] = nds[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_u == SW_EPI / 32) {
    tc_fence_after();
    tmem_dealloc(cx.tmem, 512);
  }
}

static int sw_setup() {
  static PerDeviceFlag done_pd;
  bool& done = done_pd.cur();
  if (!done) {
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_swap_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem_bytes(SW_FWD_B, true)));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_swap<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sw_smem_bytes(SW_BWD_B, true)));
    done = true;
  }
  return BCI_OK;
}
static long long* g_sw_dbg = nullptr;   // selftest only: clock stamps of CTA (0, 0)

bool rec_swap_ok(int H, const void* G, int ldg) { return H == 128 && ((uintptr_t)G & 15) == 0 && (ldg & 3) == 0; }

// BCI_TRAIN_REC=simt keeps the CUDA-core recurrences of round 1 in the fp32-parity training step and the small-batch fp32 forward
bool swap_rec_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_TRAIN_REC");
    v = (e && e[0] == 's') ? 0 : 1;
  }
  return v != 0;
}

// the 16-bit recurrent operands of every layer and direction, from the raw weights of the last load (no-op until the next load)
int pack_swap_operands(bci_lstm_s* h, cudaStream_t st) {
  if (!h->sw_stale) return BCI_OK;
  const int H = h->cfg.hidden_size, ND = num_dirs(h->cfg);
  for (int l = 0; l < h->cfg.num_layers; ++l)
    for (int d = 0; d < ND; ++d) {
      int rc = pack_whh_swap(h->raw.w_hh[l][d], h->f32.whh_sw_f[l] + (size_t)d * 2 * 4 * H * H, h->f32.whh_sw_b[l] + (size_t)d * 4 * H * H,
                             h->f32.whh_sw_b16[l] + (size_t)d * 2 * 4 * H * H, H, st);
      if (rc) return rc;
    }
  h->sw_stale = false;
  return BCI_OK;
}

int launch_rec_swap_fwd(int ND, const float* G, int ldg, const __half* whh, float* out, float* gates, float* csave, int D, int Bc, int T,
                        bool split, cudaStream_t st, const SwapDropout* drop, bool g_half) {
  int rc = sw_setup();
  if (rc) return rc;
  const dim3 grid(ceil_div(Bc, SW_NW), ND);
  const SwDrop dr = drop ? SwDrop{drop->outd, drop->outd_lo, drop->p, drop->seed, drop->site} : SwDrop{nullptr, nullptr, 0.f, 0, 0};
  if (split) lstm_rec_swap_fwd<true><<<grid, SW_BLOCK, sw_smem_bytes(SW_FWD_B, true), st>>>(G, ldg, whh, out, gates, csave, dr, D, Bc, T, g_sw_dbg, sw_jitter(), 0);
  else lstm_rec_swap_fwd<false><<<grid, SW_BLOCK, sw_smem_bytes(SW_FWD_B, false), st>>>(G, ldg, whh, out, gates, csave, dr, D, Bc, T, g_sw_dbg, sw_jitter(), g_half ? 1 : 0);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// whhT: split ? the fp16 (hi, lo) pair [ND][2][128][512] : bf16 [ND][128][512]
int launch_bptt_swap(int ND, const float* dout, const float* gates, const float* csave, const void* whhT, float* dG, float* dG_lo,
                     float* dbias, int ldg, int D, int Bc, int T, bool split, cudaStream_t st, const SwapDropout* drop) {
  int rc = sw_setup();
  if (rc) return rc;
  const dim3 grid(ceil_div(Bc, SW_NW), ND);
  const SwDrop dr = drop ? SwDrop{nullptr, nullptr, drop->p, drop->seed, drop->site} : SwDrop{nullptr, nullptr, 0.f, 0, 0};
  if (split) lstm_bptt_swap<true><<<grid, SW_BLOCK, sw_smem_bytes(SW_BWD_B, true), st>>>(dout, gates, csave, whhT, dG, dG_lo, dbias, dr, ldg, D, Bc, T, sw_jitter());
  else lstm_bptt_swap<false><<<grid, SW_BLOCK, sw_smem_bytes(SW_BWD_B, false), st>>>(dout, gates, csave, whhT, dG, dG_lo, dbias, dr, ldg, D, Bc, T, sw_jitter());
  BCI_LAUNCH_OK();
  return BCI_OK;
}


// =====================================================================================================================================
// hidden_size 256 (the reference's trained checkpoint, 04_lstm_model.py:876-877), mixed precision: the same swapped recurrences on a
// CTA PAIR.  W_hh is 512 KB in 16 bits -- neither one SM's tensor memory (256 KB, of which the accumulators need a share) nor its
// shared memory holds it, so the 256 hidden units are split between the two CTAs of a cluster: each owns 128 units, i.e. 128 rows of
// every gate block (forward) / 128 rows j of W_hh^T (BPTT), for the SAME 16 windows.  Per CTA that is 256 KB of weights: three
// quarters live in tensor memory (384 columns), the fourth (64 KB) in shared memory as the classic K-major SW128 A operand.
//
//   forward  per gate g:  D_g[u][n] = sum_{k < 256} W_hh[g*256 + 128 r + u][k] . h_{t-1}[n][k]     gates i, f, g~ from TMEM, o from smem
//   BPTT                  D[j][n]   = sum_{k < 1024} W_hh^T[128 r + j][k] . dG_t[n][k]             K quarters 0-2 from TMEM, 3 from smem
//
// The cell update / gate backward stay thread-local (a thread owns one unit and four of the 16 windows).  What the pair must exchange
// every step is the B operand: each CTA produces h_t (forward, 4 KB) / dG_t (BPTT, 16 KB) of ITS 128 units and needs the other
// half.  The producer's MMA warp pushes its half into the same place of the peer's B tile with cp.async.bulk shared::cta ->
// shared::cluster, counted on the peer's mbarrier x_in; B tiles and x_in are double-buffered by step parity, so there is no
// "receive buffer free" handshake (a CTA can only be one step ahead of its peer: it needs the peer's half of every step), and
// x_in[b] is re-armed by its single waiter right after each completed phase (lstm_bf16_fused.cu's protocol).
constexpr int S2_NW = 16, S2_WPT = 4;
constexpr uint32_t S2_FB = 4 * SW_BATOM;      // forward B tile: 16 windows x K 256 = 8 KB
constexpr uint32_t S2_BB = 16 * SW_BATOM;     // BPTT B tile:    16 windows x K 1024 = 32 KB
constexpr uint32_t S2_WSM = 4 * SW_AATOM;     // the weight quarter kept in shared memory: 128 rows x K 256 = 64 KB
constexpr size_t s2_smem_bytes(uint32_t b_bytes) { return 1024 + S2_WSM + 2 * b_bytes + 128; }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

struct S2Ctx {
  uint32_t sA, sB, acc_full, op_ready, x_in, tmem, rank;
  uint8_t* genB;
};
// wsm: this CTA's 128 rows of the shared-memory quarter, row-major [128][32 chunks of 8 halves] with row stride `wsm_stride` chunks;
// wrow: this thread's row of the tensor-memory quarters (quarter q, this thread's 64-element slice at wrow + q * q_stride chunks)
template <uint32_t B_BYTES, uint32_t D_COLS>
__device__ __forceinline__ S2Ctx s2_prologue(uint8_t* raw_ptr, const uint4* __restrict__ wsm, int wsm_stride, const uint4* __restrict__ wrow,
                                             int q_stride, uint32_t tx_bytes) {
  S2Ctx c;
  const uint32_t raw = smem_u32(raw_ptr);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* genA = raw_ptr + (base - raw);
  c.sA = base;
  c.genB = genA + S2_WSM;
  c.sB = base + S2_WSM;
  uint8_t* ctl = c.genB + 2 * B_BYTES;
  c.acc_full = smem_u32(ctl);
  c.op_ready = c.acc_full + 8;
  c.x_in = c.acc_full + 16;   // two barriers
  c.rank = cluster_ctarank();
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(c.acc_full, 1);
    mbar_init(c.op_ready, SW_EPI / 32);
    mbar_init(c.x_in, 1);
    mbar_init(c.x_in + 8, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(c.x_in, tx_bytes);
    mbar_arrive_expect_tx(c.x_in + 8, tx_bytes);
  }
  if (tid >= SW_EPI) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 128 * 32; i += SW_BLOCK) {
    const int row = i >> 5, cc = i & 31;
    *reinterpret_cast<uint4*>(genA + (uint32_t)(cc >> 3) * SW_AATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) =
        __ldg(wsm + (size_t)row * wsm_stride + cc);
  }
  for (int i = tid; i < (int)(2 * B_BYTES / 16); i += SW_BLOCK) reinterpret_cast<uint4*>(c.genB)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;
  if (tid < SW_EPI) {
    const int u = tid & 127, wq = tid >> 7;
    const uint32_t dst = c.tmem + ((uint32_t)((u >> 5) * 32) << 16) + D_COLS + (uint32_t)wq * 32u;
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 v = __ldg(wrow + (size_t)q * q_stride + k * 4 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tmem_st16(dst + (uint32_t)q * 128u + (uint32_t)k * 16u, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // the peer's barriers are initialised and armed before anything is pushed to it
  return c;
}

// grid = (2 * ceil(Bc / 16), ND), cluster (2, 1, 1)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SW_BLOCK, 1)
lstm_rec_swap256_fwd(const float* __restrict__ G,        // [T*Bc][ldg] fp32: column dir*1024 + unit*4 + gate, bias included
                     int ldg,
                     const __half* __restrict__ whh,     // [ND][2 parts][1024][256] fp16 of 16 w (the hi part is used)
                     float* __restrict__ out,            // [T][Bc][D]: column dir*256 + unit
                     float* __restrict__ gates, float* __restrict__ csave, SwDrop dr, int D, int Bc, int T, int jitter, int g_half) {
  SwJitter jit(jitter);
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, u = tid & 127, wq = (tid >> 7) & 3;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int dir = blockIdx.y, b0 = (int)(blockIdx.x >> 1) * S2_NW;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const __half* wdir = whh + (size_t)dir * 2 * 1024 * 256;
  // shared-memory quarter = the o gate's rows of this CTA; tensor-memory quarters q = gates i, f, g~ (row q*256 + 128 r + u)
  const S2Ctx cx = s2_prologue<S2_FB, 64>(sw_smem_raw, reinterpret_cast<const uint4*>(wdir + ((size_t)3 * 256 + 128 * rank) * 256), 32,
                                          reinterpret_cast<const uint4*>(wdir + ((size_t)128 * rank + u) * 256 + wq * 64), 256 * 32, 4096u);

  if (warp_u == SW_EPI / 32) {
    for (int st = 0; st < T; ++st) {
      const uint32_t buf = (uint32_t)(st & 1), sBb = cx.sB + buf * S2_FB, xin = cx.x_in + 8u * buf;
      if (st > 0) {
        jit();
        mbar_wait(cx.op_ready, (uint32_t)((st - 1) & 1));   // this CTA's half of h_{st-1} is in B[buf]
        jit();
        if (elect_one()) {
          const uint32_t mine = sBb + 2u * rank * SW_BATOM;
          bulk_copy_s2s_cluster(mapa_u32(mine, peer), mine, 4096u, mapa_u32(xin, peer));
        }
        jit();
        mbar_wait(xin, (uint32_t)(((st - 1) >> 1) & 1));    // the peer's half has landed
        jit();
        if (elect_one()) mbar_arrive_expect_tx(xin, 4096u);  // re-arm for step st + 2
      }
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = sw_idesc(128, 16, false);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t db = umma_desc_sw128(sBb + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
            if (g < 3) {
              umma_f16_ts(cx.tmem + g * 16, cx.tmem + 64 + g * 128 + k * 8, db, idesc, k != 0 ? 1u : 0u);
            } else {
              const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)(k >> 2) * SW_AATOM + (uint32_t)(k & 3) * 32u);
              umma_bf16(cx.tmem + g * 16, da, db, idesc, k != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(cx.acc_full);
      }
      __syncwarp();
    }
  } else {
    const uint32_t taddr = cx.tmem + ((uint32_t)((u >> 5) * 32) << 16) + (uint32_t)(wq * S2_WPT);
    const int unit = 128 * (int)rank + u;
    uint32_t hoff[S2_WPT];
    int brow[S2_WPT];
#pragma unroll
    for (int i = 0; i < S2_WPT; ++i) {
      const int n = wq * S2_WPT + i;
      hoff[i] = (uint32_t)(2 * rank + (u >> 6)) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(u & 7) * 2u;
      brow[i] = b0 + n < Bc ? b0 + n : Bc - 1;
    }
    const int colg = dir * 1024 + unit * 4, colh = dir * 256 + unit;
    float c[S2_WPT];
    uint4 gq[S2_WPT];
#pragma unroll
    // the next step's G row is requested RAW into registers (fp32: 16 bytes, fp16: 8) and converted when it is used a step later:
    // converting at the load would stall on the DRAM round trip right here, in front of the accumulator wait
    auto load_g = [&](int tt, int i) -> uint4 {
      const long long e = ((long long)tt * Bc + brow[i]) * ldg + colg;
      if (!g_half) return __ldg(reinterpret_cast<const uint4*>(G + e));
      const uint2 gv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(G) + e));
      return make_uint4(gv.x, gv.y, 0u, 0u);
    };
    auto cvt_g = [&](const uint4& r) -> float4 {
      if (!g_half) return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
      return make_float4(a.x, a.y, b.x, b.y);
    };
    for (int i = 0; i < S2_WPT; ++i) {
      c[i] = 0.f;
      gq[i] = load_g(dir ? T - 1 : 0, i);
    }
    for (int st = 0; st < T; ++st) {
      const int t = dir ? (T - 1 - st) : st;
      uint8_t* Bn = cx.genB + (uint32_t)((st + 1) & 1) * S2_FB;   // h_t is the B operand of step st + 1
      float4 gc[S2_WPT];
#pragma unroll
      for (int i = 0; i < S2_WPT; ++i) gc[i] = cvt_g(gq[i]);
      if (st + 1 < T) {
        const int tn = dir ? (T - 2 - st) : st + 1;
#pragma unroll
        for (int i = 0; i < S2_WPT; ++i) gq[i] = load_g(tn, i);
      }
      if (st + 4 < T && (tid & (g_half ? 15 : 7)) == 0) {
        const int t4 = dir ? (T - 5 - st) : st + 4;
#pragma unroll
        for (int i = 0; i < S2_WPT; ++i) {
          const long long e = ((long long)t4 * Bc + brow[i]) * ldg + colg;
          if (g_half) sw_prefetch_l2(reinterpret_cast<const __half*>(G) + e); else sw_prefetch_l2(G + e);
        }
      }
      float dsc[S2_WPT];
      if (dr.outd) {
#pragma unroll
        for (int i = 0; i < S2_WPT; ++i) dsc[i] = drop_scale(dr.seed, dr.site, (uint64_t)(((long long)t * Bc + brow[i]) * D + colh), dr.p);
      }
      jit();
      mbar_wait(cx.acc_full, (uint32_t)(st & 1));
      tc_fence_after();
      uint32_t a[4][S2_WPT];
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_ld4(taddr + g * 16, a[g]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < S2_WPT; ++i) {
        const float ig = sw_sigmoid(fmaf(__uint_as_float(a[0][i]), 1.0f / SW_WSCALE, gc[i].x));
        const float fg = sw_sigmoid(fmaf(__uint_as_float(a[1][i]), 1.0f / SW_WSCALE, gc[i].y));
        const float gg = tanh_mufu(fmaf(__uint_as_float(a[2][i]), 1.0f / SW_WSCALE, gc[i].z));
        const float og = sw_sigmoid(fmaf(__uint_as_float(a[3][i]), 1.0f / SW_WSCALE, gc[i].w));
        c[i] = fmaf(fg, c[i], ig * gg);
        const float hv = og * tanh_mufu(c[i]);
        *reinterpret_cast<__half*>(Bn + hoff[i]) = __float2half_rn(hv);
        if (b0 + wq * S2_WPT + i < Bc) {
          const long long row = (long long)t * Bc + b0 + wq * S2_WPT + i;
          out[row * D + colh] = hv;
          if (gates) {   // fp16, as in the H = 128 mixed kernel
            const __half2 lo2 = __floats2half2_rn(ig, fg), hi2 = __floats2half2_rn(gg, og);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(gates) + row * ldg + colg) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
          }
          if (csave) csave[row * D + colh] = c[i];
          if (dr.outd) dr.outd[row * D + colh] = hv * dsc[i];
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) { jit(); mbar_arrive(cx.op_ready); }
    }
  }
  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still push into / wait on this CTA
  if (warp_u == SW_EPI / 32) {
    tc_fence_after();
    tmem_dealloc(cx.tmem, 512);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SW_BLOCK, 1)
lstm_bptt_swap256(const float* __restrict__ dout,           // [T][Bc][D]
                  const float* __restrict__ gates,          // [T*Bc][ldg]: column dir*1024 + unit*4
                  const float* __restrict__ csave,          // [T*Bc][D]
                  const __nv_bfloat16* __restrict__ whhT,   // [ND][256 j][1024 k = gate*256 + unit] bf16
                  float* __restrict__ dG, float* __restrict__ dbias, SwDrop dr, int ldg, int D, int Bc, int T, int jitter) {
  SwJitter jit(jitter);
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, u = tid & 127, wq = (tid >> 7) & 3;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int dir = blockIdx.y, b0 = (int)(blockIdx.x >> 1) * S2_NW;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const __nv_bfloat16* wdir = whhT + (size_t)dir * 256 * 1024;
  // this CTA's rows j = 128 r + u; K quarters 0-2 in tensor memory, quarter 3 (k in [768, 1024): the o gate) in shared memory
  const S2Ctx cx = s2_prologue<S2_BB, 32>(sw_smem_raw, reinterpret_cast<const uint4*>(wdir + ((size_t)128 * rank) * 1024 + 768), 128,
                                          reinterpret_cast<const uint4*>(wdir + ((size_t)128 * rank + u) * 1024 + wq * 64), 32, 16384u);

  if (warp_u == SW_EPI / 32) {
    for (int it = 0; it + 1 < T; ++it) {
      const uint32_t buf = (uint32_t)(it & 1), sBb = cx.sB + buf * S2_BB, xin = cx.x_in + 8u * buf;
      jit();
      mbar_wait(cx.op_ready, (uint32_t)(it & 1));   // this CTA's half of dG is in B[buf]
      jit();
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // gate g: atoms 4 g + 2 r, + 1
          const uint32_t mine = sBb + (uint32_t)(4 * g + 2 * rank) * SW_BATOM;
          bulk_copy_s2s_cluster(mapa_u32(mine, peer), mine, 4096u, mapa_u32(xin, peer));
        }
      }
      jit();
      mbar_wait(xin, (uint32_t)((it >> 1) & 1));
      jit();
      if (elect_one()) mbar_arrive_expect_tx(xin, 16384u);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = sw_idesc(128, 16, true);
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const uint64_t db = umma_desc_sw128(sBb + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
          if (k < 48) {
            umma_f16_ts(cx.tmem, cx.tmem + 32 + k * 8, db, idesc, k != 0 ? 1u : 0u);
          } else {
            const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)((k - 48) >> 2) * SW_AATOM + (uint32_t)(k & 3) * 32u);
            umma_bf16(cx.tmem, da, db, idesc, 1u);
          }
        }
        umma_commit(cx.acc_full);
      }
      __syncwarp();
    }
  } else {
    const uint32_t taddr = cx.tmem + ((uint32_t)((u >> 5) * 32) << 16) + (uint32_t)(wq * S2_WPT);
    const int unit = 128 * (int)rank + u;
    uint32_t boff[S2_WPT];
    int brow[S2_WPT];
#pragma unroll
    for (int i = 0; i < S2_WPT; ++i) {
      const int n = wq * S2_WPT + i;
      boff[i] = (uint32_t)(2 * rank + (u >> 6)) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)(n & 7)) << 4) + (uint32_t)(u & 7) * 2u;
      brow[i] = b0 + n < Bc ? b0 + n : Bc - 1;
    }
    const int colg = dir * 1024 + unit * 4, colh = dir * 256 + unit;
    float dh_rec[S2_WPT], dc[S2_WPT];
    float4 pg[S2_WPT];
    float pc[S2_WPT], pcp[S2_WPT], pdo[S2_WPT];
#pragma unroll
    for (int i = 0; i < S2_WPT; ++i) { dh_rec[i] = 0.f; dc[i] = 0.f; }
    float pds[S2_WPT];
    auto fetch = [&](int s, float4* g4, float* cc, float* cp, float* dd, float* sc) {   // raw, as in lstm_bptt_swap
      const int t = dir ? (T - 1 - s) : s;
      const int tp = dir ? (t + 1) : (t - 1);
#pragma unroll
      for (int i = 0; i < S2_WPT; ++i) {
        const long long row = (long long)t * Bc + brow[i];
        const uint2 gv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(gates) + row * ldg + colg));
        g4[i] = make_float4(__uint_as_float(gv.x), __uint_as_float(gv.y), 0.f, 0.f);
        if (cc) cc[i] = __ldg(csave + row * D + colh);
        cp[i] = (s > 0) ? __ldg(csave + ((long long)tp * Bc + brow[i]) * D + colh) : 0.f;
        dd[i] = __ldg(dout + row * D + colh);
        sc[i] = dr.p > 0.f ? drop_scale(dr.seed, dr.site, (uint64_t)(row * D + colh), dr.p) : 1.0f;
      }
    };
    auto gate_vals = [&](const float4& r) -> float4 {
      const uint32_t x = __float_as_uint(r.x), y = __float_as_uint(r.y);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&x)), b = __half22float2(*reinterpret_cast<const __half2*>(&y));
      return make_float4(a.x, a.y, b.x, b.y);
    };
    fetch(T - 1, pg, pc, pcp, pdo, pds);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = T - 1; s >= 0; --s) {
      const int t = dir ? (T - 1 - s) : s;
      const int it = T - 1 - s;
      uint8_t* Bn = cx.genB + (uint32_t)(it & 1) * S2_BB;
#pragma unroll
      for (int i = 0; i < S2_WPT; ++i) {
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + wq * S2_WPT + i < Bc) {
          const float4 g = gate_vals(pg[i]);
          const float dh = fmaf(pdo[i], pds[i], dh_rec[i]);
          const float tc = tanh_mufu(pc[i]);
          const float dct = fmaf(dh * g.w, 1.0f - tc * tc, dc[i]);
          dg.x = dct * g.z * g.x * (1.0f - g.x);
          dg.y = dct * pcp[i] * g.y * (1.0f - g.y);
          dg.z = dct * g.x * (1.0f - g.z * g.z);
          dg.w = dh * tc * g.w * (1.0f - g.w);
          dc[i] = dct * g.y;
          bsum.x += dg.x; bsum.y += dg.y; bsum.z += dg.z; bsum.w += dg.w;
          const long long row = (long long)t * Bc + b0 + wq * S2_WPT + i;
          *reinterpret_cast<float4*>(dG + row * ldg + colg) = dg;
        }
        if (s > 0) {   // k = gate*256 + unit: atom 4 gate + 2 r + u / 64
          *reinterpret_cast<__nv_bfloat16*>(Bn + 0 * 4 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.x);
          *reinterpret_cast<__nv_bfloat16*>(Bn + 1 * 4 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.y);
          *reinterpret_cast<__nv_bfloat16*>(Bn + 2 * 4 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.z);
          *reinterpret_cast<__nv_bfloat16*>(Bn + 3 * 4 * SW_BATOM + boff[i]) = __float2bfloat16_rn(dg.w);
        }
      }
      if (s == 0) {
        if (dbias) {
          atomicAdd(dbias + colg + 0, bsum.x); atomicAdd(dbias + colg + 1, bsum.y);
          atomicAdd(dbias + colg + 2, bsum.z); atomicAdd(dbias + colg + 3, bsum.w);
        }
        break;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if ((tid & 31) == 0) { jit(); mbar_arrive(cx.op_ready); }
      float4 ng[S2_WPT];
      float ncp[S2_WPT], ndo[S2_WPT], nds[S2_WPT];
      fetch(s - 1, ng, nullptr, ncp, ndo, nds);
      if (s >= 4) {
        const int t4 = dir ? (T - 1 - (s - 4)) : s - 4;
#pragma unroll
        for (int i = 0; i < S2_WPT; ++i) {
          const long long row = (long long)t4 * Bc + brow[i];
          if ((tid & 15) == 0) sw_prefetch_l2(reinterpret_cast<const __half*>(gates) + row * ldg + colg);
          if ((tid & 31) == 0) { sw_prefetch_l2(csave + row * D + colh); sw_prefetch_l2(dout + row * D + colh); }
        }
      }
      mbar_wait(cx.acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t a[S2_WPT];
      tmem_ld4(taddr, a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < S2_WPT; ++i) {
        dh_rec[i] = __uint_as_float(a[i]);
        pc[i] = pcp[i]; pg[i] = ng[i]; pcp[i] = ncp[i]; pdo[i] = ndo[i]; pds[i] = nds[i];
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp_u == SW_EPI / 32) {
    tc_fence_after();
    tmem_dealloc(cx.tmem, 512);
  }
}

static int s2_setup() {
  static PerDeviceFlag done_pd;
  bool& done = done_pd.cur();
  if (!done) {
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_swap256_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2_smem_bytes(S2_FB)));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_swap256, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2_smem_bytes(S2_BB)));
    done = true;
  }
  return BCI_OK;
}

bool rec_swap256_ok(int H, const void* G, int ldg) { return H == 256 && ((uintptr_t)G & 15) == 0 && (ldg & 3) == 0; }

int launch_rec_swap256_fwd(int ND, const float* G, int ldg, const __half* whh, float* out, float* gates, float* csave, int D, int Bc, int T,
                           cudaStream_t st, const SwapDropout* drop, bool g_half) {
  int rc = s2_setup();
  if (rc) return rc;
  const SwDrop dr = drop ? SwDrop{drop->outd, drop->outd_lo, drop->p, drop->seed, drop->site} : SwDrop{nullptr, nullptr, 0.f, 0, 0};
  lstm_rec_swap256_fwd<<<dim3(2 * ceil_div(Bc, S2_NW), ND), SW_BLOCK, s2_smem_bytes(S2_FB), st>>>(G, ldg, whh, out, gates, csave, dr, D, Bc, T, sw_jitter(), g_half ? 1 : 0);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int launch_bptt_swap256(int ND, const float* dout, const float* gates, const float* csave, const __nv_bfloat16* whhT, float* dG, float* dbias,
                        int ldg, int D, int Bc, int T, cudaStream_t st, const SwapDropout* drop) {
  int rc = s2_setup();
  if (rc) return rc;
  const SwDrop dr = drop ? SwDrop{nullptr, nullptr, drop->p, drop->seed, drop->site} : SwDrop{nullptr, nullptr, 0.f, 0, 0};
  lstm_bptt_swap256<<<dim3(2 * ceil_div(Bc, S2_NW), ND), SW_BLOCK, s2_smem_bytes(S2_BB), st>>>(dout, gates, csave, whhT, dG, dbias, dr, ldg, D, Bc, T, sw_jitter());
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---- probe: the A operand of tcgen05.mma read from TENSOR MEMORY ---------------------------------------------------------------
// One M128 x N16 x K16 product whose B is the 16 x 16 identity: D[m][n] = A[m][n] as the tensor core sees it.  A is written with
// tcgen05.st.32x32b.x8 (lane = row m): column c of the 8 carries the two halves value(c, 0) | value(c, 1) << 16 with
// value(c, h) = 1 + 2 c + h, so the result tells which (column, half) the hardware reads as K index n.
__global__ void __launch_bounds__(128, 1) tmem_a_probe_kernel(float* __restrict__ out) {
  __shared__ __align__(128) uint8_t sB_raw[2048 + 1024];
  uint8_t* sB = sB_raw + (((smem_u32(sB_raw) + 1023u) & ~1023u) - smem_u32(sB_raw));
  __shared__ __align__(8) uint64_t bar_s;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = smem_u32(&bar_s);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (tid < 32) { tmem_alloc(smem_u32(&slot), 64); tmem_relinquish(); }
  for (int i = tid; i < 2048 / 2; i += 128) reinterpret_cast<__half*>(sB)[i] = __float2half_rn(0.f);
  __syncthreads();
  if (tid < 16) {  // B[n][k] = (n == k): row n, k in chunk k/8 (swizzled), element k%8
    const uint32_t n = tid, k = tid;
    *reinterpret_cast<__half*>(sB + n * 128u + (((k >> 3) ^ (n & 7u)) << 4) + (k & 7u) * 2u) = __float2half_rn(1.f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const __half lo = __float2half_rn((float)(1 + 2 * c) + (tid == 5 ? 100.f : 0.f)), hi = __float2half_rn((float)(2 + 2 * c));
    v[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(lane_addr + 32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    constexpr uint32_t idesc = sw_idesc(128, 16, false);
    const uint64_t db = umma_desc_sw128(smem_u32(sB));
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem), "r"(tmem + 32), "l"(db), "r"(idesc), "r"(0u) : "memory");
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t d[8], d2[8];
  tmem_ld8(lane_addr, d);
  tmem_ld8(lane_addr + 8, d2);
  tmem_ld_wait();
#pragma unroll
  for (int n = 0; n < 8; ++n) { out[tid * 16 + n] = __uint_as_float(d[n]); out[tid * 16 + 8 + n] = __uint_as_float(d2[n]); }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

}  // namespace bci

// diagnostics (tests/test_gpu_rec_swap.py): the swapped recurrences in isolation, fp32 weights in the PyTorch layout
// scratch layout: fwd [ND][2][512][128] fp16 | bwd [ND][128][512] bf16 | bwd16 [ND][2][128][512] fp16  = 5 ND 65 536 sixteen-bit values
static int sw_selftest_pack(const float* w_hh, void* packed, int ND, __half** f, __nv_bfloat16** b, __half** b16, cudaStream_t st) {
  using namespace bci;
  const size_t W = 512 * 128;
  *f = reinterpret_cast<__half*>(packed);
  *b = reinterpret_cast<__nv_bfloat16*>(*f + (size_t)ND * 2 * W);
  *b16 = reinterpret_cast<__half*>(*b + (size_t)ND * W);
  for (int d = 0; d < ND; ++d) {
    int rc = pack_whh_swap(w_hh + d * W, *f + d * 2 * W, *b + d * W, *b16 + d * 2 * W, 128, st);
    if (rc) return rc;
  }
  return BCI_OK;
}
extern "C" int bci_selftest_rec_swap_fwd(const float* G, const float* w_hh, void* packed, float* out, float* gates, float* csave, int32_t Bc,
                                         int32_t T, int32_t ND, int32_t split, void* stream) {
  using namespace bci;
  BCI_REQUIRE(G && w_hh && packed && out && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL, "bci_selftest_rec_swap_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half *f, *b16;
  __nv_bfloat16* b;
  int rc = sw_selftest_pack(w_hh, packed, ND, &f, &b, &b16, st);
  if (rc) return rc;
  return launch_rec_swap_fwd(ND, G, ND * 512, f, out, gates, csave, ND * 128, Bc, T, split != 0, st);
}
extern "C" int bci_selftest_bptt_swap(const float* dout, const float* gates, const float* csave, const float* w_hh, void* packed, float* dG,
                                      int32_t Bc, int32_t T, int32_t ND, int32_t split, void* stream) {
  using namespace bci;
  BCI_REQUIRE(dout && gates && csave && w_hh && packed && dG && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL,
              "bci_selftest_bptt_swap: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half *f, *b16;
  __nv_bfloat16* b;
  int rc = sw_selftest_pack(w_hh, packed, ND, &f, &b, &b16, st);
  if (rc) return rc;
  return launch_bptt_swap(ND, dout, gates, csave, split ? (const void*)b16 : (const void*)b, dG, nullptr, nullptr, ND * 512, ND * 128, Bc, T, split != 0, st);
}
// H = 256 pair kernels in isolation: G / gates / dG [T*Bc][ND*1024], out / csave / dout [T][Bc][ND*256], w_hh [ND][1024][256] fp32;
// packed: 5 x ND x 1024 x 256 sixteen-bit values of scratch
static int s2_selftest_pack(const float* w_hh, void* packed, int ND, __half** f, __nv_bfloat16** b, cudaStream_t st) {
  using namespace bci;
  const size_t W = 1024 * 256;
  *f = reinterpret_cast<__half*>(packed);
  *b = reinterpret_cast<__nv_bfloat16*>(*f + (size_t)ND * 2 * W);
  __half* b16 = reinterpret_cast<__half*>(*b + (size_t)ND * W);
  for (int d = 0; d < ND; ++d) {
    int rc = pack_whh_swap(w_hh + d * W, *f + d * 2 * W, *b + d * W, b16 + d * 2 * W, 256, st);
    if (rc) return rc;
  }
  return BCI_OK;
}
extern "C" int bci_selftest_rec_swap256_fwd(const float* G, const float* w_hh, void* packed, float* out, float* gates, float* csave,
                                            int32_t Bc, int32_t T, int32_t ND, void* stream) {
  using namespace bci;
  BCI_REQUIRE(G && w_hh && packed && out && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL, "bci_selftest_rec_swap256_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half* f;
  __nv_bfloat16* b;
  int rc = s2_selftest_pack(w_hh, packed, ND, &f, &b, st);
  if (rc) return rc;
  return launch_rec_swap256_fwd(ND, G, ND * 1024, f, out, gates, csave, ND * 256, Bc, T, st);
}
extern "C" int bci_selftest_bptt_swap256(const float* dout, const float* gates, const float* csave, const float* w_hh, void* packed,
                                         float* dG, int32_t Bc, int32_t T, int32_t ND, void* stream) {
  using namespace bci;
  BCI_REQUIRE(dout && gates && csave && w_hh && packed && dG && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL,
              "bci_selftest_bptt_swap256: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half* f;
  __nv_bfloat16* b;
  int rc = s2_selftest_pack(w_hh, packed, ND, &f, &b, st);
  if (rc) return rc;
  return launch_bptt_swap256(ND, dout, gates, csave, b, dG, nullptr, ND * 1024, ND * 256, Bc, T, st);
}
/* selftest only: clock64 stamps (8 per step, steps 100-103) of CTA (0,0) of the next forward launches; NULL switches them off */
extern "C" int bci_selftest_swap_set_debug(long long* stamps) {
  bci::g_sw_dbg = stamps;
  return BCI_OK;
}
extern "C" int bci_selftest_tmem_a_probe(float* out, void* stream) {
  using namespace bci;
  BCI_REQUIRE(out, BCI_EINVAL, "bci_selftest_tmem_a_probe: out is NULL");
  tmem_a_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(out);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
