// Training-size recurrences on the tensor cores: the "swapped" form  pre^T = W_hh . h^T  with W_hh RESIDENT as the MMA's A operand.
//
// A 512-window training batch is 4 tiles of 128 windows: the tile-per-CTA tensor-core recurrences of the inference paths
// (lstm_bf16_fused.cu, lstm_fp32_tc.cu) would keep 8 SMs busy, which is why the training step of round 1 ran its recurrences on
// the CUDA cores (thread = (unit, 4 windows), 128 CTAs: 4.9 us per forward step, 7 us per BPTT step; 62 % of the step).  Swapping
// the operands makes the WINDOWS the MMA's N dimension, which may be as small as 16:
//
//   forward   D_g[u][n] = sum_k W_hh[g*128 + u][k] . h_{t-1}[n][k]      four M128 x N16 x K128 products (one per gate) -> 64 TMEM columns
//   BPTT      D[j][n]   = sum_k W_hh^T[j][k] . dG_t[n][k],  k = (gate, unit)    one M128 x N16 x K512 product            -> 16 TMEM columns
//
//   A = the weights, 128 KB of 16-bit values, K-major SWIZZLE_128B atoms, loaded ONCE per CTA (the PyTorch (4H, H) row order is
//       already "gate block g, row u": the forward operand is a plain fp16 cast; BPTT's is the transpose in bf16)
//   B = h_{t-1} (4 KB) / dG_t (16 KB): rows = windows, written by the epilogue threads of the previous step
//   D = TMEM lane u (hidden unit) x column n (window): thread u of the CTA owns unit u for the CTA's 8 windows, reads its four
//       gate pre-activations with four tcgen05.ld.32x32b.x8 -- the whole cell update is thread-local, no exchange of any kind.
//
// One CTA = 8 windows x one direction (N = 16 is the smallest legal N at M = 128; rows 8-15 of B stay zero), 128 threads, one CTA
// per SM: 512 windows x 2 directions = 128 CTAs.  Per step: thread 0 issues 32 MMAs (4 KB of A each: the product is bound by the
// shared-memory read of the weights, ~1 000 cycles) and commits to an mbarrier; everybody prefetches the next step's G_t / saved
// activations (coalesced: consecutive threads = consecutive units) while the product runs.
//
// Precision ("mixed" training mode, the analogue of the reference's autocast training, 04_lstm_model.py:486-490): forward operands
// fp16 (h in [-1, 1], 11 significant bits), BPTT operands bf16 (gradients need the exponent range, not the bits), accumulation,
// gates, cell state, dG in fp32.  The fp32-parity training step keeps the CUDA-core recurrences.
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int SW_NW = 8;                     // windows per CTA
constexpr int SW_THREADS = 128;              // thread = hidden unit
constexpr uint32_t SW_AATOM = 128 * 128;     // A atom: 128 rows x 64 halves
constexpr uint32_t SW_BATOM = 16 * 128;      // B atom: 16 rows x 64 halves (rows 8-15 zero)
constexpr uint32_t SW_A_BYTES = 8 * SW_AATOM;             // forward: [gate 4][K atom 2]; BPTT: [K atom 8]
constexpr uint32_t SW_FWD_B = 2 * SW_BATOM, SW_BWD_B = 8 * SW_BATOM;
constexpr size_t SW_FWD_SMEM = 1024 + SW_A_BYTES + SW_FWD_B + 64;
constexpr size_t SW_BWD_SMEM = 1024 + SW_A_BYTES + SW_BWD_B + 64;

__host__ __device__ constexpr uint32_t sw_idesc(int M, int N, bool bf16) {
  return (1u << 4) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}

// ---- operand packing --------------------------------------------------------------------------------------------------------
// w_hh (4H, H) fp32 -> fp16, same order
__global__ void pack_whh_swap_fwd_kernel(const float* __restrict__ w, __half* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(w[i]);
}
// w_hh (4H, H) fp32 -> dst [j][k = gate*H + unit] bf16 = w_hh[k][j]
__global__ void pack_whh_swap_bwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H * H) return;
  const int j = i / (4 * H), k = i - j * 4 * H;
  dst[i] = __float2bfloat16_rn(w[(size_t)k * H + j]);
}
int pack_whh_swap(const float* w_hh, __half* fwd, __nv_bfloat16* bwd, int H, cudaStream_t st) {
  pack_whh_swap_fwd_kernel<<<ceil_div(4 * H * H, 256), 256, 0, st>>>(w_hh, fwd, 4 * H * H);
  BCI_LAUNCH_OK();
  pack_whh_swap_bwd_kernel<<<ceil_div(4 * H * H, 256), 256, 0, st>>>(w_hh, bwd, H);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// common prologue: 1024-aligned dynamic shared memory, one mbarrier, TMEM columns
struct SwCtx {
  uint32_t base, sA, sB, bar, tmem;
  uint8_t* gen;
};
template <uint32_t B_BYTES, uint32_t TMEM_COLS>
__device__ __forceinline__ SwCtx sw_prologue(uint8_t* raw_ptr, const uint4* __restrict__ wsrc, int chunks_per_row) {
  SwCtx c;
  const uint32_t raw = smem_u32(raw_ptr);
  c.base = (raw + 1023u) & ~1023u;
  c.gen = raw_ptr + (c.base - raw);
  c.sA = c.base;
  c.sB = c.base + SW_A_BYTES;
  uint8_t* ctl = c.gen + SW_A_BYTES + B_BYTES;
  c.bar = smem_u32(ctl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 16);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(c.bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  // weights: global row-major [block][128 rows][chunks_per_row x 8 halves] (128 KB) -> atoms [block][K atom][row][64], SW128
  const int katoms = chunks_per_row >> 3;
  for (int i = tid; i < (int)(SW_A_BYTES / 16); i += SW_THREADS) {
    const int row_g = i / chunks_per_row, cc = i - row_g * chunks_per_row;
    const int blk = row_g >> 7, row = row_g & 127;
    const uint4 v = __ldg(wsrc + i);
    *reinterpret_cast<uint4*>(c.gen + (uint32_t)(blk * katoms + (cc >> 3)) * SW_AATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) = v;
  }
  for (int i = tid; i < (int)(B_BYTES / 16); i += SW_THREADS) reinterpret_cast<uint4*>(c.gen + SW_A_BYTES)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = *tmem_slot;
  return c;
}

// ---- forward --------------------------------------------------------------------------------------------------------------------
// grid = (ceil(Bc / 8), ND)
__global__ void __launch_bounds__(SW_THREADS, 1)
lstm_rec_swap_fwd(const float* __restrict__ G,        // [T*Bc][ldg] fp32: column dir*512 + unit*4 + gate, bias included
                  int ldg,
                  const __half* __restrict__ whh,     // [ND][512][128] fp16, PyTorch row order
                  float* __restrict__ out,            // [T][Bc][D]: h_t at column dir*128 + unit
                  float* __restrict__ gates,          // optional [T*Bc][ldg] gate ACTIVATIONS, same layout as G
                  float* __restrict__ csave,          // optional [T*Bc][D] cell states
                  int D, int Bc, int T) {
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, u = tid;
  const int dir = blockIdx.y, b0 = blockIdx.x * SW_NW;
  const SwCtx cx = sw_prologue<SW_FWD_B, 64>(sw_smem_raw, reinterpret_cast<const uint4*>(whh + (size_t)dir * 512 * 128), 16);
  const uint32_t taddr = cx.tmem + ((uint32_t)(warp * 32) << 16);

  // this thread's element of row n of the B tile: k = u
  uint32_t hoff[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n)
    hoff[n] = (uint32_t)(u >> 6) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)n) << 4) + (uint32_t)(u & 7) * 2u;
  uint8_t* genB = cx.gen + SW_A_BYTES;
  int brow[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n) brow[n] = b0 + n < Bc ? b0 + n : Bc - 1;   // dead windows read a valid row and store nothing
  const int colg = dir * 512 + u * 4, colh = dir * 128 + u;

  float c[SW_NW];
  float4 gq[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n) c[n] = 0.f;
  {
    const int t0 = dir ? T - 1 : 0;
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) gq[n] = __ldg(reinterpret_cast<const float4*>(G + ((long long)t0 * Bc + brow[n]) * ldg + colg));
  }
  for (int st = 0; st < T; ++st) {
    const int t = dir ? (T - 1 - st) : st;
    if (tid == 0) {
      constexpr uint32_t idesc = sw_idesc(128, 16, false);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)(g * 2 + (k >> 2)) * SW_AATOM + (uint32_t)(k & 3) * 32u);
          const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
          umma_bf16(cx.tmem + g * 16, da, db, idesc, k != 0 ? 1u : 0u);
        }
      }
      umma_commit(cx.bar);
    }
    float4 gc[SW_NW];
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) gc[n] = gq[n];
    if (st + 1 < T) {
      const int tn = dir ? (T - 2 - st) : st + 1;
#pragma unroll
      for (int n = 0; n < SW_NW; ++n) gq[n] = __ldg(reinterpret_cast<const float4*>(G + ((long long)tn * Bc + brow[n]) * ldg + colg));
    }
    mbar_wait(cx.bar, (uint32_t)(st & 1));
    tc_fence_after();
    uint32_t a[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_ld8(taddr + g * 16, a[g]);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) {
      const float ig = rec_sigmoid(__uint_as_float(a[0][n]) + gc[n].x);
      const float fg = rec_sigmoid(__uint_as_float(a[1][n]) + gc[n].y);
      const float gg = rec_tanh(__uint_as_float(a[2][n]) + gc[n].z);
      const float og = rec_sigmoid(__uint_as_float(a[3][n]) + gc[n].w);
      c[n] = fmaf(fg, c[n], ig * gg);
      const float hv = og * rec_tanh(c[n]);
      if (b0 + n < Bc) {
        const long long row = (long long)t * Bc + b0 + n;
        out[row * D + colh] = hv;
        if (gates) *reinterpret_cast<float4*>(gates + row * ldg + colg) = make_float4(ig, fg, gg, og);
        if (csave) csave[row * D + colh] = c[n];
      }
      *reinterpret_cast<__half*>(genB + hoff[n]) = __float2half_rn(hv);
    }
    fence_proxy_async_smem();  // h_t (generic-proxy stores) -> visible to the next step's tcgen05.mma
    tc_fence_before();         // this thread's TMEM reads are ordered before the barrier
    __syncthreads();
    tc_fence_after();
  }
  if (tid < 32) tmem_dealloc(cx.tmem, 64);
}

// ---- BPTT -----------------------------------------------------------------------------------------------------------------------
// The mirror of lstm_bptt_f32 (lstm_train.cu): walks the direction's time order backwards, dG_t to global (fp32, + optional tf32
// remainder for the split-precision GEMMs) and, as bf16, into the B tile of  dh_{t-1}[j] = sum_k dG_t[k] W_hh[k][j].
__global__ void __launch_bounds__(SW_THREADS, 1)
lstm_bptt_swap(const float* __restrict__ dout,           // [T][Bc][D]
               const float* __restrict__ gates,          // [T*Bc][ldg]: i,f,g,o of (dir, unit) at column dir*512 + unit*4
               const float* __restrict__ csave,          // [T*Bc][D]
               const __nv_bfloat16* __restrict__ whhT,   // [ND][128 j][512 k = gate*128 + unit] bf16
               float* __restrict__ dG,                   // [T*Bc][ldg]
               float* __restrict__ dG_lo,                // optional
               int ldg, int D, int Bc, int T) {
  extern __shared__ uint8_t sw_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, u = tid;
  const int dir = blockIdx.y, b0 = blockIdx.x * SW_NW;
  const SwCtx cx = sw_prologue<SW_BWD_B, 32>(sw_smem_raw, reinterpret_cast<const uint4*>(whhT + (size_t)dir * 512 * 128), 64);
  const uint32_t taddr = cx.tmem + ((uint32_t)(warp * 32) << 16);
  uint8_t* genB = cx.gen + SW_A_BYTES;
  // element (row n, k = gate*128 + u) of the B tile: atom gate*2 + u/64
  uint32_t boff[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n)
    boff[n] = (uint32_t)(u >> 6) * SW_BATOM + (uint32_t)n * 128u + (((uint32_t)((u & 63) >> 3) ^ (uint32_t)n) << 4) + (uint32_t)(u & 7) * 2u;
  int brow[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n) brow[n] = b0 + n < Bc ? b0 + n : Bc - 1;
  const int colg = dir * 512 + u * 4, colh = dir * 128 + u;

  float dh_rec[SW_NW], dc[SW_NW];
#pragma unroll
  for (int n = 0; n < SW_NW; ++n) { dh_rec[n] = 0.f; dc[n] = 0.f; }
  float4 pg[SW_NW];
  float pc[SW_NW], pcp[SW_NW], pdo[SW_NW];
  auto fetch = [&](int s, float4* g4, float* cc, float* cp, float* dd) {
    const int t = dir ? (T - 1 - s) : s;
    const int tp = dir ? (t + 1) : (t - 1);
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) {
      const long long row = (long long)t * Bc + brow[n];
      g4[n] = __ldg(reinterpret_cast<const float4*>(gates + row * ldg + colg));
      if (cc) cc[n] = __ldg(csave + row * D + colh);
      cp[n] = (s > 0) ? __ldg(csave + ((long long)tp * Bc + brow[n]) * D + colh) : 0.f;
      dd[n] = __ldg(dout + row * D + colh);
    }
  };
  fetch(T - 1, pg, pc, pcp, pdo);
  for (int s = T - 1; s >= 0; --s) {
    const int t = dir ? (T - 1 - s) : s;
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) {
      float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + n < Bc) {
        const float4 g = pg[n];
        const float dh = pdo[n] + dh_rec[n];
        const float tc = rec_tanh(pc[n]);  // the forward's own tanh(c)
        const float dct = fmaf(dh * g.w, 1.0f - tc * tc, dc[n]);
        dg.x = dct * g.z * g.x * (1.0f - g.x);
        dg.y = dct * pcp[n] * g.y * (1.0f - g.y);
        dg.z = dct * g.x * (1.0f - g.z * g.z);
        dg.w = dh * tc * g.w * (1.0f - g.w);
        dc[n] = dct * g.y;
        const long long row = (long long)t * Bc + b0 + n;
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
      if (s > 0) {
        *reinterpret_cast<__nv_bfloat16*>(genB + 0 * 2 * SW_BATOM + boff[n]) = __float2bfloat16_rn(dg.x);
        *reinterpret_cast<__nv_bfloat16*>(genB + 1 * 2 * SW_BATOM + boff[n]) = __float2bfloat16_rn(dg.y);
        *reinterpret_cast<__nv_bfloat16*>(genB + 2 * 2 * SW_BATOM + boff[n]) = __float2bfloat16_rn(dg.z);
        *reinterpret_cast<__nv_bfloat16*>(genB + 3 * 2 * SW_BATOM + boff[n]) = __float2bfloat16_rn(dg.w);
      }
    }
    if (s == 0) break;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      constexpr uint32_t idesc = sw_idesc(128, 16, true);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const uint64_t da = umma_desc_sw128(cx.sA + (uint32_t)(k >> 2) * SW_AATOM + (uint32_t)(k & 3) * 32u);
        const uint64_t db = umma_desc_sw128(cx.sB + (uint32_t)(k >> 2) * SW_BATOM + (uint32_t)(k & 3) * 32u);
        umma_bf16(cx.tmem, da, db, idesc, k != 0 ? 1u : 0u);
      }
      umma_commit(cx.bar);
    }
    // step s-1's saved activations are requested while the product runs; c(s-1) is this step's cprev
    float4 ng[SW_NW];
    float ncp[SW_NW], ndo[SW_NW];
    fetch(s - 1, ng, nullptr, ncp, ndo);
    mbar_wait(cx.bar, (uint32_t)((T - 1 - s) & 1));
    tc_fence_after();
    uint32_t a[8];
    tmem_ld8(taddr, a);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < SW_NW; ++n) {
      dh_rec[n] = __uint_as_float(a[n]);
      pc[n] = pcp[n]; pg[n] = ng[n]; pcp[n] = ncp[n]; pdo[n] = ndo[n];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(cx.tmem, 32);
}

static int sw_setup() {
  static PerDeviceFlag done_pd;
  bool& done = done_pd.cur();
  if (!done) {
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_swap_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_FWD_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_swap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_BWD_SMEM));
    done = true;
  }
  return BCI_OK;
}

bool rec_swap_ok(int H, const void* G, int ldg) { return H == 128 && ((uintptr_t)G & 15) == 0 && (ldg & 3) == 0; }

int launch_rec_swap_fwd(int ND, const float* G, int ldg, const __half* whh, float* out, float* gates, float* csave, int D, int Bc, int T,
                        cudaStream_t st) {
  int rc = sw_setup();
  if (rc) return rc;
  lstm_rec_swap_fwd<<<dim3(ceil_div(Bc, SW_NW), ND), SW_THREADS, SW_FWD_SMEM, st>>>(G, ldg, whh, out, gates, csave, D, Bc, T);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int launch_bptt_swap(int ND, const float* dout, const float* gates, const float* csave, const __nv_bfloat16* whhT, float* dG, float* dG_lo,
                     int ldg, int D, int Bc, int T, cudaStream_t st) {
  int rc = sw_setup();
  if (rc) return rc;
  lstm_bptt_swap<<<dim3(ceil_div(Bc, SW_NW), ND), SW_THREADS, SW_BWD_SMEM, st>>>(dout, gates, csave, whhT, dG, dG_lo, ldg, D, Bc, T);
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
extern "C" int bci_selftest_rec_swap_fwd(const float* G, const float* w_hh, void* packed, float* out, float* gates, float* csave, int32_t Bc,
                                         int32_t T, int32_t ND, void* stream) {
  using namespace bci;
  BCI_REQUIRE(G && w_hh && packed && out && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL, "bci_selftest_rec_swap_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half* f = reinterpret_cast<__half*>(packed);
  __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(f + (size_t)ND * 512 * 128);
  for (int d = 0; d < ND; ++d) {
    int rc = pack_whh_swap(w_hh + (size_t)d * 512 * 128, f + (size_t)d * 512 * 128, b + (size_t)d * 512 * 128, 128, st);
    if (rc) return rc;
  }
  return launch_rec_swap_fwd(ND, G, ND * 512, f, out, gates, csave, ND * 128, Bc, T, st);
}
extern "C" int bci_selftest_bptt_swap(const float* dout, const float* gates, const float* csave, const float* w_hh, void* packed, float* dG,
                                      int32_t Bc, int32_t T, int32_t ND, void* stream) {
  using namespace bci;
  BCI_REQUIRE(dout && gates && csave && w_hh && packed && dG && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL,
              "bci_selftest_bptt_swap: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __half* f = reinterpret_cast<__half*>(packed);
  __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(f + (size_t)ND * 512 * 128);
  for (int d = 0; d < ND; ++d) {
    int rc = pack_whh_swap(w_hh + (size_t)d * 512 * 128, f + (size_t)d * 512 * 128, b + (size_t)d * 512 * 128, 128, st);
    if (rc) return rc;
  }
  return launch_bptt_swap(ND, dout, gates, csave, b, dG, nullptr, ND * 512, ND * 128, Bc, T, st);
}
extern "C" int bci_selftest_tmem_a_probe(float* out, void* stream) {
  using namespace bci;
  BCI_REQUIRE(out, BCI_EINVAL, "bci_selftest_tmem_a_probe: out is NULL");
  tmem_a_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(out);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
