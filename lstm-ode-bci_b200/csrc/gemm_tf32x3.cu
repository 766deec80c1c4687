// Split-precision (3 x TF32) tcgen05 GEMMs for the fp32 parity path.
//
// The fp32 mode must stay within 1e-5 of the reference's fp32 CPU arithmetic (04_lstm_model.py:181-188 through torch's
// CPU kernels), which plain TF32 (10-bit mantissa) cannot do.  Every fp32 operand x is therefore written as
//     x = hi + lo,   hi = tf32(x) (the tensor core reads the upper 19 bits),  lo = tf32(x - hi)
// and the product is accumulated in fp32 in TMEM as  A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  (the dropped lo.lo term is
// 2^-22 relative): three kind::tf32 MMAs per product instead of CUDA-core FFMAs -- ~370 TFLOP/s of fp32-grade math per
// GPU against 71.6 TFLOP/s on the FMA pipe.  The lo parts are produced by one elementwise pass (split_tf32_kernel; weights:
// once per load_weights) so that the GEMM itself is a pure TMA -> tcgen05 pipeline.
//
//   NT  C[M][N] (=|+=) A[M][K] . W[N][K]^T (+ bias)      time-parallel projections G = in . W_ih^T (forward) and data
//                                                        gradients d_in = dG . W_ih (backward, W given as [Kin][8H])
//   TN  C[P][Q]  =     sum_r A[r][P] . B[r][Q]           weight gradients dW = dG^T . in over the R = T*Bc rows; operands
//                                                        are read as MN-major UMMA tiles straight from the row-major
//                                                        activations (no transpose pass), split-K over CTAs, partial
//                                                        tiles combined in L2 by the TMA reduce-add store
//
// F16 = true (NT only): the same pipeline with every operand split into two FP16 numbers instead of two TF32 numbers
// (x = hi + lo, hi = fp16(x), lo = fp16(x - hi): 22 significant bits as long as lo stays out of fp16's subnormals -- activations
// of the LSTM are O(1), the weights are pre-scaled by 16 and the epilogue multiplies by 1/16).  kind::f16 runs at twice the rate
// of kind::tf32 and the operand tiles hold 64 instead of 32 K values in the same 16 KB, so the forward projections of the fp32
// inference path (G = in . W_ih^T) cost about half; fp16 x fp16 products are exact in the fp32 accumulator.  Not used for
// gradients (their magnitudes leave fp16's range).
//
// Kernel shape (both): persistent CTAs, warp 0 = TMA producer (3-stage ring of {A_hi, A_lo, B_hi, B_lo} 128x32 fp32 tiles,
// SWIZZLE_128B), warp 1 = MMA issuer (M128 x N128 x K8; two accumulators per tile -- hi.hi and the
// correction terms -- double-buffered: all 512 TMEM columns), warps 2-9 = epilogue (two per TMEM lane quarter, two 32-column slabs each)
// (tcgen05.ld -> +bias -> swizzled staging -> TMA store / reduce-add, one 32x32 box per warp).
#include "lstm_handle.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int TX_BM = 128, TX_BN = 128, TX_BK = 32, TX_STAGES = 3;
constexpr int TX_THREADS = 320;                     // TMA warp, MMA warp, 8 epilogue warps
constexpr uint32_t TX_TILE = TX_BM * TX_BK * 4;     // 16 KB: one operand tile
constexpr uint32_t TX_STAGE = 4 * TX_TILE;          // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t TX_CSTAGE = 8 * 32 * 128;        // per epilogue warp: 32 rows x 128 B
constexpr size_t TX_SMEM = 1024 + (size_t)TX_STAGES * TX_STAGE + TX_CSTAGE + 128 * sizeof(float) + 256;

// x -> lo (and optionally hi).  hi == nullptr: the tensor core is trusted to ignore the low 13 bits of the raw operand,
// so lo is the remainder after truncation; otherwise hi is the round-to-nearest tf32 value and lo its remainder.
__global__ void split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(x + i);
  float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t hb;
    if (hi) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[j]));
    else hb = __float_as_uint(in[j]) & 0xFFFFE000u;
    h[j] = __uint_as_float(hb);
    uint32_t lb;
    const float rem = in[j] - h[j];  // exact
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
    l[j] = __uint_as_float(lb);
  }
  if (hi) hi[i] = make_float4(h[0], h[1], h[2], h[3]);
  lo[i] = make_float4(l[0], l[1], l[2], l[3]);
}

int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st) {
  BCI_REQUIRE(n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)hi & 15) == 0, BCI_EINVAL,
              "split_tf32: 16-byte aligned arrays with n %% 4 == 0 required");
  if (n == 0) return BCI_OK;
  split_tf32_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                                                       reinterpret_cast<float4*>(lo), n / 4);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// Instruction descriptor, kind::tf32: D fp32, A/B tf32 (format 2); bit 15 / 16 = A / B is MN-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(mn_major ? 1 : 0) << 15) | ((uint32_t)(mn_major ? 1 : 0) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand tile: [MN group of 32 floats][k row][128 B].  32-bit MN-major operands exist only in the
// "128-byte swizzle with a 32-byte base" layout (UMMA layout type 1 = TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: the 32-byte
// chunk index of a row is XORed with the row index mod 4; plain SWIZZLE_128B is silently read as zeros): atoms of 4 k-rows
// (512 B, SBO), MN groups 4096 B apart (LBO); one K = 8 MMA spans two atoms = 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}

struct TxMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo, c; };

// TN = false: tile (mb, nb) of C[M][N], K loop over [0, K) (k_splits == 1) -- operands K-major, 2-D maps {k, row}
// TN = true : tile (mb, nb) of C[P=M][Q=N], K loop over the R rows in k_splits ranges -- operands MN-major (four {32 floats, 32 rows}
//             boxes per tile); partial tiles are reduce-added
__device__ __forceinline__ void umma_f16_1sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16 with FP16 inputs (a_format = b_format = 0), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16_k(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool TN, bool F16>
__global__ void __launch_bounds__(TX_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ TxMaps maps, const float* __restrict__ bias, int M, int N, long long K, int k_splits,
                   int reduce_add, float out_scale, int single) {   // single: plain one-pass product of the hi operands (mixed training mode)
  static_assert(!(TN && F16), "the fp16 split is built for the NT projections only");
  constexpr int BK = F16 ? 2 * TX_BK : TX_BK;   // K values per stage: 128-byte rows of fp16 / fp32
  extern __shared__ uint8_t tx_smem_raw[];
  const uint32_t raw = smem_u32(tx_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = tx_smem_raw + (base - raw);
  const uint32_t sRing = base, sC = base + TX_STAGES * TX_STAGE;
  uint8_t* genC = gen + TX_STAGES * TX_STAGE;
  float* bias_s = reinterpret_cast<float*>(genC + TX_CSTAGE);
  uint8_t* ctl = genC + TX_CSTAGE + 128 * sizeof(float);
  const uint32_t bar0 = smem_u32(ctl);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  // single-pass mode loads two tiles per k-block instead of four: every ring stage then holds TWO k-blocks (A in the hi / lo slot of
  // the A pair, B likewise), i.e. a ring of 6.  With 4 MMAs (256 tensor cycles) per k-block, three slots covered 770 cycles of load
  // latency -- less than an L2-miss TMA round trip: the mainloop waited on loads
  const int nslots = single ? 2 * TX_STAGES : TX_STAGES;
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 * TX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (4 * TX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (4 * TX_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * (4 * TX_STAGES + 4));
  // slot -> shared-memory offset of its A tile (its B tile is 2 * TX_TILE further)
  auto slot_base = [&](int s) { return single ? (uint32_t)(s >> 1) * TX_STAGE + (uint32_t)(s & 1) * TX_TILE : (uint32_t)s * TX_STAGE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (M + TX_BM - 1) / TX_BM, n_blocks = (N + TX_BN - 1) / TX_BN;
  const long long kb_total = (K + BK - 1) / BK;
  const long long kb_per = (kb_total + k_splits - 1) / k_splits;
  const long long tiles = (long long)m_blocks * n_blocks * k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_hi); tma_prefetch_desc(&maps.a_lo);
    tma_prefetch_desc(&maps.b_hi); tma_prefetch_desc(&maps.b_lo);
    tma_prefetch_desc(&maps.c);
    for (int s = 0; s < 2 * TX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 4 * TX_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int nb = (int)(t % n_blocks);
        const int mb = (int)((t / n_blocks) % m_blocks);
        const int ks = (int)(t / ((long long)n_blocks * m_blocks));
        const long long kb0 = ks * kb_per, kb1 = (kb0 + kb_per < kb_total) ? kb0 + kb_per : kb_total;
        for (long long kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), single ? TX_STAGE / 2 : TX_STAGE);
          const uint32_t s0 = sRing + slot_base(stage);
          if (single) {
            if (TN) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                tma_load_2d(s0 + g * 4096, &maps.a_hi, mb * TX_BM + g * 32, (int)(kb * TX_BK), full_bar(stage));
                tma_load_2d(s0 + 2 * TX_TILE + g * 4096, &maps.b_hi, nb * TX_BN + g * 32, (int)(kb * TX_BK), full_bar(stage));
              }
            } else {
              tma_load_2d(s0, &maps.a_hi, (int)(kb * BK), mb * TX_BM, full_bar(stage));
              tma_load_2d(s0 + 2 * TX_TILE, &maps.b_hi, (int)(kb * BK), nb * TX_BN, full_bar(stage));
            }
          } else if (TN) {
            // MN-major tile = four {32 floats, 32 rows} boxes side by side: [group][k row][128 B]
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              tma_load_2d(s0 + g * 4096, &maps.a_hi, mb * TX_BM + g * 32, (int)(kb * TX_BK), full_bar(stage));
              tma_load_2d(s0 + TX_TILE + g * 4096, &maps.a_lo, mb * TX_BM + g * 32, (int)(kb * TX_BK), full_bar(stage));
              tma_load_2d(s0 + 2 * TX_TILE + g * 4096, &maps.b_hi, nb * TX_BN + g * 32, (int)(kb * TX_BK), full_bar(stage));
              tma_load_2d(s0 + 3 * TX_TILE + g * 4096, &maps.b_lo, nb * TX_BN + g * 32, (int)(kb * TX_BK), full_bar(stage));
            }
          } else {
            tma_load_2d(s0, &maps.a_hi, (int)(kb * BK), mb * TX_BM, full_bar(stage));
            tma_load_2d(s0 + TX_TILE, &maps.a_lo, (int)(kb * BK), mb * TX_BM, full_bar(stage));
            tma_load_2d(s0 + 2 * TX_TILE, &maps.b_hi, (int)(kb * BK), nb * TX_BN, full_bar(stage));
            tma_load_2d(s0 + 3 * TX_TILE, &maps.b_lo, (int)(kb * BK), nb * TX_BN, full_bar(stage));
          }
          if (++stage == nslots) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = F16 ? umma_idesc_f16_k(TX_BM, TX_BN) : umma_idesc_tf32(TX_BM, TX_BN, TN ? 1 : 0);
      constexpr uint32_t KSTEP = TN ? 1024u : 32u;  // 8 k-rows of an MN-major tile / 8 floats inside a K-major row
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ks = (int)(t / ((long long)n_blocks * m_blocks));
        const long long kb0 = ks * kb_per, kb1 = (kb0 + kb_per < kb_total) ? kb0 + kb_per : kb_total;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        // two accumulators per tile: the tensor core adds into TMEM with truncation, an error that grows with the number of
        // accumulation steps times the accumulator's magnitude -- the small correction terms are summed apart from the
        // hi.hi products (a third of the steps on the big accumulator) and the epilogue adds the two in fp32 registers
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * (2 * TX_BN), d_lo = d_tmem + TX_BN;
        for (long long kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t s0 = sRing + slot_base(stage);
#pragma unroll
          for (int kk = 0; kk < TX_BK / 8; ++kk) {
            const uint32_t o = kk * KSTEP;
            const uint64_t ah = TN ? umma_desc_sw128_mn(s0 + o) : umma_desc_sw128(s0 + o);
            const uint64_t al = TN ? umma_desc_sw128_mn(s0 + TX_TILE + o) : umma_desc_sw128(s0 + TX_TILE + o);
            const uint64_t bh = TN ? umma_desc_sw128_mn(s0 + 2 * TX_TILE + o) : umma_desc_sw128(s0 + 2 * TX_TILE + o);
            const uint64_t bl = TN ? umma_desc_sw128_mn(s0 + 3 * TX_TILE + o) : umma_desc_sw128(s0 + 3 * TX_TILE + o);
            const uint32_t first = (kb != kb0 || kk != 0) ? 1u : 0u;
            if (single) {
              if (F16) umma_f16_1sm(d_tmem, ah, bh, idesc, first);
              else umma_tf32(d_tmem, ah, bh, idesc, first);
            } else if (F16) {
              umma_f16_1sm(d_lo, al, bh, idesc, first);
              umma_f16_1sm(d_lo, ah, bl, idesc, 1u);
              umma_f16_1sm(d_tmem, ah, bh, idesc, first);
            } else {
              umma_tf32(d_lo, al, bh, idesc, first);
              umma_tf32(d_lo, ah, bl, idesc, 1u);
              umma_tf32(d_tmem, ah, bh, idesc, first);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == nslots) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // epilogue warp w: TMEM lanes / tile rows [32 (w % 4), +32), two of the four 32-column slabs (two warps per lane quarter: with
    // four warps the single-pass products were bound by this epilogue -- ~7 000 cycles per tile against 2 048 of MMA time), each slab
    // staged as a swizzled 32 x 128 B box
    const int quarter = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    uint8_t* cst = genC + (warp - 2) * 4096;
    const uint32_t cst_s = sC + (warp - 2) * 4096;
    const int et = (warp - 2) * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int nb = (int)(t % n_blocks);
      const int mb = (int)((t / n_blocks) % m_blocks);
      const int ks = (int)(t / ((long long)n_blocks * m_blocks));
      const bool add_bias = bias != nullptr && ks == 0;
      // all eight epilogue warps: previous tile's bias reads are done before it is overwritten
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et < TX_BN) bias_s[et] = (add_bias && nb * TX_BN + et < N) ? __ldg(bias + nb * TX_BN + et) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * (2 * TX_BN);
#pragma unroll 1
      for (int slab = 2 * ehalf; slab < 2 * ehalf + 2; ++slab) {
        uint32_t r[32], rl[32];
        tmem_ld32(taddr + slab * 32, r);
        if (single) {
#pragma unroll
          for (int q = 0; q < 32; ++q) rl[q] = 0u;
        } else {
          tmem_ld32(taddr + TX_BN + slab * 32, rl);
        }
        if (lane == 0) tma_store_wait_read();  // the staging box has been drained by the previous store
        __syncwarp();
        tmem_ld_wait();
        if (slab == 2 * ehalf + 1) {  // all TMEM reads of this accumulator by this thread are done
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v;
          v.x = fmaf(__uint_as_float(r[4 * q + 0]) + __uint_as_float(rl[4 * q + 0]), out_scale, bias_s[slab * 32 + 4 * q + 0]);
          v.y = fmaf(__uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1]), out_scale, bias_s[slab * 32 + 4 * q + 1]);
          v.z = fmaf(__uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2]), out_scale, bias_s[slab * 32 + 4 * q + 2]);
          v.w = fmaf(__uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3]), out_scale, bias_s[slab * 32 + 4 * q + 3]);
          *reinterpret_cast<float4*>(cst + sw128_chunk_off((uint32_t)lane, (uint32_t)q)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int c0 = nb * TX_BN + slab * 32, c1 = mb * TX_BM + quarter * 32;
          if (c0 < N && c1 < M) {
            if (reduce_add) tma_reduce_add_2d(&maps.c, cst_s, c0, c1);
            else tma_store_2d(&maps.c, cst_s, c0, c1);
          }
          tma_store_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 4 * TX_BN);
}

static int make_tmap_f32_2d(CUtensorMap* tm, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                            uint32_t box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode_fn();
  BCI_REQUIRE(enc, BCI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BCI_REQUIRE(r == CUDA_SUCCESS, BCI_ECUDA, "cuTensorMapEncodeTiled(f32 2D) failed with CUresult %d", (int)r);
  return BCI_OK;
}
static int tx_prepare() {
  static PerDeviceFlag done_pd;
  bool& done = done_pd.cur();
  if (!done) {
    BCI_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TX_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TX_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TX_SMEM));
    done = true;
  }
  return BCI_OK;
}

bool tf32x3_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_FP32_GEMM");  // "simt": CUDA-core GEMMs (the first version of the fp32 path)
    v = (e && e[0] == 's') ? 0 : 1;
  }
  return v != 0;
}
bool tf32x3_nt_ok(const void* A, int lda, const void* W, int ldw, const void* C, int ldc, int M, int N, int K) {
  return tf32x3_enabled() && M >= 128 && N % 32 == 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0 &&
         ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0;
}
bool tf32x3_tn_ok(const void* A, int lda, const void* B, int ldb, const void* C, int ldc, long long R, int P, int Q) {
  return tf32x3_enabled() && R >= 1024 && R < (1ll << 31) && P % 128 == 0 && Q % 128 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
         ldc % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0;
}

static bool tf32_pair_enabled();
static int tf32_pair_max_clusters();
static int gemm_tf32_pair_nt(const float* A, const float* A_lo, int lda, const float* W, const float* W_lo, int ldw, const float* bias,
                             float* C, int ldc, int M, int N, int K, int accumulate, cudaStream_t st);
static int gemm_tf32_pair_tn(const float* A, const float* A_lo, int lda, const float* B, const float* B_lo, int ldb, float* C, int ldc,
                             long long R, int P, int Q, cudaStream_t st, int force_splits, int qcols);

// C[M][N] (ldc) (=|+=) A[M][K] (lda) . W[N][K]^T (ldw) + bias[N];  *_lo from split_tf32 (hi = the raw array, or the rounded copy)
int gemm_tf32x3_nt(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw, const float* bias,
                   float* C, int ldc, int M, int N, int K, int accumulate, cudaStream_t st) {
  int rc = tx_prepare();
  if (rc) return rc;
  TxMaps maps;
  // A_lo == W_lo == nullptr: single-pass TF32 product (the mixed-precision training step)
  const int single = (A_lo == nullptr && W_lo == nullptr) ? 1 : 0;
  BCI_REQUIRE(single || (A_lo && W_lo), BCI_EINVAL, "gemm_tf32x3_nt: both remainders or neither");
  if (M >= 512 && N >= (single ? 256 : 128) && tf32_pair_enabled() && tf32_pair_max_clusters() > 0)
    return gemm_tf32_pair_nt(A_hi, single ? nullptr : A_lo, lda, W_hi, single ? nullptr : W_lo, ldw, bias, C, ldc, M, N, K, accumulate, st);
  if (single) { A_lo = A_hi; W_lo = W_hi; }
  if ((rc = make_tmap_f32_2d(&maps.a_hi, A_hi, M, K, lda, TX_BK, TX_BM))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.a_lo, A_lo, M, K, lda, TX_BK, TX_BM))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_hi, W_hi, N, K, ldw, TX_BK, TX_BN))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_lo, W_lo, N, K, ldw, TX_BK, TX_BN))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.c, C, M, N, ldc, 32, 32))) return rc;
  const long long tiles = (long long)ceil_div(M, TX_BM) * ceil_div(N, TX_BN);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  gemm_tf32x3_kernel<false, false><<<grid, TX_THREADS, TX_SMEM, st>>>(maps, bias, M, N, (long long)K, 1, accumulate, 1.0f, single);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---- single-pass TF32 NT product on CTA PAIRS (cta_group::2): 256 x 256 output tiles ------------------------------------------
// The single-pass products of the mixed training step (one TF32 MMA per product, fp32 operands) are bound by the L2 -> shared
// memory operand stream, not by the tensor pipe: a 128 x 128 tile moves 32 KB per 32 K values = 32 FLOP per byte, and the step's
// GEMMs sat at ~320 TFLOP/s = 11 TB/s of L2 reads with the ring six deep and the epilogue on eight warps.  A CTA pair computes a
// 256 x 256 tile with ONE tcgen05.mma.cta_group::2 (M 256 x N 256 x K 8) per K slice: each CTA loads its own 128 rows of A and HALF
// of the 256 rows of W (the pair's MMA reads both halves), i.e. the same 32 KB per k-block per CTA for twice the FLOPs.
// Roles per CTA: warp 0 = TMA producer (its loads are counted on the LEADER's full barrier), warp 1 = MMA issuer (leader) /
// accumulator-drained relay (peer), warps 2-9 = epilogue of the CTA's own 128 rows x 256 columns (TMEM lanes = rows).
constexpr int TP_STAGES = 5;
constexpr uint32_t TP_STAGE = 2 * TX_TILE;   // A (this CTA's 128 rows) + B (this CTA's 128 of the tile's 256 W rows), 32 K values each
constexpr size_t TP_SMEM = 1024 + (size_t)TP_STAGES * TP_STAGE + TX_CSTAGE + 256 * sizeof(float) + 256;

__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
struct TpMaps { CUtensorMap a, b, c, a_lo, b_lo; };

// TN = false: C[M][N] (=|+=) A[M][K] . W[N][K]^T + bias, K-major operands, 256 x 256 tiles
// TN = true : C[P=M][Q=N] += sum_r A[r][P] . B[r][Q] over the R = K rows in k_splits ranges (MN-major operands, boxes of {32 floats, 32
//             rows}, partial tiles reduce-added); 256 x (2 nhalf) tiles, nhalf = 128 or 64 columns of B per CTA
template <bool TN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TX_THREADS, 1)
gemm_tf32_pair_kernel(const __grid_constant__ TpMaps maps, const float* __restrict__ bias, int M, int N, long long K, int k_splits,
                      int reduce_add, int nhalf, int terms, int c_half, int f16, float out_scale) {   // terms = 3: split precision (nhalf = 64)
  // f16 (NT only): operands are fp16 (hi, lo) pairs -- 64 K values per 128-byte tile row, kind::f16 MMAs (twice the TF32 rate),
  // C = out_scale * acc + bias (the weights of that form are stored x 16)
  // c_half: C is fp16 (the mixed step's G): 64-column TMA boxes, two accumulator slabs per store
  extern __shared__ uint8_t tx_smem_raw[];
  const uint32_t raw = smem_u32(tx_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = tx_smem_raw + (base - raw);
  const uint32_t sRing = base, sC = base + TP_STAGES * TP_STAGE;
  uint8_t* genC = gen + TP_STAGES * TP_STAGE;
  float* bias_s = reinterpret_cast<float*>(genC + TX_CSTAGE);
  uint8_t* ctl = genC + TX_CSTAGE + 256 * sizeof(float);
  const uint32_t bar0 = smem_u32(ctl);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                            // leader: both CTAs' tiles of the stage have landed
  auto empty_bar = [&](int s) { return bar0 + 8u * (TP_STAGES + s); };             // every CTA: the pair's MMAs have read the stage
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * TP_STAGES + a); };         // every CTA: accumulator a is complete
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * TP_STAGES + 2 + a); };    // every CTA: its epilogue has drained accumulator a
  auto peer_tempty_bar = [&](int a) { return bar0 + 8u * (2 * TP_STAGES + 4 + a); };  // leader: relay of the peer's tempty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * (2 * TP_STAGES + 6));

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ntile = 2 * nhalf;                      // tile width
  const int m_blocks = (M + 255) / 256, n_blocks = (N + ntile - 1) / ntile;
  const int bk = f16 ? 2 * TX_BK : TX_BK;   // K values per k-block
  const long long kb_total = (K + bk - 1) / bk;
  const long long kb_per = (kb_total + k_splits - 1) / k_splits;
  const long long tiles = (long long)m_blocks * n_blocks * k_splits;
  const int n_clusters = (int)cluster_nclusters_x(), cid = (int)cluster_id_x();
  // a stage: A (this CTA's 128 rows) [, its remainder], this CTA's nhalf rows / columns of B [, their remainder]
  const uint32_t b_bytes = (uint32_t)nhalf * TX_BK * 4u;
  const uint32_t stage_bytes = terms == 3 ? 2 * TX_TILE + 2 * b_bytes : TX_TILE + b_bytes;
  const uint32_t stage_stride = terms == 3 ? 3 * TX_TILE : TP_STAGE;       // 48 KB x 3 stages or 32 KB x 5
  const int nstages = terms == 3 ? 3 : TP_STAGES;
  const uint32_t off_b = terms == 3 ? 2 * TX_TILE : TX_TILE;
  auto decode = [&](long long t, int& nb, int& mb, long long& kb0, long long& kb1) {
    nb = (int)(t % n_blocks);
    mb = (int)((t / n_blocks) % m_blocks);
    const int ks = (int)(t / ((long long)n_blocks * m_blocks));
    kb0 = ks * kb_per;
    kb1 = (kb0 + kb_per < kb_total) ? kb0 + kb_per : kb_total;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.a); tma_prefetch_desc(&maps.b); tma_prefetch_desc(&maps.c);
    for (int s = 0; s < TP_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); mbar_init(peer_tempty_bar(a), 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();   // both CTAs' barriers exist before anything is signalled across the pair

  if (warp == 0) {
    // ---- TMA producer: this CTA's A rows / columns and its half of the B operand; bytes are counted on the leader's full barrier
    int stage = 0; uint32_t phase = 0;
    for (long long t = cid; t < tiles; t += n_clusters) {
      int nb, mb; long long kb0, kb1;
      decode(t, nb, mb, kb0, kb1);
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (tp_elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * stage_bytes);
          const uint32_t s0 = sRing + stage * stage_stride;
          if (TN) {
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_load_2d_2sm(s0 + g * 4096, &maps.a, mb * 256 + (int)rank * 128 + g * 32, (int)(kb * TX_BK), full_bar(stage));
            for (int g = 0; g < nhalf / 32; ++g)
              tma_load_2d_2sm(s0 + off_b + g * 4096, &maps.b, nb * ntile + (int)rank * nhalf + g * 32, (int)(kb * TX_BK), full_bar(stage));
            if (terms == 3) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                tma_load_2d_2sm(s0 + TX_TILE + g * 4096, &maps.a_lo, mb * 256 + (int)rank * 128 + g * 32, (int)(kb * TX_BK), full_bar(stage));
              for (int g = 0; g < nhalf / 32; ++g)
                tma_load_2d_2sm(s0 + off_b + b_bytes + g * 4096, &maps.b_lo, nb * ntile + (int)rank * nhalf + g * 32, (int)(kb * TX_BK), full_bar(stage));
            }
          } else {
            tma_load_2d_2sm(s0, &maps.a, (int)(kb * bk), mb * 256 + (int)rank * 128, full_bar(stage));
            tma_load_2d_2sm(s0 + off_b, &maps.b, (int)(kb * bk), nb * ntile + (int)rank * nhalf, full_bar(stage));
            if (terms == 3) {
              tma_load_2d_2sm(s0 + TX_TILE, &maps.a_lo, (int)(kb * bk), mb * 256 + (int)rank * 128, full_bar(stage));
              tma_load_2d_2sm(s0 + off_b + b_bytes, &maps.b_lo, (int)(kb * bk), nb * ntile + (int)rank * nhalf, full_bar(stage));
            }
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---- MMA issuer
      const uint32_t idesc = f16 ? umma_idesc_f16_k(256, ntile) : umma_idesc_tf32(256, ntile, TN ? 1 : 0);
      constexpr uint32_t KSTEP = TN ? 1024u : 32u;
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (long long t = cid; t < tiles; t += n_clusters) {
        int nb, mb; long long kb0, kb1;
        decode(t, nb, mb, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        mbar_wait_cluster(peer_tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u, d_lo = d_tmem + 128u;   // split precision: second accumulator (ntile = 128)
        for (long long kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (tp_elect_one()) {
            const uint32_t s0 = sRing + stage * stage_stride;
#pragma unroll
            for (int kk = 0; kk < TX_BK / 8; ++kk) {
              const uint32_t o = kk * KSTEP;
              const uint64_t ah = TN ? umma_desc_sw128_mn(s0 + o) : umma_desc_sw128(s0 + o);
              const uint64_t bh = TN ? umma_desc_sw128_mn(s0 + off_b + o) : umma_desc_sw128(s0 + off_b + o);
              const uint32_t first = (kb != kb0 || kk != 0) ? 1u : 0u;
              if (terms == 3) {
                const uint64_t al = TN ? umma_desc_sw128_mn(s0 + TX_TILE + o) : umma_desc_sw128(s0 + TX_TILE + o);
                const uint64_t bl = TN ? umma_desc_sw128_mn(s0 + off_b + b_bytes + o) : umma_desc_sw128(s0 + off_b + b_bytes + o);
                if (f16) { umma_bf16_2sm(d_lo, al, bh, idesc, first); umma_bf16_2sm(d_lo, ah, bl, idesc, 1u); }
                else { umma_tf32_2sm(d_lo, al, bh, idesc, first); umma_tf32_2sm(d_lo, ah, bl, idesc, 1u); }
              }
              if (f16) umma_bf16_2sm(d_tmem, ah, bh, idesc, first);
              else umma_tf32_2sm(d_tmem, ah, bh, idesc, first);
            }
            umma_commit_2sm_mc(empty_bar(stage), (uint16_t)3);
            if (kb == kb1 - 1) umma_commit_2sm_mc(tfull_bar(acc), (uint16_t)3);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else {
      // ---- peer: tells the leader when this CTA's epilogue has drained an accumulator (relaxed: it publishes no data)
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t remote0 = mapa_u32(peer_tempty_bar(0), 0);
      for (long long t = cid; t < tiles; t += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase);
        if (tp_elect_one()) mbar_arrive_cluster_relaxed(remote0 + 8u * (uint32_t)acc);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ---- epilogue: this CTA's 128 rows (TMEM lanes) x ntile columns; warp w: lane quarter w % 4, half of the 32-column slabs
    const int quarter = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    const int spw = ntile / 64;   // slabs per warp: 4 (256 columns) or 2 (128)
    uint8_t* cst = genC + (warp - 2) * 4096;
    const uint32_t cst_s = sC + (warp - 2) * 4096;
    const int et = (warp - 2) * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long t = cid; t < tiles; t += n_clusters) {
      int nb, mb; long long kb0, kb1;
      decode(t, nb, mb, kb0, kb1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bias_s[et] = (bias != nullptr && kb0 == 0 && nb * ntile + et < N) ? __ldg(bias + nb * ntile + et) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * 256u;
      if (c_half) {
#pragma unroll 1
        for (int sp = 0; sp < spw / 2; ++sp) {
          const int slab = spw * ehalf + 2 * sp;
          uint32_t r[32], r2[32];
          tmem_ld32(taddr + slab * 32, r);
          tmem_ld32(taddr + slab * 32 + 32, r2);
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          tmem_ld_wait();
          if (sp == spw / 2 - 1) {
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {   // chunk q = columns [8 q, 8 q + 8) of the 64-column box
            const uint32_t* src = q < 4 ? r + 8 * q : r2 + 8 * (q - 4);
            const float* bs = bias_s + slab * 32 + 8 * q;
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h2 = __floats2half2_rn(__uint_as_float(src[2 * e]) + bs[2 * e], __uint_as_float(src[2 * e + 1]) + bs[2 * e + 1]);
              pk[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(cst + sw128_chunk_off((uint32_t)lane, (uint32_t)q)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int c0 = nb * ntile + slab * 32, c1 = mb * 256 + (int)rank * 128 + quarter * 32;
            if (c0 < N && c1 < M) tma_store_2d(&maps.c, cst_s, c0, c1);
            tma_store_commit();
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
#pragma unroll 1
      for (int slab = spw * ehalf; slab < spw * ehalf + spw; ++slab) {
        uint32_t r[32], rl[32];
        tmem_ld32(taddr + slab * 32, r);
        if (terms == 3) {
          tmem_ld32(taddr + 128 + slab * 32, rl);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) rl[q] = 0u;
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        tmem_ld_wait();
        if (slab == spw * ehalf + spw - 1) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v;
          v.x = fmaf(__uint_as_float(r[4 * q + 0]) + __uint_as_float(rl[4 * q + 0]), out_scale, bias_s[slab * 32 + 4 * q + 0]);
          v.y = fmaf(__uint_as_float(r[4 * q + 1]) + __uint_as_float(rl[4 * q + 1]), out_scale, bias_s[slab * 32 + 4 * q + 1]);
          v.z = fmaf(__uint_as_float(r[4 * q + 2]) + __uint_as_float(rl[4 * q + 2]), out_scale, bias_s[slab * 32 + 4 * q + 2]);
          v.w = fmaf(__uint_as_float(r[4 * q + 3]) + __uint_as_float(rl[4 * q + 3]), out_scale, bias_s[slab * 32 + 4 * q + 3]);
          *reinterpret_cast<float4*>(cst + sw128_chunk_off((uint32_t)lane, (uint32_t)q)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int c0 = nb * ntile + slab * 32, c1 = mb * 256 + (int)rank * 128 + quarter * 32;
          if (c0 < N && c1 < M) {
            if (reduce_add) tma_reduce_add_2d(&maps.c, cst_s, c0, c1);
            else tma_store_2d(&maps.c, cst_s, c0, c1);
          }
          tma_store_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its shared memory or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int make_tmap_f16_2d(CUtensorMap* tm, const __half* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                            uint32_t box_rows);
// BCI_GEMM_PAIR=off keeps the one-CTA kernel for the single-pass products
static bool tf32_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_GEMM_PAIR");
    v = (e && e[0] == 'o') ? 0 : 1;
  }
  return v != 0;
}
static int tf32_pair_max_clusters() {
  static PerDeviceInt state_pd, max_pd;   // state: 0 = not tried, 1 = ok, -1 = unavailable
  int& state = state_pd.cur();
  int& mx = max_pd.cur();
  if (state == 0) {
    state = -1;
    if (cudaFuncSetAttribute(gemm_tf32_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(gemm_tf32_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count(), 1, 1);
    cfg.blockDim = dim3(TX_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TP_SMEM;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    cfg.attrs = la; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&mx, gemm_tf32_pair_kernel<false>, &cfg) != cudaSuccess || mx <= 0) { mx = 0; return 0; }
    state = 1;
  }
  return state == 1 ? mx : 0;
}
static int gemm_tf32_pair_nt(const float* A, const float* A_lo, int lda, const float* W, const float* W_lo, int ldw, const float* bias,
                             float* C, int ldc, int M, int N, int K, int accumulate, cudaStream_t st) {
  TpMaps maps;
  int rc;
  const int terms = A_lo ? 3 : 1;
  const int nhalf = terms == 3 ? 64 : 128;   // split precision keeps two accumulators per tile: 256 x 128 tiles
  if ((rc = make_tmap_f32_2d(&maps.a, A, M, K, lda, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b, W, N, K, ldw, TX_BK, nhalf))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.a_lo, A_lo ? A_lo : A, M, K, lda, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_lo, W_lo ? W_lo : W, N, K, ldw, TX_BK, nhalf))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.c, C, M, N, ldc, 32, 32))) return rc;
  const long long tiles = (long long)ceil_div(M, 256) * ceil_div(N, 2 * nhalf);
  const int mx = tf32_pair_max_clusters();
  const int clusters = (int)(tiles < mx ? tiles : mx);
  gemm_tf32_pair_kernel<false><<<2 * clusters, TX_THREADS, TP_SMEM, st>>>(maps, bias, M, N, (long long)K, 1, accumulate, nhalf, terms, 0, 0, 1.0f);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
// single-pass NT product with an fp16 result (the mixed step's projected inputs G): C16[M][N] = fp16(A . W^T + bias).  Needs the pair
// kernel (M >= 512, N >= 256, N % 64 == 0, ldc % 8 == 0); callers check gemm_tf32_half_ok first
bool gemm_tf32_half_ok(const void* A, int lda, const void* W, int ldw, const void* C, int ldc, int M, int N, int K) {
  return tf32x3_enabled() && tf32_pair_enabled() && M >= 512 && N >= 256 && N % 64 == 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 &&
         ldc % 8 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)C & 15) == 0 && tf32_pair_max_clusters() > 0;
}
int gemm_tf32_nt_half(const float* A, int lda, const float* W, int ldw, const float* bias, __half* C, int ldc, int M, int N, int K,
                      cudaStream_t st) {
  TpMaps maps;
  int rc;
  if ((rc = make_tmap_f32_2d(&maps.a, A, M, K, lda, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b, W, N, K, ldw, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.a_lo, A, M, K, lda, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_lo, W, N, K, ldw, TX_BK, 128))) return rc;
  if ((rc = make_tmap_f16_2d(&maps.c, C, M, N, ldc, 64, 32))) return rc;
  const long long tiles = (long long)ceil_div(M, 256) * ceil_div(N, 256);
  const int mx = tf32_pair_max_clusters();
  const int clusters = (int)(tiles < mx ? tiles : mx);
  gemm_tf32_pair_kernel<false><<<2 * clusters, TX_THREADS, TP_SMEM, st>>>(maps, bias, M, N, (long long)K, 1, 0, 128, 1, 1, 0, 1.0f);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
// C[P][Q] = sum_r A[r][P] . B[r][Q] on CTA pairs: P % 256 == 0, Q % 128 == 0
static int gemm_tf32_pair_tn(const float* A, const float* A_lo, int lda, const float* B, const float* B_lo, int ldb, float* C, int ldc,
                             long long R, int P, int Q, cudaStream_t st, int force_splits, int qcols) {
  TpMaps maps;
  int rc;
  const int terms = A_lo ? 3 : 1;
  if ((rc = make_tmap_f32_2d(&maps.a, A, R, P, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b, B, R, qcols, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.a_lo, A_lo ? A_lo : A, R, P, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_lo, B_lo ? B_lo : B, R, qcols, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.c, C, P, Q, ldc, 32, 32))) return rc;
  const int nhalf = (terms == 1 && Q % 256 == 0) ? 128 : 64;
  const int out_tiles = (P / 256) * (Q / (2 * nhalf));
  const int mx = tf32_pair_max_clusters();
  const long long kb_total = (R + TX_BK - 1) / TX_BK;
  long long splits = (mx + out_tiles - 1) / out_tiles;
  if (splits > kb_total / 8) splits = kb_total / 8;
  // split precision: at most 1024 rows per accumulator (the TMEM accumulation error grows with the range, see gemm_tf32x3_tn)
  if (terms == 3 && splits < (kb_total + 31) / 32) splits = (kb_total + 31) / 32;
  if (splits < 1) splits = 1;
  if (force_splits > 0) splits = force_splits;
  const long long per = (kb_total + splits - 1) / splits;
  splits = (kb_total + per - 1) / per;
  if (splits > 1) BCI_CUDA_OK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)Q * 4, P, st));
  const long long tiles = out_tiles * splits;
  const int clusters = (int)(tiles < mx ? tiles : mx);
  gemm_tf32_pair_kernel<true><<<2 * clusters, TX_THREADS, TP_SMEM, st>>>(maps, nullptr, P, Q, R, (int)splits, splits > 1 ? 1 : 0, nhalf, terms, 0, 0, 1.0f);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// C[P][Q] (ldc) = sum over r < R of A[r][P] (lda) . B[r][Q] (ldb)   (C is overwritten)
// q_valid > 0: B has only q_valid (< Q) columns in memory (row stride ldb); the rest of the tile is zero-filled by the TMA unit
int gemm_tf32x3_tn(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                   long long R, int P, int Q, cudaStream_t st, int force_splits, int q_valid) {
  int rc = tx_prepare();
  if (rc) return rc;
  TxMaps maps;
  const int single = (A_lo == nullptr && B_lo == nullptr) ? 1 : 0;
  BCI_REQUIRE(single || (A_lo && B_lo), BCI_EINVAL, "gemm_tf32x3_tn: both remainders or neither");
  const int qcols = q_valid > 0 ? q_valid : Q;
  if (P % 256 == 0 && Q % 128 == 0 && tf32_pair_enabled() && tf32_pair_max_clusters() > 0)
    return gemm_tf32_pair_tn(A_hi, single ? nullptr : A_lo, lda, B_hi, single ? nullptr : B_lo, ldb, C, ldc, R, P, Q, st, force_splits, qcols);
  if (single) { A_lo = A_hi; B_lo = B_hi; }
  if ((rc = make_tmap_f32_2d(&maps.a_hi, A_hi, R, P, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.a_lo, A_lo, R, P, lda, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_hi, B_hi, R, qcols, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.b_lo, B_lo, R, qcols, ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.c, C, P, Q, ldc, 32, 32))) return rc;
  const int out_tiles = (P / TX_BM) * (Q / TX_BN);
  const long long kb_total = (R + TX_BK - 1) / TX_BK;
  // K ranges of at most 1024 rows per accumulator (the TMEM accumulation error grows with the range: 3e-5 relative at 4096
  // rows in one accumulator, 2e-6 when the same sum is split four ways and combined by the fp32 reduce-add), and enough
  // of them to fill the machine
  long long splits = (sm_count() + out_tiles - 1) / out_tiles;
  if (splits > kb_total / 8) splits = kb_total / 8;
  if (splits < (kb_total + 31) / 32) splits = (kb_total + 31) / 32;
  if (splits < 1) splits = 1;
  if (force_splits > 0) splits = force_splits;
  const long long per = (kb_total + splits - 1) / splits;
  splits = (kb_total + per - 1) / per;  // every split owns at least one k block
  if (splits > 1) BCI_CUDA_OK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)Q * 4, P, st));
  const long long tiles = out_tiles * splits;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  gemm_tf32x3_kernel<true, false><<<grid, TX_THREADS, TX_SMEM, st>>>(maps, nullptr, P, Q, R, (int)splits, splits > 1 ? 1 : 0, 1.0f, single);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---- fp16-split variant (NT): x -> (hi, lo) fp16 pair, optionally pre-scaled -------------------------------------------------
__global__ void split_f16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, long long n4, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldg(x + i);
  const float in[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  __half2 h[2], l[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    h[j] = __floats2half2_rn(in[2 * j], in[2 * j + 1]);
    const float2 back = __half22float2(h[j]);
    l[j] = __floats2half2_rn(in[2 * j] - back.x, in[2 * j + 1] - back.y);
  }
  hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h[0]), *reinterpret_cast<const uint32_t*>(&h[1]));
  lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l[0]), *reinterpret_cast<const uint32_t*>(&l[1]));
}

int split_f16(const float* x, __half* hi, __half* lo, long long n, float scale, cudaStream_t st) {
  BCI_REQUIRE(n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)hi & 7) == 0 && ((uintptr_t)lo & 7) == 0, BCI_EINVAL,
              "split_f16: aligned arrays with n %% 4 == 0 required");
  if (n == 0) return BCI_OK;
  split_f16_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(hi),
                                                                      reinterpret_cast<uint2*>(lo), n / 4, scale);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

static int make_tmap_f16_2d(CUtensorMap* tm, const __half* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                            uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  BCI_REQUIRE(enc, BCI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BCI_REQUIRE(r == CUDA_SUCCESS, BCI_ECUDA, "cuTensorMapEncodeTiled(f16 2D) failed with CUresult %d", (int)r);
  return BCI_OK;
}

bool f16x3_nt_ok(const void* A_hi, int lda, const void* W_hi, int ldw, const void* C, int ldc, int M, int N, int K) {
  return tf32x3_enabled() && M >= 128 && N % 32 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 4 == 0 &&
         ((uintptr_t)A_hi & 15) == 0 && ((uintptr_t)W_hi & 15) == 0 && ((uintptr_t)C & 15) == 0;
}

// C[M][N] (ldc, fp32) = out_scale * (A[M][K] . W[N][K]^T) + bias[N], operands as fp16 (hi, lo) pairs
int gemm_f16x3_nt(const __half* A_hi, const __half* A_lo, int lda, const __half* W_hi, const __half* W_lo, int ldw, const float* bias,
                  float* C, int ldc, int M, int N, int K, float out_scale, cudaStream_t st) {
  int rc = tx_prepare();
  if (rc) return rc;
  if (M >= 512 && N >= 128 && tf32_pair_enabled() && tf32_pair_max_clusters() > 0) {
    // CTA pairs, 256 x 128 tiles (two accumulators per tile): each CTA loads its own 128 rows of A (hi, lo) and 64 of the 128 W rows
    TpMaps pm;
    if ((rc = make_tmap_f16_2d(&pm.a, A_hi, M, K, lda, 2 * TX_BK, 128))) return rc;
    if ((rc = make_tmap_f16_2d(&pm.a_lo, A_lo, M, K, lda, 2 * TX_BK, 128))) return rc;
    if ((rc = make_tmap_f16_2d(&pm.b, W_hi, N, K, ldw, 2 * TX_BK, 64))) return rc;
    if ((rc = make_tmap_f16_2d(&pm.b_lo, W_lo, N, K, ldw, 2 * TX_BK, 64))) return rc;
    if ((rc = make_tmap_f32_2d(&pm.c, C, M, N, ldc, 32, 32))) return rc;
    const long long ptiles = (long long)ceil_div(M, 256) * ceil_div(N, 128);
    const int mx = tf32_pair_max_clusters();
    const int clusters = (int)(ptiles < mx ? ptiles : mx);
    gemm_tf32_pair_kernel<false><<<2 * clusters, TX_THREADS, TP_SMEM, st>>>(pm, bias, M, N, (long long)K, 1, 0, 64, 3, 0, 1, out_scale);
    BCI_LAUNCH_OK();
    return BCI_OK;
  }
  TxMaps maps;
  if ((rc = make_tmap_f16_2d(&maps.a_hi, A_hi, M, K, lda, 2 * TX_BK, TX_BM))) return rc;
  if ((rc = make_tmap_f16_2d(&maps.a_lo, A_lo, M, K, lda, 2 * TX_BK, TX_BM))) return rc;
  if ((rc = make_tmap_f16_2d(&maps.b_hi, W_hi, N, K, ldw, 2 * TX_BK, TX_BN))) return rc;
  if ((rc = make_tmap_f16_2d(&maps.b_lo, W_lo, N, K, ldw, 2 * TX_BK, TX_BN))) return rc;
  if ((rc = make_tmap_f32_2d(&maps.c, C, M, N, ldc, 32, 32))) return rc;
  const long long tiles = (long long)ceil_div(M, TX_BM) * ceil_div(N, TX_BN);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  gemm_tf32x3_kernel<false, true><<<grid, TX_THREADS, TX_SMEM, st>>>(maps, bias, M, N, (long long)K, 1, 0, out_scale, 0);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci

// Diagnostic entry point: mode 0  C[M][N] = A[M][K] . B[N][K]^T + bias;  mode 1  C[M][N] = A[K][M]^T . B[K][N].
// explicit_hi != 0 rounds hi into a separate array instead of passing the raw operand.  Scratch is allocated here (test only).
extern "C" int bci_selftest_gemm_tf32x3(int32_t mode, const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N,
                                        int64_t K, int32_t explicit_hi, void* stream) {
  using namespace bci;
  cudaStream_t st = (cudaStream_t)stream;
  const long long na = (long long)M * K, nb = (long long)N * K;
  // modes 2 / 3: diagnostics -- NT accumulating into the caller's C (reduce-add epilogue) / TN without split-K (plain store)
  const bool nt_acc = mode == 2, tn_single = mode == 3;
  if (nt_acc) mode = 0;
  if (tn_single) mode = 1;
  BCI_REQUIRE(mode == 0 || mode == 1, BCI_EINVAL, "bci_selftest_gemm_tf32x3: mode must be 0 (NT) or 1 (TN)");
  BCI_REQUIRE(na % 4 == 0 && nb % 4 == 0, BCI_EINVAL, "bci_selftest_gemm_tf32x3: operand sizes must be multiples of 4");
  float* scratch = nullptr;
  BCI_CUDA_OK(cudaMalloc(&scratch, (size_t)(na + nb) * 8));
  float *alo = scratch, *blo = scratch + na, *ahi = blo + nb, *bhi = ahi + na;
  int rc = split_tf32(A, explicit_hi ? ahi : nullptr, alo, na, st);
  if (!rc) rc = split_tf32(B, explicit_hi ? bhi : nullptr, blo, nb, st);
  const float* Ah = explicit_hi ? ahi : A;
  const float* Bh = explicit_hi ? bhi : B;
  if (!rc) {
    if (mode == 0) {
      rc = tf32x3_nt_ok(A, (int)K, B, (int)K, C, N, M, N, (int)K)
               ? gemm_tf32x3_nt(Ah, alo, (int)K, Bh, blo, (int)K, bias, C, N, M, N, (int)K, nt_acc ? 1 : 0, st) : BCI_EINVAL;
    } else {
      rc = tf32x3_tn_ok(A, M, B, N, C, N, K, M, N) ? gemm_tf32x3_tn(Ah, alo, M, Bh, blo, N, C, N, K, M, N, st, tn_single ? 1 : 0) : BCI_EINVAL;
    }
    if (rc == BCI_EINVAL) set_error("bci_selftest_gemm_tf32x3: shape not supported by the tcgen05 path");
  }
  const cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(scratch);
  BCI_REQUIRE(se == cudaSuccess, BCI_ECUDA, "bci_selftest_gemm_tf32x3: %s", cudaGetErrorString(se));
  return rc;
}


// Diagnostic entry point of the fp16-split form: C[M][N] = A[M][K] . B[N][K]^T + bias, A split as is, B scaled by 16 as the packed
// weights are.  Scratch is allocated here (test only).
extern "C" int bci_selftest_gemm_f16x3(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                                       void* stream) {
  using namespace bci;
  cudaStream_t st = (cudaStream_t)stream;
  const long long na = (long long)M * K, nb = (long long)N * K;
  BCI_REQUIRE(na % 4 == 0 && nb % 4 == 0, BCI_EINVAL, "bci_selftest_gemm_f16x3: operand sizes must be multiples of 4");
  __half* scratch = nullptr;
  BCI_CUDA_OK(cudaMalloc(&scratch, (size_t)(na + nb) * 4));
  __half *ahi = scratch, *alo = ahi + na, *bhi = alo + na, *blo = bhi + nb;
  int rc = split_f16(A, ahi, alo, na, 1.0f, st);
  if (!rc) rc = split_f16(B, bhi, blo, nb, F16X3_WSCALE, st);
  if (!rc) {
    if (f16x3_nt_ok(ahi, K, bhi, K, C, N, M, N, K)) rc = gemm_f16x3_nt(ahi, alo, K, bhi, blo, K, bias, C, N, M, N, K, 1.0f / F16X3_WSCALE, st);
    else { set_error("bci_selftest_gemm_f16x3: shape not supported"); rc = BCI_EINVAL; }
  }
  const cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(scratch);
  BCI_REQUIRE(se == cudaSuccess, BCI_ECUDA, "bci_selftest_gemm_f16x3: %s", cudaGetErrorString(se));
  return rc;
}

// single-pass TF32 NT product (the mixed training step's GEMM form): C[M][N] = A[M][K] . B[N][K]^T + bias, accumulate != 0: C += ...
// M >= 512 and N >= 256 run on CTA pairs (gemm_tf32_pair_kernel), smaller shapes on the one-CTA kernel
extern "C" int bci_selftest_gemm_tf32_single(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                                             int32_t accumulate, void* stream) {
  using namespace bci;
  BCI_REQUIRE(A && B && C && M >= 1 && N >= 1 && K >= 1, BCI_EINVAL, "bci_selftest_gemm_tf32_single: bad arguments");
  BCI_REQUIRE(tf32x3_nt_ok(A, K, B, K, C, N, M, N, K), BCI_EINVAL, "bci_selftest_gemm_tf32_single: shape not supported by the tensor-core path");
  return gemm_tf32x3_nt(A, nullptr, K, B, nullptr, K, bias, C, N, M, N, K, accumulate, (cudaStream_t)stream);
}
// ... and of the TN product (weight gradients): C[M][N] = A[K][M]^T . B[K][N]; M % 256 == 0 and N % 128 == 0 run on CTA pairs
extern "C" int bci_selftest_gemm_tf32_single_tn(const float* A, const float* B, float* C, int32_t M, int32_t N, int64_t K, void* stream) {
  using namespace bci;
  BCI_REQUIRE(A && B && C && M >= 1 && N >= 1 && K >= 1, BCI_EINVAL, "bci_selftest_gemm_tf32_single_tn: bad arguments");
  BCI_REQUIRE(tf32x3_tn_ok(A, M, B, N, C, N, K, M, N), BCI_EINVAL, "bci_selftest_gemm_tf32_single_tn: shape not supported by the tensor-core path");
  return gemm_tf32x3_tn(A, nullptr, M, B, nullptr, N, C, N, K, M, N, (cudaStream_t)stream, 0);
}

