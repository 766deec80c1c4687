// bf16 tensor-core mode of the BiLSTM forward (H = 128), hand-written for sm_100a.
//
//   K2  proj_gemm_bf16 : G = in . W_ih^T + b for all T steps and both directions -- a persistent,
//       warp-specialised tcgen05 GEMM: TMA (cp.async.bulk.tensor, SWIZZLE_128B) feeds a 4-stage
//       smem ring, one thread issues tcgen05.mma (M128 x N256 x K16, bf16 -> fp32 in TMEM), the
//       accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue warps (tcgen05.ld
//       -> +bias -> bf16 -> global) of tile i overlap the MMAs of tile i+1.
//   K3  lstm_rec_bf16  : the serial recurrence, one CTA per (128-window tile, direction), resident
//       for the whole sequence: W_hh (512x128 bf16, 128 KB) stays in shared memory in UMMA layout,
//       each step issues h_{t-1} . W_hh^T as 16 tcgen05.mma into the 512 TMEM columns, eight epilogue
//       warps read their gate slabs back (tcgen05.ld), add the projected input G_t, apply
//       sigma/tanh (MUFU tanh.approx), update the fp32 cell state held in registers for all 256
//       steps, and write h_t as bf16 straight into the swizzled A-operand buffer of the next step.
//       Gate columns are permuted (unit/8, gate, unit%8) so one 32-column TMEM slab carries i,f,g,o
//       of 8 hidden units and the cell update is thread-local.  The two N=256 halves are committed
//       to separate mbarriers, so half of the epilogue starts while the second half of the MMAs runs.
//
// Reference semantics: nn.LSTM inside EnhancedLSTMModel (04_lstm_model.py:181-188,211); on CUDA the
// reference itself runs this under autocast (04:486-490, 06:348-351).
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <cuda_fp16.h>

namespace bci {
using namespace sm100;

// ---------------------------------------------------------------------------------------------
// weight packing for the tensor-core path
// ---------------------------------------------------------------------------------------------
// Two column orders for the 512 gate pre-activations of one direction (H = 128):
//   perm_T : TMEM / W_hh row order   half*256 + slab*32 + gate*8 + u      (unit = half*64 + slab*8 + u)
//            -> thread (row, half) reads slab `slab` of its half as one 32-column tcgen05.ld
//   perm_G : memory order of G       slab*64 + half*32 + gate*8 + u
//            -> slab `slab` of BOTH halves is one contiguous 128-byte row segment = one TMA box
__host__ __device__ constexpr int perm_T(int unit, int gate) { return (unit >> 3) * 32 + gate * 8 + (unit & 7); }
__host__ __device__ constexpr int perm_G(int unit, int gate) {
  return ((unit >> 3) & 7) * 64 + (unit >> 6) * 32 + gate * 8 + (unit & 7);
}

// src (4H, K) gate-major rows -> dst bf16 [row0 + perm(unit,gate)][K];  order 0 = perm_T, 1 = perm_G
__global__ void pack_rows_perm_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H, int K, int row0, int order) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int gate = row / H, unit = row - gate * H;
  const int pr = order ? perm_G(unit, gate) : perm_T(unit, gate);
  // sigmoid(x) = 0.5 + 0.5 tanh(x/2): the 1/2 is folded into the i,f,o rows here (exact in binary floating point)
  const float sc = (gate == 2) ? 1.0f : 0.5f;
  dst[(long long)(row0 + pr) * K + k] = __float2bfloat16_rn(sc * src[i]);
}
__global__ void pack_bias_perm(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ dst, int H, int col0,
                               int order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H) return;
  const int gate = i / H, unit = i - gate * H;
  dst[col0 + (order ? perm_G(unit, gate) : perm_T(unit, gate))] = ((gate == 2) ? 1.0f : 0.5f) * (bih[i] + bhh[i]);
}

// ---- which recurrence path the bf16 forward uses -----------------------------------------------------------------------
//   fused : lstm_fused_bf16 (4-CTA clusters, W_ih + W_hh resident, G never materialised)          [default when it fits]
//   split : proj_gemm_bf16 (K2) -> G in HBM -> lstm_rec_bf16 (K3)                                  [BCI_BF16_PATH=split]
static int bf16_path_fused() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_BF16_PATH");
    if (e && e[0] == 's') v = 0;
    else v = fused_max_clusters() > 0 ? 1 : 0;
  }
  return v;
}
int bf16_chunk_windows() {
  // split: 74 x 128 windows = exactly one wave of (tile, direction) CTAs of the recurrence on 148 SMs.
  // fused: every cluster runs two work items (both directions of one tile quad) per chunk: 4 tiles per co-resident cluster.
  return bf16_path_fused() ? fused_max_clusters() * 4 * 128 : 74 * 128;
}

size_t lstm_workspace_h256(const bci_lstm_config& c, int batch, int T);
int lstm_forward_h256(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn, void* ws,
                      size_t ws_bytes, cudaStream_t st);

size_t lstm_store_bytes_bf16(const bci_lstm_config& c) {
  if (c.precision != BCI_PRECISION_BF16) return 0;
  const size_t H = c.hidden_size;
  size_t n = 0;
  if (H == 256) {
    for (int l = 0; l < c.num_layers; ++l)
      n += align_up((size_t)2048 * layer_in_width(c, l) * 2, 256) + align_up((size_t)2048 * 256 * 2, 256) + align_up(2048 * 4, 256);
    n += align_up((size_t)256 * 512 * 2, 256) + align_up(256 * sizeof(float4), 256) + align_up(256 * 4, 256);  // attention W1', params, zeros
    n += align_up((size_t)256 * 64 * 2, 256);                                                                  // input projection W0 (bf16, K padded)
    return n + 1024;
  }
  for (int l = 0; l < c.num_layers; ++l)
    n += 2 * align_up((size_t)8 * H * layer_in_width(c, l) * 2, 256) + 2 * align_up(4 * H * H * 2, 256) + 2 * align_up(8 * H * 4, 256);
  n += align_up(H * 2 * H * 2, 256) + align_up(H * sizeof(float4), 256);  // attention W1' (bf16) + per-unit params
  n += align_up(H * 64 * 2, 256) + align_up(H * sizeof(float4), 256);     // input projection W0 (bf16, K padded) + params
  return n + 1024;
}

void lstm_carve_bf16(bci_lstm_s* h, char* base) {
  const bci_lstm_config& c = h->cfg;
  if (c.precision != BCI_PRECISION_BF16) return;
  const size_t H = c.hidden_size;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base + off; off += align_up(bytes, 256); return p; };
  if (H == 256) {
    for (int l = 0; l < c.num_layers; ++l) {
      h->bf16.wih256[l] = reinterpret_cast<__nv_bfloat16*>(take((size_t)2048 * layer_in_width(c, l) * 2));
      h->bf16.whh256[l] = reinterpret_cast<__nv_bfloat16*>(take((size_t)2048 * 256 * 2));
      h->bf16.bias256[l] = reinterpret_cast<float*>(take(2048 * 4));
    }
    h->bf16.aw1_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)256 * 512 * 2));
    h->bf16.apar = reinterpret_cast<float4*>(take(256 * sizeof(float4)));
    h->bf16.zero_bias = reinterpret_cast<float*>(take(256 * 4));
    h->bf16.w0_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)256 * 64 * 2));
    return;
  }
  for (int l = 0; l < c.num_layers; ++l) {
    h->bf16.wih_bf[l] = reinterpret_cast<__nv_bfloat16*>(take((size_t)8 * H * layer_in_width(c, l) * 2));
    h->bf16.whh_bf[l][0] = reinterpret_cast<__nv_bfloat16*>(take(4 * H * H * 2));
    h->bf16.whh_bf[l][1] = reinterpret_cast<__nv_bfloat16*>(take(4 * H * H * 2));
    h->bf16.bias_p[l] = reinterpret_cast<float*>(take(8 * H * 4));
    h->bf16.wih_t_bf[l] = reinterpret_cast<__nv_bfloat16*>(take((size_t)8 * H * layer_in_width(c, l) * 2));
    h->bf16.bias_t[l] = reinterpret_cast<float*>(take(8 * H * 4));
  }
  h->bf16.aw1_bf = reinterpret_cast<__nv_bfloat16*>(take(H * 2 * H * 2));
  h->bf16.apar = reinterpret_cast<float4*>(take(H * sizeof(float4)));
  h->bf16.w0_bf = reinterpret_cast<__nv_bfloat16*>(take(H * 64 * 2));
  h->bf16.par0 = reinterpret_cast<float4*>(take(H * sizeof(float4)));
}

int lstm_pack_bf16(bci_lstm_s* h, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const bci_lstm_weights& w = h->raw;
  const int H = c.hidden_size;
  if (H == 256) return pack_h256_bf16(h, st);
  BCI_REQUIRE(H == 128, BCI_EINVAL, "bf16 (tcgen05) mode is built for hidden_size 128 and 256 (got %d)", H);
  for (int l = 0; l < c.num_layers; ++l) {
    const int K = layer_in_width(c, l);
    for (int d = 0; d < 2; ++d) {
      pack_rows_perm_bf16<<<(unsigned)ceil_div64((long long)4 * H * K, 256), 256, 0, st>>>(w.w_ih[l][d], h->bf16.wih_bf[l], H, K, d * 4 * H, 1);
      pack_rows_perm_bf16<<<(unsigned)ceil_div64((long long)4 * H * H, 256), 256, 0, st>>>(w.w_hh[l][d], h->bf16.whh_bf[l][d], H, H, 0, 0);
      pack_bias_perm<<<ceil_div(4 * H, 256), 256, 0, st>>>(w.b_ih[l][d], w.b_hh[l][d], h->bf16.bias_p[l], H, d * 4 * H, 1);
      pack_rows_perm_bf16<<<(unsigned)ceil_div64((long long)4 * H * K, 256), 256, 0, st>>>(w.w_ih[l][d], h->bf16.wih_t_bf[l], H, K, d * 4 * H, 0);
      pack_bias_perm<<<ceil_div(4 * H, 256), 256, 0, st>>>(w.b_ih[l][d], w.b_hh[l][d], h->bf16.bias_t[l], H, d * 4 * H, 0);
    }
  }
  int rc = pack_pool_bf16(h, st);
  if (rc) return rc;
  rc = pack_inproj_bf16(h, st);
  if (rc) return rc;
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---------------------------------------------------------------------------------------------
// K2: persistent TMA + tcgen05 GEMM.  C[M][N] (bf16) = A[M][K] (bf16) . W[N][K]^T (bf16) + bias[N]
// Each CTA owns ONE 256-column block of W for its whole life (W block resident in smem, loaded once
// by TMA) and streams 128-row A tiles through a 4-stage ring, so per output tile only A (K*256 B)
// and C cross L2 -- the first version re-fetched the W block per tile and was L2-bound (ncu:
// lts throughput 72 %, tensor pipe 26 %).
// ---------------------------------------------------------------------------------------------
constexpr int GB_BM = 128, GB_BK = 64, GB_MAX_STAGES = 8;
// BN = 256 (K <= 256) is the H = 128 configuration; BN = 128 keeps a 128-column W block of up to K = 512 resident (H = 256:
// layers 1-2 have a 512-wide input)
template <int BN> struct GbCfg { static constexpr int MAX_K = BN == 256 ? 256 : 512; };
constexpr int GB_THREADS = 320;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue (2 groups of 4)
constexpr uint32_t GB_A_BYTES = GB_BM * GB_BK * 2;
constexpr uint32_t GB_C_BYTES = GB_BM * 128 * 2;  // C staging: one [128][64] bf16 SW128 atom per epilogue group
// smem: [W block: k_blocks x BN x 128 B][A ring: stages x 16 KB][C staging 32 KB][bias][barriers]
static inline int gb_stages(int K) { return K <= 128 ? 8 : 4; }
static inline size_t gb_smem(int K, int BN) {
  return 1024 + (size_t)(K / GB_BK) * (size_t)(BN * GB_BK * 2) + (size_t)gb_stages(K) * GB_A_BYTES + GB_C_BYTES + 256 * sizeof(float) + 256;
}

// BLOCKED = false: C row-major [M][N] through a tensor map.
// BLOCKED = true : C in the recurrence's streaming layout  [m_block = row/128][n/8 (16-byte chunk)][row%128][8 bf16]
//                  (N = 1024: 256 KB per m_block).  A 64-column pass of the epilogue is then 8 chunks x 2 KB = one contiguous
//                  16 KB block: staged as [chunk][row] (conflict-free 16-byte stores) and written with one 1-D bulk store;
//                  the recurrence reads it back with fully coalesced 16-byte loads (lane = row).
template <bool BLOCKED, int GB_BN>
__global__ void __launch_bounds__(GB_THREADS, 1)
proj_gemm_bf16(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, __nv_bfloat16* __restrict__ Cblk, const float* __restrict__ bias, int M,
               int N, int K, int GB_STAGES) {
  extern __shared__ uint8_t gb_smem_raw[];
  const uint32_t raw = smem_u32(gb_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = gb_smem_raw + (base - raw);  // generic pointer to the aligned base
  constexpr uint32_t GB_B_ATOM = GB_BN * GB_BK * 2;
  constexpr int GCOLS = GB_BN / 2;      // columns per epilogue group
  constexpr int PASSES = GCOLS / 64;    // 64-column staging passes per group (2 for BN = 256, 1 for BN = 128)
  const uint32_t b_bytes = (uint32_t)(K / GB_BK) * GB_B_ATOM;
  const uint32_t sB = base, sA = base + b_bytes, sC = sA + GB_STAGES * GB_A_BYTES;
  uint8_t* genC = gen + b_bytes + GB_STAGES * GB_A_BYTES;
  float* bias_s = reinterpret_cast<float*>(genC + GB_C_BYTES);  // [BN]
  uint8_t* ctl = genC + GB_C_BYTES + 256 * sizeof(float);
  const uint32_t bar0 = smem_u32(ctl);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (GB_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * GB_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * GB_MAX_STAGES + 2 + a); };
  const uint32_t bfull_bar = bar0 + 8u * (2 * GB_MAX_STAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * (2 * GB_MAX_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / GB_BN, m_blocks = (M + GB_BM - 1) / GB_BM, k_blocks = K / GB_BK;
  const int n_blk = blockIdx.x % n_blocks, n0 = n_blk * GB_BN;
  const int m_first = blockIdx.x / n_blocks, m_step = gridDim.x / n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < GB_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int et = (warp - 2) * 32 + lane;  // 0..255
    if (et < GB_BN) bias_s[et] = __ldg(bias + n0 + et);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull_bar, (uint32_t)k_blocks * GB_B_ATOM);
      for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d(sB + kb * GB_B_ATOM, &tmB, kb * GB_BK, n0, bfull_bar);
      int stage = 0; uint32_t phase = 0;
      for (int mb = m_first; mb < m_blocks; mb += m_step) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), GB_A_BYTES);
          tma_load_2d(sA + stage * GB_A_BYTES, &tmA, kb * GB_BK, mb * GB_BM, full_bar(stage));
          if (++stage == GB_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GB_BM, GB_BN);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      mbar_wait(bfull_bar, 0);
      for (int mb = m_first; mb < m_blocks; mb += m_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * GB_BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < GB_BK / 16; ++kk) {
            const uint64_t da = umma_desc_sw128(sA + stage * GB_A_BYTES + kk * 32);
            const uint64_t db = umma_desc_sw128(sB + kb * GB_B_ATOM + kk * 32);
            umma_bf16(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the A slot when these MMAs retire
          if (++stage == GB_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // epilogue: 8 warps = 2 groups of 4; warp w owns TMEM lane quarter (w % 4), group g owns columns [128 g, 128 g + 128)
    // of the tile and its own 16 KB staging atom.  Each 64-column pass is converted to bf16 into the swizzled staging
    // atom and written with one TMA store (full 128-byte lines); per-lane 16-byte global stores to 32 different rows
    // would cost 32 L1 wavefronts per instruction.
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int rt = quarter * 32 + lane;  // row inside the tile
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    uint8_t* cst = genC + grp * (GB_BM * 128);
    const uint32_t cst_s = sC + grp * (GB_BM * 128);
    int acc = 0; uint32_t acc_phase = 0;
    for (int mb = m_first; mb < m_blocks; mb += m_step) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * GB_BN + grp * GCOLS;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
#pragma unroll
      for (int ch = 0; ch < 2 * PASSES; ++ch) {
        if ((ch & 1) == 0) {
          // the group's staging atom must have been drained by its previous TMA store
          if (issuer) tma_store_wait_read();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
        tmem_ld_wait();
        if (ch + 1 < 2 * PASSES) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);  // next slab in flight
        const uint32_t* rc = r[ch & 1];
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v0 = __uint_as_float(rc[2 * j]) + bias_s[grp * GCOLS + ch * 32 + 2 * j];
          const float v1 = __uint_as_float(rc[2 * j + 1]) + bias_s[grp * GCOLS + ch * 32 + 2 * j + 1];
          __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
          o[j] = *reinterpret_cast<uint32_t*>(&p);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t cidx = (uint32_t)((ch & 1) * 4 + q);  // 16-byte chunk inside the 64-column pass
          uint8_t* dstp = BLOCKED ? cst + cidx * 2048u + (uint32_t)rt * 16u : cst + sw128_chunk_off((uint32_t)rt, cidx);
          *reinterpret_cast<uint4*>(dstp) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        if ((ch & 1) == 1) {
          if (ch == 2 * PASSES - 1) {  // all TMEM reads of this accumulator (by this thread) are done
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (issuer) {
            const int col = n0 + grp * GCOLS + (ch >> 1) * 64;
            if (BLOCKED) bulk_store_s2g(Cblk + ((size_t)mb * N + (size_t)col) * GB_BM, cst_s, 16384u);
            else tma_store_2d(&tmC, cst_s, col, mb * GB_BM);
            tma_store_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN>
static int launch_proj_gemm_bn(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, __nv_bfloat16* C, int M, int N,
                               int K, bool blocked, cudaStream_t st) {
  CUtensorMap tmC;
  {
    // (for the blocked layout the map is unused; encode a valid one over the same buffer anyway)
    int rc0 = make_tmap_bf16(&tmC, C, (uint64_t)M, (uint64_t)N, 64, GB_BM);
    if (rc0) return rc0;
  }
  BCI_REQUIRE(N % BN == 0 && K % GB_BK == 0 && K <= GbCfg<BN>::MAX_K && M > 0 && N / BN <= sm_count(), BCI_EINVAL,
              "proj_gemm_bf16: unsupported shape M=%d N=%d K=%d", M, N, K);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, 64, GB_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, 64, BN);
  if (rc) return rc;
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    BCI_CUDA_OK(cudaFuncSetAttribute(proj_gemm_bf16<false, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gb_smem(GbCfg<BN>::MAX_K, BN)));
    BCI_CUDA_OK(cudaFuncSetAttribute(proj_gemm_bf16<true, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gb_smem(GbCfg<BN>::MAX_K, BN)));
    attr = true;
  }
  const int n_blocks = N / BN, m_blocks = ceil_div(M, GB_BM);
  int per_n = sm_count() / n_blocks;
  if (per_n > m_blocks) per_n = m_blocks;
  const int grid = per_n * n_blocks;
  if (blocked) proj_gemm_bf16<true, BN><<<grid, GB_THREADS, gb_smem(K, BN), st>>>(tmA, tmB, tmC, C, bias, M, N, K, gb_stages(K));
  else proj_gemm_bf16<false, BN><<<grid, GB_THREADS, gb_smem(K, BN), st>>>(tmA, tmB, tmC, C, bias, M, N, K, gb_stages(K));
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int launch_proj_gemm_bf16(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, __nv_bfloat16* C, int M, int N,
                          int K, bool blocked, cudaStream_t st) {
  // large products: CTA pairs, 256 x 256 tiles (lstm_bf16_gemm_pair.cu) -- half the L2 -> SM operand bytes per flop
  const int rc_pair = launch_proj_gemm_bf16_pair(A, W, bias, C, M, N, K, blocked, st);
  if (rc_pair != 1) return rc_pair;
  if (K <= 256 && N % 256 == 0) return launch_proj_gemm_bn<256>(A, W, bias, C, M, N, K, blocked, st);
  return launch_proj_gemm_bn<128>(A, W, bias, C, M, N, K, blocked, st);
}

// ---------------------------------------------------------------------------------------------
// K3: persistent tcgen05 recurrence (H = 128)
// warps 0-7: epilogue (thread = window row x half of the hidden units); warp 8: MMA issuer + TMEM owner + TMA stores of
// h_t; warp 9: L2 prefetcher for the next step's G block.
// G arrives in the blocked streaming layout written by proj_gemm_bf16<true>: for a fixed 16-byte chunk (8 units of one
// gate) the 128 rows of an m-block are contiguous, so the per-thread loads of a warp (lane = row) coalesce into 512-byte
// requests.  History (profiles/): v1 per-thread loads from row-major G: 62 % long-scoreboard stalls (32 L1 wavefronts
// per load); v2 G through a 4-slot TMA smem ring: smem bandwidth became the limiter (MMA operands 192 KB + G 256 KB + h
// 64 KB per step against 128 B/clk); v3 (this): no G traffic through smem, h double-buffered so the TMA store of h_{t-1}
// never delays the epilogue of step t.
// ---------------------------------------------------------------------------------------------
constexpr int RB_H = 128, RB_M = 128, RB_N = 4 * RB_H;  // 512 gate columns per direction
constexpr int RB_EPI_WARPS = 8, RB_THREADS = (RB_EPI_WARPS + 2) * 32;
constexpr uint32_t RB_W_BYTES = RB_N * RB_H * 2;   // 131072: two K-atoms of [512][64]
constexpr uint32_t RB_W_ATOM = RB_N * 128;         // 65536
constexpr uint32_t RB_H_BYTES = RB_M * RB_H * 2;   // 32768: two K-atoms of [128][64]
constexpr uint32_t RB_H_ATOM = RB_M * 128;         // 16384
constexpr size_t RB_SMEM = 1024 + RB_W_BYTES + 2 * RB_H_BYTES + 128;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x already carries the 1/2 of sigmoid(2x') = 0.5 + 0.5 tanh(x') (folded into the weights at pack time)
__device__ __forceinline__ float sigmoid_half_arg(float xh) { return fmaf(0.5f, tanh_fast(xh), 0.5f); }
// two activations per MUFU op: tanh.approx.f16x2 (max rel. error 2^-10.99, below the bf16 rounding of h)
__device__ __forceinline__ __half2 tanh2_f16(float a, float b) {
  __half2 x = __floats2half2_rn(a, b);
  uint32_t xi = *reinterpret_cast<uint32_t*>(&x), yi;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(yi) : "r"(xi));
  return *reinterpret_cast<__half2*>(&yi);
}

// 10 warps put 3 on one SM sub-partition (16 K registers each): 3 x 32 x 168 is the per-thread ceiling, hence the few
// spilled words ptxas reports; 200 registers per thread compile but do not launch.
template <bool STATS, bool F16ACT>
__global__ void __launch_bounds__(RB_THREADS, 1)
lstm_rec_bf16(const __nv_bfloat16* __restrict__ G,        // blocked: [row/128][dir*64 + chunk][row%128][8], bias included
              const __grid_constant__ CUtensorMap tmOut,  // out [T][Bc][256] bf16 (3D), box 64 cols x 128 rows x 1
              const __nv_bfloat16* __restrict__ whh_f,    // [512][128] rows in perm_T order, forward
              const __nv_bfloat16* __restrict__ whh_r,    // reverse
              float2* __restrict__ stats,                 // STATS: [T][dir*4 + half*2 + k][Bc] (sum, sum of squares) of h over 32 units
              int Bc, int T, int dbg) {
  extern __shared__ uint8_t rb_smem_raw[];
  const uint32_t raw = smem_u32(rb_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = rb_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + RB_W_BYTES;       // sH: two h buffers of RB_H_BYTES
  uint8_t* genW = gen;
  uint8_t* genH = gen + RB_W_BYTES;
  uint8_t* ctl = genH + 2 * RB_H_BYTES;
  const uint32_t bar_half0 = smem_u32(ctl), bar_half1 = bar_half0 + 8, bar_h = bar_half0 + 16, bar_hfree = bar_half0 + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * RB_M;

  // stage W_hh into the SW128 K-major UMMA layout; zero both h buffers (h_{-1} = 0 lives in buffer 1)
  {
    const uint4* src = reinterpret_cast<const uint4*>(dir ? whh_r : whh_f);  // 16 chunks of 16 B per row
    for (int q = tid; q < RB_N * 16; q += RB_THREADS) {
      const uint32_t row = q >> 4, cc = q & 15, atom = cc >> 3, c = cc & 7;
      *reinterpret_cast<uint4*>(genW + atom * RB_W_ATOM + sw128_chunk_off(row, c)) = __ldg(src + q);
    }
    for (int q = tid; q < (int)(2 * RB_H_BYTES / 16); q += RB_THREADS) reinterpret_cast<uint4*>(genH)[q] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    mbar_init(bar_half0, 1);
    mbar_init(bar_half1, 1);
    mbar_init(bar_h, RB_EPI_WARPS * 32);
    mbar_init(bar_hfree, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmOut);
  }
  if (warp == RB_EPI_WARPS) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // byte address of chunk 0 of (row, dir) in the blocked G layout
  auto g_row_ptr = [&](long long row) {
    return reinterpret_cast<const uint8_t*>(G) + (((row >> 7) * 2 + dir) * 64) * 2048ll + (row & 127) * 16ll;
  };

  if (warp == RB_EPI_WARPS + 1) {
    // ---------------- L2 prefetcher: the G block of the NEXT step (128 rows x 512 columns = 128 KB per direction) ----------------
    if (!(dbg & 1)) {
      for (int s = 0; s < T; ++s) {
        const int sn = s + 1;
        if (sn >= T) break;
        const long long row0 = (long long)(dir ? (T - 1 - sn) : sn) * Bc + b0;
        // the tile's rows live in one m-block when Bc % 128 == 0, else in two consecutive ones
        const uint8_t* blk0 = reinterpret_cast<const uint8_t*>(G) + (((row0 >> 7) * 2 + dir) * 64) * 2048ll;
        if (lane < 8) bulk_prefetch_l2(blk0 + lane * 16384, 16384u);
        if ((row0 & 127) != 0 && lane >= 8 && lane < 16) bulk_prefetch_l2(blk0 + 2 * 64 * 2048ll + (lane - 8) * 16384, 16384u);
        // pace: one step of prefetch per step of compute (bar_h completes once per step)
        mbar_wait(bar_h, (uint32_t)(s & 1));
      }
    }
  } else if (warp == RB_EPI_WARPS) {
    // ---------------- MMA issuer (also streams h_t to global with TMA stores) ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(RB_M, 256);
      for (int s = 0; s <= T; ++s) {
        if (s > 0) {
          mbar_wait(bar_h, (uint32_t)((s - 1) & 1));  // h_{s-1} written, TMEM drained
          tc_fence_after();
        }
        const uint32_t hprev = sH + (uint32_t)((s + 1) & 1) * RB_H_BYTES;  // h_{s-1} lives in buffer (s-1)&1
        if (s < T) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int k = 0; k < ((dbg & 4) ? 1 : RB_H / 16); ++k) {
              const uint32_t atom = k >> 2, kk = k & 3;
              const uint64_t da = umma_desc_sw128(hprev + atom * RB_H_ATOM + kk * 32);
              const uint64_t db = umma_desc_sw128(sW + atom * RB_W_ATOM + half * (256 * 128) + kk * 32);
              umma_bf16(tmem_base + half * 256, da, db, idesc, k != 0 ? 1u : 0u);
            }
            umma_commit(half == 0 ? bar_half0 : bar_half1);
          }
        }
        if (s > 0) {
          // h_{s-1} sits in its operand buffer as two [128 x 64] SW128 atoms == two TMA store boxes.  The other buffer
          // (being written by the epilogue of step s) was last read by the store issued one iteration ago: allow one
          // store group in flight and release that buffer.
          if (!(dbg & 2)) {
            const int tp = dir ? (T - s) : (s - 1);
            tma_store_3d(&tmOut, hprev, dir * 128, b0, tp);
            tma_store_3d(&tmOut, hprev + RB_H_ATOM, dir * 128 + 64, b0, tp);
            tma_store_commit();
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          mbar_arrive(bar_hfree);
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ---------------- epilogue: thread = (window row, half of the hidden units) ----------------
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;  // window row inside the tile == TMEM lane
    const bool live = b0 + r < Bc;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)half * 256;
    const uint32_t my_bar = half == 0 ? bar_half0 : bar_half1;
    float c[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) c[i] = 0.f;

    for (int s = 0; s < T; ++s) {
      const int t = dir ? (T - 1 - s) : s;
      // chunk (sl*8 + half*4 + gate) of this thread's row: gates i,f,g,o of units 64*half + 8*sl .. +7
      const uint8_t* gp = g_row_ptr((long long)t * Bc + (live ? b0 + r : b0)) + (half * 4) * 2048;
      // G slabs are loaded two slabs ahead (gbuf[sl % 3]); slabs 0 and 1 are issued before waiting on the MMA
      uint4 gbuf[3][4];
      if (!(dbg & 1)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) gbuf[0][q] = ldg_stream_v4(gp + q * 2048);
#pragma unroll
        for (int q = 0; q < 4; ++q) gbuf[1][q] = ldg_stream_v4(gp + (8 + q) * 2048);
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) gbuf[i][q] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
      }
      uint8_t* hrow = genH + (uint32_t)(s & 1) * RB_H_BYTES + half * RB_H_ATOM;  // h_s -> buffer s&1, K-atom `half`
      mbar_wait(my_bar, (uint32_t)(s & 1));
      tc_fence_after();
      float ssum = 0.f, ssq = 0.f, ssum_lo = 0.f, ssq_lo = 0.f;
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {  // fully unrolled: c[] must stay in registers
        if (STATS && sl == 4) { ssum_lo = ssum; ssq_lo = ssq; ssum = 0.f; ssq = 0.f; }  // 32-unit partials (shared layout with the fused kernel)
        uint32_t acc[32];
        tmem_ld32(taddr + sl * 32, acc);
        if (sl < 6 && !(dbg & 1)) {
#pragma unroll
          for (int q = 0; q < 4; ++q) gbuf[(sl + 2) % 3][q] = ldg_stream_v4(gp + ((sl + 2) * 8 + q) * 2048);
        }
        tmem_ld_wait();
        // gbuf[sl % 3][g] = 8 bf16 of gate g (units u = 0..7): element u sits in word u/2, low or high half
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(gbuf[sl % 3]);
        uint32_t hp[4];
#pragma unroll
        for (int u2 = 0; u2 < 4; ++u2) {
          float hv[2];
          auto gval = [&](int g, int u) {
            const uint32_t w = gw[g * 4 + (u >> 1)];
            return __uint_as_float((u & 1) ? (w & 0xFFFF0000u) : (w << 16));
          };
          if (F16ACT) {
            const int u0 = u2 * 2, u1 = u2 * 2 + 1;
            const __half2 half_h2 = __float2half2_rn(0.5f);
            const float2 ig = __half22float2(__hfma2(tanh2_f16(__uint_as_float(acc[0 * 8 + u0]) + gval(0, u0),
                                                               __uint_as_float(acc[0 * 8 + u1]) + gval(0, u1)), half_h2, half_h2));
            const float2 fg = __half22float2(__hfma2(tanh2_f16(__uint_as_float(acc[1 * 8 + u0]) + gval(1, u0),
                                                               __uint_as_float(acc[1 * 8 + u1]) + gval(1, u1)), half_h2, half_h2));
            const float2 gg = __half22float2(tanh2_f16(__uint_as_float(acc[2 * 8 + u0]) + gval(2, u0),
                                                       __uint_as_float(acc[2 * 8 + u1]) + gval(2, u1)));
            const float2 og = __half22float2(__hfma2(tanh2_f16(__uint_as_float(acc[3 * 8 + u0]) + gval(3, u0),
                                                               __uint_as_float(acc[3 * 8 + u1]) + gval(3, u1)), half_h2, half_h2));
            float& c0 = c[sl * 8 + u0];
            float& c1 = c[sl * 8 + u1];
            c0 = fmaf(fg.x, c0, ig.x * gg.x);
            c1 = fmaf(fg.y, c1, ig.y * gg.y);
            const float2 tc = __half22float2(tanh2_f16(c0, c1));
            hv[0] = og.x * tc.x;
            hv[1] = og.y * tc.y;
            if (STATS) { ssum += hv[0] + hv[1]; ssq = fmaf(hv[0], hv[0], fmaf(hv[1], hv[1], ssq)); }
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int u = u2 * 2 + e;
              const float ig = sigmoid_half_arg(__uint_as_float(acc[0 * 8 + u]) + gval(0, u));
              const float fg = sigmoid_half_arg(__uint_as_float(acc[1 * 8 + u]) + gval(1, u));
              const float gg = tanh_fast(__uint_as_float(acc[2 * 8 + u]) + gval(2, u));
              const float og = sigmoid_half_arg(__uint_as_float(acc[3 * 8 + u]) + gval(3, u));
              float& cc = c[sl * 8 + u];
              cc = fmaf(fg, cc, ig * gg);
              hv[e] = og * tanh_fast(cc);
              if (STATS) { ssum += hv[e]; ssq = fmaf(hv[e], hv[e], ssq); }
            }
          }
          __nv_bfloat162 p = __floats2bfloat162_rn(hv[0], hv[1]);
          hp[u2] = *reinterpret_cast<uint32_t*>(&p);
        }
        const uint4 hvec = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        // buffer s&1 was last read by the TMA store of h_{s-2}; the MMA thread's arrival of iteration s certifies it is done
        if (sl == 0 && s > 0) mbar_wait(bar_hfree, (uint32_t)((s - 1) & 1));
        // units 64*half + 8*sl .. +7  ->  chunk sl of row r in K-atom `half`
        *reinterpret_cast<uint4*>(hrow + sw128_chunk_off((uint32_t)r, (uint32_t)sl)) = hvec;
      }
      fence_proxy_async_smem();  // h_t (generic-proxy stores) -> visible to the next step's tcgen05.mma
      tc_fence_before();         // order this thread's TMEM reads before the barrier
      mbar_arrive(bar_h);
      if (STATS) {
        // partial LayerNorm statistics of the last layer's output row (consumed by attn_score_bf16 / attn_pool_finish)
        if (live) {  // [T][8][Bc]: a warp's 32 windows are 256 contiguous bytes per slot (full sectors, no DRAM read-modify-write)
          float2* sp = stats + ((long long)t * 8 + dir * 4 + half * 2) * Bc + b0 + r;
          sp[0] = make_float2(ssum_lo, ssq_lo);
          sp[Bc] = make_float2(ssum, ssq);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RB_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// activation variant of the recurrence epilogue: 0 = tanh.approx.f32 (default), 1 = tanh.approx.f16x2 (BCI_REC_ACT=f16).
// The packed form was tried to halve the MUFU load, but ptxas splits it into two MUFU.TANH.F16 (same MUFU count, measured
// no faster, slightly less accurate) -- kept only as an experiment switch.
static int rec_act_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("BCI_REC_ACT");
    mode = (e && e[0] == 'f' && e[1] == '1') ? 1 : 0;  // "f16" -> 1; default tanh.approx.f32
  }
  return mode;
}
static int rec_dbg() {  // timing experiments only (results are wrong when set): 1 = no G traffic, 2 = no h TMA store, 4 = no MMA
#ifdef BCI_DEBUG_SWITCHES  // compiled out of the shipped library: a stray environment variable must not corrupt inference
  static int v = -1;
  if (v < 0) { const char* e = getenv("BCI_REC_DBG"); v = e ? atoi(e) : 0; }
  return v;
#else
  return 0;
#endif
}

int launch_rec_bf16(const __nv_bfloat16* G, const __nv_bfloat16* whh_f, const __nv_bfloat16* whh_r, __nv_bfloat16* out,
                    float2* stats, int Bc, int T, cudaStream_t st) {
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_bf16<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_bf16<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_bf16<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_bf16<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM));
    attr = true;
  }
  CUtensorMap tmOut;
  int rc = make_tmap_bf16_3d(&tmOut, out, (uint64_t)T, (uint64_t)Bc, 256, 64, RB_M);
  if (rc) return rc;
  dim3 grid(ceil_div(Bc, RB_M), 2);
  if (rec_act_mode()) {
    if (stats) lstm_rec_bf16<true, true><<<grid, RB_THREADS, RB_SMEM, st>>>(G, tmOut, whh_f, whh_r, stats, Bc, T, rec_dbg());
    else lstm_rec_bf16<false, true><<<grid, RB_THREADS, RB_SMEM, st>>>(G, tmOut, whh_f, whh_r, nullptr, Bc, T, rec_dbg());
  } else {
    if (stats) lstm_rec_bf16<true, false><<<grid, RB_THREADS, RB_SMEM, st>>>(G, tmOut, whh_f, whh_r, stats, Bc, T, rec_dbg());
    else lstm_rec_bf16<false, false><<<grid, RB_THREADS, RB_SMEM, st>>>(G, tmOut, whh_f, whh_r, nullptr, Bc, T, rec_dbg());
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
static size_t chunk_bytes_bf16(const bci_lstm_config& c, int Bc, int T) {
  const size_t H = c.hidden_size, rows = (size_t)Bc * T;
  const size_t rows_pad = (rows + 127) / 128 * 128;  // G is stored in whole 128-row blocks (split path only)
  const size_t g_bytes = bf16_path_fused() ? 0 : align_up(rows_pad * 8 * H * 2, 1024);
  return align_up(rows * H * 2, 1024) + g_bytes + 2 * align_up(rows * 2 * H * 2, 1024) +
         align_up(rows * 4, 1024) + align_up(rows * 8 * sizeof(float2), 1024);
}

// Batches up to this many windows do not fill one wave of the tile-per-CTA tcgen05 kernels (128 windows per CTA): they run the
// swapped recurrence (8 windows per CTA, W_hh in tensor memory: lstm_rec_swap.cu) in its one-chain fp16 form, fed by single-pass
// TF32 projections, through the fp32 path's small-batch code.  BCI_BF16_SMALL=off keeps the tile kernels for them.
// (measured: 256 windows 3.10 -> 1.21 ms, 512 windows 3.12 -> 1.78 ms, single window 3.04 -> 0.72 ms; from one wave of the swapped
// kernel -- 148 CTAs x 8 windows / 2 directions -- upwards the 3.1 ms latency of the fused cluster kernel is the shorter one)
constexpr int BF16_SMALL_BATCH = 592;
static bool bf16_small_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_BF16_SMALL");
    v = (e && e[0] == 'o') ? 0 : 1;
  }
  return v != 0 && swap_rec_enabled();
}

size_t lstm_workspace_bf16(const bci_lstm_config& c, int batch, int T) {
  if (c.hidden_size == 256) return lstm_workspace_h256(c, batch, T);
  const int Bc = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  size_t n = chunk_bytes_bf16(c, Bc > 0 ? Bc : 1, T) + 1024;  // + slack: the TMA-addressed buffers are aligned to 1 KB internally
  if (batch <= BF16_SMALL_BATCH && bf16_small_enabled()) {
    const size_t m = lstm_chunk_bytes_f32(c, batch > 0 ? batch : 1, T);
    if (m > n) n = m;
  }
  return n;
}

static int forward_chunk_bf16(bci_lstm_s* h, const InputView& x, int Bc, int T, float* logits, float* probs, float* attn, char* ws,
                              cudaStream_t st) {
  constexpr int H = 128;
  const bci_lstm_config& c = h->cfg;
  const size_t rows = (size_t)Bc * T;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = ws + off; off += align_up(bytes, 1024); return p; };
  __nv_bfloat16* z = reinterpret_cast<__nv_bfloat16*>(take(rows * H * 2));
  const bool fused = bf16_path_fused() != 0;
  __nv_bfloat16* g = fused ? nullptr : reinterpret_cast<__nv_bfloat16*>(take((rows + 127) / 128 * 128 * 8 * H * 2));
  __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(take(rows * 2 * H * 2));
  __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(take(rows * 2 * H * 2));
  float* scores = reinterpret_cast<float*>(take(rows * 4));
  float2* stats = reinterpret_cast<float2*>(take(rows * 8 * sizeof(float2)));
  h->prof.mark(-1, st);
  // tensor-core input projection needs whole 128-row tiles inside one window; other lengths use the CUDA-core kernel
  int rc = input_proj_bf16_ok(h, x, T) ? launch_input_proj_bf16(h, x, Bc, T, z, st) : launch_input_proj<H, __nv_bfloat16>(h, x, Bc, T, z, st);
  if (rc) return rc;
  h->prof.mark(0, st);
  const __nv_bfloat16* in = z;
  __nv_bfloat16* outs[2] = {o0, o1};
  for (int l = 0; l < c.num_layers; ++l) {
    __nv_bfloat16* o = outs[l & 1];
    float2* st_l = l == c.num_layers - 1 ? stats : nullptr;
    if (fused) {
      // projection and recurrence in one cluster kernel (phase 2 of the profile; phase 1 stays empty)
      rc = launch_fused_rec_bf16(in, h->bf16.wih_t_bf[l], h->bf16.whh_bf[l][0], h->bf16.whh_bf[l][1], h->bf16.bias_t[l], o, st_l, Bc, T,
                                 layer_in_width(c, l), st);
      if (rc) return rc;
    } else {
      rc = launch_proj_gemm_bf16(in, h->bf16.wih_bf[l], h->bf16.bias_p[l], g, (int)rows, 8 * H, layer_in_width(c, l), true, st);
      if (rc) return rc;
      h->prof.mark(1, st);
      rc = launch_rec_bf16(g, h->bf16.whh_bf[l][0], h->bf16.whh_bf[l][1], o, st_l, Bc, T, st);
      if (rc) return rc;
    }
    h->prof.mark(2, st);
    in = o;
  }
  // single pass over the sequence when the batch fills the machine (the score buffer then holds the pooled context)
  rc = pool_stream_ok(h, Bc, T) ? launch_pool_stream_bf16(h, in, stats, scores, Bc, T, logits, probs, attn, st)
                                : launch_pool_bf16(h, in, stats, scores, Bc, T, logits, probs, attn, st);
  h->prof.mark(3, st);
  return rc;
}

int lstm_forward_bf16(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  if (c.hidden_size == 256) return lstm_forward_h256(h, x, batch, T, logits, probs, attn, ws, ws_bytes, st);
  BCI_REQUIRE(c.hidden_size == 128, BCI_EINVAL, "bf16 mode supports hidden_size 128 and 256");
  if (batch <= BF16_SMALL_BATCH && bf16_small_enabled()) {
    h->infer_fast = true;
    const int rc = lstm_forward_fp32(h, x, batch, T, logits, probs, attn, ws, ws_bytes, st);
    h->infer_fast = false;
    return rc;
  }
  const int chunk = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  char* ws_al = reinterpret_cast<char*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  BCI_REQUIRE(ws_bytes >= chunk_bytes_bf16(c, chunk, T) + (size_t)(ws_al - (char*)ws), BCI_ENOMEM,
              "bci_lstm_forward: workspace %zu < %zu bytes", ws_bytes, chunk_bytes_bf16(c, chunk, T) + 1024);
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int Bc = (batch - b0) < chunk ? (batch - b0) : chunk;
    int rc = forward_chunk_bf16(h, chunk_view(x, b0), Bc, T, logits + (size_t)b0 * c.num_classes,
                                probs ? probs + (size_t)b0 * c.num_classes : nullptr, attn ? attn + (size_t)b0 * T : nullptr,
                                ws_al, st);
    if (rc) return rc;
  }
  return BCI_OK;
}

}  // namespace bci

// ---- diagnostics exported for the GPU unit tests (tests/test_gpu_tensorcore.py) ---------------------
extern "C" int bci_selftest_proj_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N,
                                           int32_t K, void* stream) {
  return bci::launch_proj_gemm_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)W, bias, (__nv_bfloat16*)C, M, N, K, false,
                                    (cudaStream_t)stream);
}
// the same product written in the recurrence's blocked streaming layout [row / 128][n / 8][row % 128][8] (C holds ceil(M / 128) row blocks)
extern "C" int bci_selftest_proj_gemm_bf16_blocked(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N,
                                                   int32_t K, void* stream) {
  return bci::launch_proj_gemm_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)W, bias, (__nv_bfloat16*)C, M, N, K, true,
                                    (cudaStream_t)stream);
}
extern "C" int bci_selftest_rec_bf16(const void* G, const void* whh_f, const void* whh_r, void* out, int32_t Bc, int32_t T,
                                     void* stream) {
  return bci::launch_rec_bf16((const __nv_bfloat16*)G, (const __nv_bfloat16*)whh_f, (const __nv_bfloat16*)whh_r,
                              (__nv_bfloat16*)out, nullptr, Bc, T, (cudaStream_t)stream);
}
