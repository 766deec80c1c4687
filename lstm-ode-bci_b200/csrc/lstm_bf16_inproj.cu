// bf16 mode of K1: z = GELU(LayerNorm(x W0^T + b0))   (04_lstm_model.py:173-178,208) on tcgen05.
//
// x is fp32 (B,T,C) with C = 61 channels: rows are 244 bytes, which no TMA tensor map can describe
// (strides must be multiples of 16 B).  But a tile of 128 consecutive (b,t) rows is one CONTIGUOUS block of
// 128*C*4 bytes, so it is fetched with a 1-D bulk copy (cp.async.bulk, mbarrier-completed) into a raw
// staging buffer.  The windows need not be packed: an InputView (lstm_handle.cuh) gives the element offset of
// every window, so the 50 %-overlapping windows of a (samples, C) recording (02_preprocessing.py:157-180) are
// read in place -- half the HBM and host->device bytes of materialised windows -- and the elements may already
// be bf16 (the converter warps round fp32 input to bf16 anyway, so bf16 input gives bit-identical results at
// half the bytes again); four converter warps then turn it into the bf16 K-major SWIZZLE_128B A operand (K zero-padded
// 61 -> 64), one thread issues 4 tcgen05.mma (M128 x N128 x K16), and four epilogue warps apply bias +
// LayerNorm + GELU thread-locally (one thread owns one row's 128 accumulator columns in TMEM) and write
// the bf16 tile back time-major with TMA stores.  Every stage is double-buffered.
//
// The first version of K1 (one warp per row on CUDA cores, lstm_shared_kernels.cuh) took 6.1 ms per
// 18944-window step -- 26 % of the whole forward; it remains the path for sequence lengths that are not a
// multiple of 128 and for fp32 mode.
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"

namespace bci {
using namespace sm100;

__global__ void pack_w0_bf16_kernel(const float* __restrict__ w0, __nv_bfloat16* __restrict__ dst, int H, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * 64) return;
  const int j = i >> 6, k = i & 63;
  dst[i] = __float2bfloat16_rn(k < C ? w0[j * C + k] : 0.f);
}
__global__ void pack_par0_kernel(const float* __restrict__ b0, const float* __restrict__ lnw, const float* __restrict__ lnb,
                                 float4* __restrict__ par, int H) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) par[j] = make_float4(b0[j], lnw[j], lnb[j], 0.f);
}

int pack_inproj_bf16(bci_lstm_s* h, cudaStream_t st) {
  const int H = h->cfg.hidden_size, C = h->cfg.input_size;
  const bci_lstm_weights& w = h->raw;
  pack_w0_bf16_kernel<<<ceil_div(H * 64, 256), 256, 0, st>>>(w.input_proj_w, h->bf16.w0_bf, H, C);
  pack_par0_kernel<<<1, 128, 0, st>>>(w.input_proj_b, w.input_ln_w, w.input_ln_b, h->bf16.par0, H);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

constexpr int IP_M = 128, IP_N = 128, IP_K = 64, IP_THREADS = 448;  // producer, MMA, 4 converter warps, 8 epilogue warps
constexpr uint32_t IP_A_BYTES = IP_M * 128;       // 16 KB, one SW128 atom
constexpr uint32_t IP_B_BYTES = IP_N * 128;       // 16 KB
constexpr uint32_t IP_OUT_BYTES = 2 * IP_M * 128; // two [128][64] bf16 atoms
constexpr uint32_t IP_RAW_MAX = IP_M * 64 * 4;    // 32 KB per raw buffer (C <= 64)
constexpr size_t IP_SMEM = 1024 + 2 * IP_RAW_MAX + 2 * IP_A_BYTES + IP_B_BYTES + IP_OUT_BYTES + IP_N * sizeof(float4) + 2 * IP_M * sizeof(float2) + 256;

// GELU in its tanh form, 0.5 y (1 + tanh(sqrt(2/pi) (y + 0.044715 y^3))): 1 MUFU + 6 FMA-pipe instructions.  It differs from the
// reference's erf form (04:176) by at most 5e-4 absolute -- a quarter of the bf16 rounding step of an output near 1 -- and the
// bf16-mode errors against the fp32 oracle are unchanged (tests/test_gpu_tensorcore.py prints them).  The erf form used before
// (Abramowitz-Stegun 7.1.26: 2 MUFU + ~12 FMA) made this kernel issue-bound at 0.95 ms per 16 896 windows; this one takes 0.58 ms.
__device__ __forceinline__ float gelu_tanh_fast(float y) {
  const float u = y * fmaf(0.0356774081f, y * y, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hy = 0.5f * y;
  return fmaf(hy, t, hy);
}

// the same on a pair of values (two scalar MUFUs, everything else packed)
__device__ __forceinline__ float2 gelu_tanh_fast2(float2 y) {
  const float2 u = __fmul2_rn(y, __ffma2_rn(make_float2(0.0356774081f, 0.0356774081f), __fmul2_rn(y, y), make_float2(0.7978845608f, 0.7978845608f)));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hy = __fmul2_rn(make_float2(0.5f, 0.5f), y);
  return __ffma2_rn(hy, t, hy);
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <typename InT>
__global__ void __launch_bounds__(IP_THREADS, 1)
input_proj_bf16(const InputView x,                         // windows of T x C elements of InT (fp32 or bf16)
                const __grid_constant__ CUtensorMap tmB,   // W0 bf16 [128][64], box 64 x 128
                const __grid_constant__ CUtensorMap tmZ,   // z [T][Bc][128] bf16 (3D), box 64 x 1 x 128
                const float4* __restrict__ par,            // [128] {b0, ln_w, ln_b, 0}
                int Bc, int T, int C) {
  extern __shared__ uint8_t ip_smem_raw[];
  const uint32_t raw_addr = smem_u32(ip_smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = ip_smem_raw + (base - raw_addr);
  // layout: [A0][A1][B][OUT][RAW0][RAW1][par][ctl]
  const uint32_t sA = base, sB = sA + 2 * IP_A_BYTES, sO = sB + IP_B_BYTES, sR = sO + IP_OUT_BYTES;
  uint8_t* genA = gen;
  uint8_t* genO = gen + 2 * IP_A_BYTES + IP_B_BYTES;
  const InT* genR = reinterpret_cast<const InT*>(genO + IP_OUT_BYTES);
  float4* par_s = reinterpret_cast<float4*>(gen + 2 * IP_A_BYTES + IP_B_BYTES + IP_OUT_BYTES + 2 * IP_RAW_MAX);
  float2* part_s = reinterpret_cast<float2*>(par_s + IP_N);  // [2 column halves][128 rows] partial (sum, sumsq)
  uint8_t* ctl = reinterpret_cast<uint8_t*>(part_s + 2 * IP_M);
  const uint32_t bar0 = smem_u32(ctl);
  auto raw_full = [&](int i) { return bar0 + 8u * i; };
  auto raw_empty = [&](int i) { return bar0 + 8u * (2 + i); };
  auto a_full = [&](int i) { return bar0 + 8u * (4 + i); };
  auto a_empty = [&](int i) { return bar0 + 8u * (6 + i); };
  auto t_full = [&](int i) { return bar0 + 8u * (8 + i); };
  auto t_empty = [&](int i) { return bar0 + 8u * (10 + i); };
  const uint32_t b_full = bar0 + 8u * 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_win = T / IP_M;
  const int tiles = Bc * tiles_per_win;
  const uint32_t tile_bytes = (uint32_t)IP_M * C * (uint32_t)sizeof(InT);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmZ);
    for (int i = 0; i < 2; ++i) {
      mbar_init(raw_full(i), 1); mbar_init(raw_empty(i), 128);
      mbar_init(a_full(i), 128); mbar_init(a_empty(i), 1);
      mbar_init(t_full(i), 1);   mbar_init(t_empty(i), 256);
    }
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  // parameters of column pair j = (2j, 2j+1), the unit of the packed-fp32 epilogue: par_s[2j] = {b0, b1, gamma0, gamma1}, par_s[2j+1] = {beta0, beta1, -, -}
  if (warp >= 6 && warp < 10) {
    const int col = (warp - 6) * 32 + lane;
    const float4 pc = __ldg(par + col);   // {bias, LN weight, LN bias, -} of one column
    float* pf = reinterpret_cast<float*>(par_s) + (col >> 1) * 8 + (col & 1);
    pf[0] = pc.x; pf[2] = pc.y; pf[4] = pc.z;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- producer: W0 once, then one contiguous raw fp32 tile per iteration ----------------
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, IP_B_BYTES);
      tma_load_2d(sB, &tmB, 0, 0, b_full);
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(raw_empty(s), ((it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(raw_full(s), tile_bytes);
        const int b = tile / tiles_per_win, part = tile - b * tiles_per_win;
        const InT* src = reinterpret_cast<const InT*>(x.data) + x.elem_off(b) + (long long)part * IP_M * C;
        bulk_copy_g2s(sR + s * IP_RAW_MAX, src, tile_bytes, raw_full(s));
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(IP_M, IP_N);
      mbar_wait(b_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1u;
        mbar_wait(a_full(s), ph);
        mbar_wait(t_empty(s), ph ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < IP_K / 16; ++kk)
          umma_bf16(tmem_base + s * IP_N, umma_desc_sw128(sA + s * IP_A_BYTES + kk * 32), umma_desc_sw128(sB + kk * 32), idesc,
                    kk != 0 ? 1u : 0u);
        umma_commit(a_empty(s));
        umma_commit(t_full(s));
      }
    }
  } else if (warp < 6) {
    // ---------------- converters: raw fp32 row -> bf16 SW128 K-major row (K padded to 64) ----------------
    const int r = (warp - 2) * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1u;
      mbar_wait(raw_full(s), ph);
      mbar_wait(a_empty(s), ph ^ 1u);
      const InT* row = genR + s * (IP_RAW_MAX / (int)sizeof(InT)) + r * C;
      uint8_t* arow = genA + s * IP_A_BYTES;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = c * 8 + 2 * j;
          const float v0 = k < C ? to_f32<InT>(row[k]) : 0.f;
          const float v1 = (k + 1) < C ? to_f32<InT>(row[k + 1]) : 0.f;
          __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
          w[j] = *reinterpret_cast<uint32_t*>(&p);
        }
        *reinterpret_cast<uint4*>(arow + sw128_chunk_off((uint32_t)r, (uint32_t)c)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(a_full(s));
      mbar_arrive(raw_empty(s));
    }
  } else {
    // ---------------- epilogue: bias + LayerNorm + GELU; 8 warps, two threads per row (column halves) ----------------
    const int quarter = warp & 3;            // TMEM lane quarter of this warp
    const int grp = (warp - 6) >> 2;         // column half: columns [64 grp, 64 grp + 64) == staging atom grp
    const int r = quarter * 32 + lane;
    const bool issuer = (((warp - 6) & 3) == 0 && lane == 0);
    uint8_t* oatom = genO + grp * (IP_M * 128);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait(t_full(s), (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)s * IP_N + grp * 64;
      // the row's 64 accumulator columns of this half stay in registers for both passes
      uint32_t a[2][32];
      tmem_ld32(taddr, a[0]);
      tmem_ld32(taddr + 32, a[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(t_empty(s));               // TMEM accumulator free again
      // packed fp32 (FADD2 / FFMA2 / FMUL2) on column pairs: the accumulator registers of a pair are adjacent
      float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 pa = par_s[(grp * 32 + ch * 16 + j) * 2];
          const float2 v = __fadd2_rn(make_float2(__uint_as_float(a[ch][2 * j]), __uint_as_float(a[ch][2 * j + 1])), make_float2(pa.x, pa.y));
          a[ch][2 * j] = __float_as_uint(v.x);
          a[ch][2 * j + 1] = __float_as_uint(v.y);
          sum2 = __fadd2_rn(sum2, v);
          sq2 = __ffma2_rn(v, v, sq2);
        }
      const float sum = sum2.x + sum2.y, sq = sq2.x + sq2.y;
      part_s[grp * IP_M + r] = make_float2(sum, sq);
      // staging atom free? (this group's previous TMA store has finished reading it)
      if (issuer) tma_store_wait_read();
      asm volatile("bar.sync 2, 256;" ::: "memory");   // partial sums of both halves visible; staging free
      const float2 other = part_s[(grp ^ 1) * IP_M + r];
      const float mean = (sum + other.x) * (1.0f / IP_N);
      const float rstd = 1.0f / sqrtf(fmaxf((sq + other.y) * (1.0f / IP_N) - mean * mean, 0.f) + 1e-5f);
      const float2 nmean2 = make_float2(-mean, -mean), rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 pa = par_s[(grp * 32 + ch * 16 + j) * 2], pb = par_s[(grp * 32 + ch * 16 + j) * 2 + 1];
          const float2 v = make_float2(__uint_as_float(a[ch][2 * j]), __uint_as_float(a[ch][2 * j + 1]));
          const float2 y = __ffma2_rn(__fmul2_rn(__fadd2_rn(v, nmean2), rstd2), make_float2(pa.z, pa.w), make_float2(pb.x, pb.y));
          const float2 g = gelu_tanh_fast2(y);
          __nv_bfloat162 pk = __floats2bfloat162_rn(g.x, g.y);
          o[j] = *reinterpret_cast<uint32_t*>(&pk);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(oatom + sw128_chunk_off((uint32_t)r, (uint32_t)(ch * 4 + q))) =
              make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");   // both atoms written (also protects part_s for the next tile)
      if (issuer) {
        const int b = tile / tiles_per_win, t0 = (tile - b * tiles_per_win) * IP_M;
        tma_store_3d(&tmZ, sO + grp * (IP_M * 128), grp * 64, b, t0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// z [T][Bc][128] viewed as a 3-D tensor {128 cols, Bc, T}; a box {64, 1, 128} is one window's 128 time steps
static int make_tmap_z(CUtensorMap* tm, const void* z, uint64_t T, uint64_t Bc) {
  EncodeTiledFn enc = get_encode_fn();
  BCI_REQUIRE(enc, BCI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {128, Bc, T};
  cuuint64_t strides[2] = {128 * 2, Bc * 128 * 2};
  cuuint32_t box[3] = {64, 1, 128};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(z), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BCI_REQUIRE(r == CUDA_SUCCESS, BCI_ECUDA, "cuTensorMapEncodeTiled(z) failed with CUresult %d", (int)r);
  return BCI_OK;
}

// whole 128-row tiles inside one window, and every tile start 16-byte aligned (bulk-copy source); anything else runs the CUDA-core K1
bool input_proj_bf16_ok(const bci_lstm_s* h, const InputView& x, int T) {
  const long long es = x.esize();
  return T % IP_M == 0 && h->cfg.input_size <= 64 && h->cfg.hidden_size == 128 && ((uintptr_t)x.data & 15) == 0 &&
         (x.wstride * es) % 16 == 0 && (x.rstride * es) % 16 == 0 && ((long long)IP_M * h->cfg.input_size * es) % 16 == 0;
}

int launch_input_proj_bf16(bci_lstm_s* h, const InputView& x, int Bc, int T, __nv_bfloat16* z, cudaStream_t st) {
  const int C = h->cfg.input_size;
  BCI_REQUIRE(input_proj_bf16_ok(h, x, T), BCI_EINVAL, "input_proj_bf16: unsupported shape or alignment");
  CUtensorMap tmB, tmZ;
  int rc = make_tmap_bf16(&tmB, h->bf16.w0_bf, 128, 64, 64, IP_N);
  if (rc) return rc;
  rc = make_tmap_z(&tmZ, z, (uint64_t)T, (uint64_t)Bc);
  if (rc) return rc;
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    BCI_CUDA_OK(cudaFuncSetAttribute(input_proj_bf16<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IP_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(input_proj_bf16<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IP_SMEM));
    attr = true;
  }
  const int tiles = Bc * (T / IP_M);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (x.dtype == BCI_IN_BF16) input_proj_bf16<__nv_bfloat16><<<grid, IP_THREADS, IP_SMEM, st>>>(x, tmB, tmZ, h->bf16.par0, Bc, T, C);
  else input_proj_bf16<float><<<grid, IP_THREADS, IP_SMEM, st>>>(x, tmB, tmZ, h->bf16.par0, Bc, T, C);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci
