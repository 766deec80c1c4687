// bf16 mode of K4/K5: LayerNorm + additive-attention pooling + classifier (04_lstm_model.py:112-128,
// 192-204,212-218) with the score GEMM on tcgen05 and the sequence read from HBM exactly twice.
//
// LayerNorm is folded into the GEMM instead of being materialised:
//     y = (x - mean) rstd w + b                          (per row x of the last LSTM layer, 2H = 256 wide)
//     W1 y + b1 = rstd (x . W1' - mean s) + c            W1' = W1 diag(w),  s_j = sum_d W1'_jd,  c_j = b1_j + sum_d b_d W1_jd
//     ctx = sum_t a_t y_t = w (sum_t beta_t x_t - sum_t beta_t mean_t) + b          beta_t = a_t rstd_t   (sum_t a_t = 1)
// so both passes run on the raw bf16 rows.  Row sums / sums of squares come from the recurrence
// epilogue of the last layer (lstm_rec_bf16<STATS>), four partials per row.
//
//   attn_score_bf16   : persistent TMA + tcgen05 GEMM [T*Bc,256] x [256,128]; W1' resident in smem; the epilogue
//                       warps turn each 128-wide accumulator row into one score: sum_j w2_j tanh(rstd (acc_j - mean s_j) + c_j)
//   attn_pool_finish  : one CTA per window: softmax over T (block shuffles), beta-weighted row sum (coalesced
//                       bf16x2 stream), affine, classifier MLP, softmax -> logits / probs / attention.
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"

namespace bci {
using namespace sm100;

// ---- packing ---------------------------------------------------------------------------------
__global__ void pack_attn_w1_kernel(const float* __restrict__ w1, const float* __restrict__ lnw, __nv_bfloat16* __restrict__ dst, int H, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * D) return;
  dst[i] = __float2bfloat16_rn(w1[i] * lnw[i % D]);
}
__global__ void pack_attn_par_kernel(const float* __restrict__ w1, const __nv_bfloat16* __restrict__ w1p, const float* __restrict__ lnb,
                                     const float* __restrict__ b1, const float* __restrict__ w2, float4* __restrict__ par, int H, int D) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float s = 0.f, c = b1[j];
  for (int d = 0; d < D; ++d) {
    s += __bfloat162float(w1p[j * D + d]);
    c = fmaf(lnb[d], w1[j * D + d], c);
  }
  par[j] = make_float4(s, c, w2[j], 0.f);
}

int pack_pool_bf16(bci_lstm_s* h, cudaStream_t st) {
  const int H = h->cfg.hidden_size, D = 2 * H;
  const bci_lstm_weights& w = h->raw;
  pack_attn_w1_kernel<<<ceil_div(H * D, 256), 256, 0, st>>>(w.attn_w1, w.ln_w, h->bf16.aw1_bf, H, D);
  pack_attn_par_kernel<<<ceil_div(H, 128), 128, 0, st>>>(w.attn_w1, h->bf16.aw1_bf, w.ln_b, w.attn_b1, w.attn_w2, h->bf16.apar, H, D);
  BCI_LAUNCH_OK();
  // weight-only bound on |score| = |sum_j w2_j tanh(.)| for the single-pass pooling kernel (host value: one small copy per
  // load_weights of a bf16 handle)
  float4 parh[128];
  BCI_REQUIRE(H == 128, BCI_EINVAL, "pack_pool_bf16: hidden_size 128 expected");
  BCI_CUDA_OK(cudaMemcpyAsync(parh, h->bf16.apar, (size_t)H * sizeof(float4), cudaMemcpyDeviceToHost, st));
  BCI_CUDA_OK(cudaStreamSynchronize(st));
  float smax = 0.f;
  for (int j = 0; j < H; ++j) {
    smax += fabsf(parh[j].z);
    h->bf16.pool_par[0][j] = parh[j].x; h->bf16.pool_par[1][j] = parh[j].y; h->bf16.pool_par[2][j] = parh[j].z;
  }
  h->bf16.pool_smax = (smax == smax) ? fmaxf(smax, 1e-6f) : 0.f;  // NaN weights: keep the two-kernel path
  return BCI_OK;
}

// ---- scores on tensor cores ----------------------------------------------------------------------
constexpr int SC_BM = 128, SC_N = 128, SC_K = 256, SC_BK = 64, SC_STAGES = 8, SC_THREADS = 192;
constexpr uint32_t SC_A_BYTES = SC_BM * SC_BK * 2;  // 16 KB
constexpr uint32_t SC_B_ATOM = SC_N * SC_BK * 2;    // 16 KB
constexpr uint32_t SC_B_BYTES = (SC_K / SC_BK) * SC_B_ATOM;
constexpr size_t SC_SMEM = 1024 + SC_B_BYTES + (size_t)SC_STAGES * SC_A_BYTES + SC_N * sizeof(float4) + 256;

__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(SC_THREADS, 1)
attn_score_bf16(const __grid_constant__ CUtensorMap tmA,   // seq [M][256] bf16, box 64 x 128
                const __grid_constant__ CUtensorMap tmB,   // W1' [128][256] bf16, box 64 x 128
                const float4* __restrict__ par,            // [128] {s_j, c_j, w2_j, 0}
                float2* __restrict__ stats,                // [T][8][Bc] partial (sum, sumsq) over 32 units each; slot 0 of every
                                                           // row is REPLACED by (mean, rstd) for attn_pool_finish_bf16
                float b2, float* __restrict__ scores, int M, int Bc) {
  extern __shared__ uint8_t sc_smem_raw[];
  const uint32_t raw = smem_u32(sc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = sc_smem_raw + (base - raw);
  const uint32_t sB = base, sA = base + SC_B_BYTES;
  float4* par_s = reinterpret_cast<float4*>(gen + SC_B_BYTES + SC_STAGES * SC_A_BYTES);
  uint8_t* ctl = reinterpret_cast<uint8_t*>(par_s + SC_N);
  const uint32_t bar0 = smem_u32(ctl);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (SC_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * SC_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * SC_STAGES + 2 + a); };
  const uint32_t bfull_bar = bar0 + 8u * (2 * SC_STAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * (2 * SC_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (M + SC_BM - 1) / SC_BM;
  constexpr int k_blocks = SC_K / SC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < SC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  if (warp >= 2) par_s[(warp - 2) * 32 + lane] = __ldg(par + (warp - 2) * 32 + lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull_bar, SC_B_BYTES);
      for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d(sB + kb * SC_B_ATOM, &tmB, kb * SC_BK, 0, bfull_bar);
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), SC_A_BYTES);
          tma_load_2d(sA + stage * SC_A_BYTES, &tmA, kb * SC_BK, tile * SC_BM, full_bar(stage));
          if (++stage == SC_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SC_BM, SC_N);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      mbar_wait(bfull_bar, 0);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * SC_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < SC_BK / 16; ++kk)
            umma_bf16(d_tmem, umma_desc_sw128(sA + stage * SC_A_BYTES + kk * 32), umma_desc_sw128(sB + kb * SC_B_ATOM + kk * 32),
                      idesc, (kb | kk) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == SC_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int gr = tile * SC_BM + quarter * 32 + lane;
      float rs = 0.f, mp = 0.f;
      if (gr < M) {
        // slot-major layout: the 32 lanes (consecutive windows of one time step) read 256 contiguous bytes per slot
        const int tt = gr / Bc, bb = gr - tt * Bc;
        float2* sp = stats + ((long long)tt * 8) * Bc + bb;
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float2 v = sp[(long long)k * Bc]; sum += v.x; sq += v.y; }
        const float mean = sum * (1.0f / SC_K);
        const float var = fmaxf(sq * (1.0f / SC_K) - mean * mean, 0.f);
        rs = 1.0f / sqrtf(var + 1e-5f);
        mp = -rs * mean;
        sp[0] = make_float2(mean, rs);  // only this thread ever reads this row's partials
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * SC_N;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
      float score = 0.f;
#pragma unroll
      for (int ch = 0; ch < SC_N / 32; ++ch) {
        tmem_ld_wait();
        if (ch + 1 < SC_N / 32) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
        const uint32_t* rc = r[ch & 1];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 p = par_s[ch * 32 + j];
          const float pre = fmaf(rs, __uint_as_float(rc[j]), fmaf(mp, p.x, p.y));
          score = fmaf(p.z, tanh_mufu(pre), score);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (gr < M) scores[gr] = score + b2;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- softmax over time, weighted sum, head -------------------------------------------------------
constexpr int PF_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < PF_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// PF_WPC windows per CTA: the classifier weights (128 KB + 32 KB fp32) are read once per PF_WPC windows instead of once per
// window (the first version moved as many bytes of weights through L1 as of sequence data), and the sequence is read with
// 16-byte loads, one warp per time step (512 contiguous bytes), 4 steps in flight per warp.
constexpr int PF_WPC = 4;

// D = LSTM output width (256 for hidden_size 128, 512 for 256); rowstat[t * rs_tstride + b] = (mean, rstd) of row (t, b)
template <int D>
__global__ void __launch_bounds__(PF_THREADS)
attn_pool_finish_bf16(const __nv_bfloat16* __restrict__ seq,  // [T][Bc][D]
                      const float* __restrict__ scores,       // [T][Bc]
                      const float2* __restrict__ stats,       // H=128: [T][8][Bc], slot 0 = (mean, rstd) written by attn_score_bf16
                      long long rs_tstride,
                      int Bc, int T, int classes,
                      const float* __restrict__ lnw, const float* __restrict__ lnb,
                      const float* __restrict__ c0t, const float* __restrict__ cb0,
                      const float* __restrict__ c3t, const float* __restrict__ cb3,
                      const float* __restrict__ c6, const float* __restrict__ cb6,
                      float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ attn) {
  constexpr int H = D / 2, NW = PF_THREADS / 32, CPL = D / 256;  // CPL: 16-byte chunks per lane and row
  extern __shared__ __align__(16) float pf_smem[];
  float* beta = pf_smem;                 // [T]  raw scores first, then beta_t = a_t * rstd_t
  float* bmean = beta + T;               // [T]  row means
  float* part = bmean + T;               // [NW][D] per-warp partial sums
  float* ctx_s = part + NW * D;          // [PF_WPC][D]
  float* h1_s = ctx_s + PF_WPC * D;      // [PF_WPC][H]
  float* h2_s = h1_s + PF_WPC * H;       // [PF_WPC][H/2]
  float* red = h2_s + PF_WPC * (H / 2);  // [8] + logits [PF_WPC][8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b_first = blockIdx.x * PF_WPC;
  const int nwin = (Bc - b_first) < PF_WPC ? (Bc - b_first) : PF_WPC;

  for (int w = 0; w < nwin; ++w) {
    const int b = b_first + w;
    float lmax = -INFINITY;
    for (int t = tid; t < T; t += PF_THREADS) {
      const float s = __ldg(scores + (long long)t * Bc + b);
      beta[t] = s;
      bmean[t] = stats[(long long)t * rs_tstride + b].x;
      lmax = fmaxf(lmax, s);
    }
    const float m = block_reduce(lmax, red, true);
    float lsum = 0.f;
    for (int t = tid; t < T; t += PF_THREADS) lsum += expf(beta[t] - m);
    const float inv_l = 1.0f / block_reduce(lsum, red, false);
    float lgam = 0.f;
    for (int t = tid; t < T; t += PF_THREADS) {
      const float a = expf(beta[t] - m) * inv_l;
      if (attn) attn[(long long)b * T + t] = a;
      const float mean = bmean[t];
      const float bt = a * stats[(long long)t * rs_tstride + b].y;
      beta[t] = bt;
      lgam = fmaf(bt, mean, lgam);
    }
    const float gamma = block_reduce(lgam, red, false);  // also makes beta[] visible to all threads

    // beta-weighted sum of the raw rows: warp = time step (mod NW), lane = 8 consecutive features (one 16-byte load)
    float acc[8 * CPL];
#pragma unroll
    for (int i = 0; i < 8 * CPL; ++i) acc[i] = 0.f;
    const uint4* rowp = reinterpret_cast<const uint4*>(seq) + (long long)b * (D / 8) + lane;
    const long long tstride = (long long)Bc * (D / 8);
    auto fma8 = [&](const uint4& v, float bt, int cp) {
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[cp * 8 + 2 * i] = fmaf(bt, __uint_as_float(wv[i] << 16), acc[cp * 8 + 2 * i]);
        acc[cp * 8 + 2 * i + 1] = fmaf(bt, __uint_as_float(wv[i] & 0xFFFF0000u), acc[cp * 8 + 2 * i + 1]);
      }
    };
    constexpr int UN = 4 / CPL;  // time steps in flight per warp (4 x 16-byte loads per lane either way)
    int t = warp;
    for (; t + (UN - 1) * NW < T; t += UN * NW) {
      uint4 v[UN][CPL];
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int cp = 0; cp < CPL; ++cp) v[u][cp] = ldg_stream_v4(rowp + (long long)(t + u * NW) * tstride + cp * 32);
#pragma unroll
      for (int u = 0; u < UN; ++u)
#pragma unroll
        for (int cp = 0; cp < CPL; ++cp) fma8(v[u][cp], beta[t + u * NW], cp);
    }
    for (; t < T; t += NW)
#pragma unroll
      for (int cp = 0; cp < CPL; ++cp) fma8(ldg_stream_v4(rowp + (long long)t * tstride + cp * 32), beta[t], cp);
#pragma unroll
    for (int cp = 0; cp < CPL; ++cp)
#pragma unroll
      for (int i = 0; i < 8; i += 4)
        *reinterpret_cast<float4*>(part + warp * D + (cp * 32 + lane) * 8 + i) =
            make_float4(acc[cp * 8 + i], acc[cp * 8 + i + 1], acc[cp * 8 + i + 2], acc[cp * 8 + i + 3]);
    __syncthreads();
#pragma unroll
    for (int f = 0; f < CPL; ++f) {
      const int d = f * PF_THREADS + tid;
      float raw = 0.f;
#pragma unroll
      for (int k = 0; k < NW; ++k) raw += part[k * D + d];
      ctx_s[w * D + d] = fmaf(lnw[d], raw - gamma, lnb[d]);
    }
    __syncthreads();
  }
  // classifier for the CTA's windows at once: Linear(2H,H) GELU Linear(H,H/2) GELU Linear(H/2,classes); softmax
  if (tid < H) {
    float a[PF_WPC];
#pragma unroll
    for (int w = 0; w < PF_WPC; ++w) a[w] = cb0[tid];
    for (int d = 0; d < D; ++d) {
      const float wt = __ldg(c0t + (long long)d * H + tid);
#pragma unroll
      for (int w = 0; w < PF_WPC; ++w) a[w] = fmaf(ctx_s[w * D + d], wt, a[w]);
    }
#pragma unroll
    for (int w = 0; w < PF_WPC; ++w) h1_s[w * H + tid] = gelu_erf(a[w]);
  }
  __syncthreads();
  if (tid < H / 2) {
    float a[PF_WPC];
#pragma unroll
    for (int w = 0; w < PF_WPC; ++w) a[w] = cb3[tid];
    for (int k = 0; k < H; ++k) {
      const float wt = __ldg(c3t + k * (H / 2) + tid);
#pragma unroll
      for (int w = 0; w < PF_WPC; ++w) a[w] = fmaf(h1_s[w * H + k], wt, a[w]);
    }
#pragma unroll
    for (int w = 0; w < PF_WPC; ++w) h2_s[w * (H / 2) + tid] = gelu_erf(a[w]);
  }
  __syncthreads();
  for (int i = warp; i < nwin * classes; i += NW) {
    const int w = i / classes, c = i - w * classes;
    float a = 0.f;
    for (int k = lane; k < H / 2; k += 32) a = fmaf(h2_s[w * (H / 2) + k], __ldg(c6 + c * (H / 2) + k), a);
    a = warp_sum(a) + cb6[c];
    if (lane == 0) { logits[(long long)(b_first + w) * classes + c] = a; red[8 + w * 8 + c] = a; }
  }
  if (probs) {
    __syncthreads();
    if (tid < nwin) {
      const float* lg = red + 8 + tid * 8;
      float mx = -INFINITY;
      for (int c = 0; c < classes; ++c) mx = fmaxf(mx, lg[c]);
      float den = 0.f;
      for (int c = 0; c < classes; ++c) den += expf(lg[c] - mx);
      for (int c = 0; c < classes; ++c) probs[(long long)(b_first + tid) * classes + c] = expf(lg[c] - mx) / den;
    }
  }
}

static inline size_t pool_finish_smem(int D, int T) {
  return (size_t)(2 * T + (PF_THREADS / 32) * D + PF_WPC * (D + D / 2 + D / 4) + 8 + PF_WPC * 8) * sizeof(float);
}

// ---- H = 256: scores from a plain tcgen05 GEMM ------------------------------------------------------------------------------
// PRE_raw = X . W1'^T is computed by proj_gemm_bf16 (N = 256, K = 512, bf16 output, zero bias) on the raw rows; this kernel (one
// warp per row) computes the row's LayerNorm statistics from X, folds them in -- pre_j = rstd (PRE_raw_j - mean s_j) + c_j --
// and reduces  sum_j w2_j tanh(pre_j)  to the row's score; (mean, rstd) are kept for the finish kernel.
__global__ void __launch_bounds__(256)
attn_score256_kernel(const __nv_bfloat16* __restrict__ seq,   // [M][512]
                     const __nv_bfloat16* __restrict__ pre,   // [M][256]
                     const float4* __restrict__ par,          // [256] {s_j, c_j, w2_j, 0}
                     float* __restrict__ scores, float2* __restrict__ rowstat, long long M) {
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * 8;
  float4 pj[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pj[i] = __ldg(par + lane * 8 + i);
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < M; r += wstride) {
    const uint4* xr = reinterpret_cast<const uint4*>(seq + r * 512);
    const uint4 x0 = ldg_stream_v4(xr + lane), x1 = ldg_stream_v4(xr + 32 + lane);
    const uint4 pr = ldg_stream_v4(reinterpret_cast<const uint4*>(pre + r * 256) + lane);
    float sm = 0.f, sq = 0.f;
    const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = __uint_as_float(xw[i] << 16), b = __uint_as_float(xw[i] & 0xFFFF0000u);
      sm += a + b;
      sq = fmaf(a, a, fmaf(b, b, sq));
    }
    sm = warp_sum(sm);
    sq = warp_sum(sq);
    const float mean = sm * (1.0f / 512.0f);
    const float rs = 1.0f / sqrtf(fmaxf(sq * (1.0f / 512.0f) - mean * mean, 0.f) + 1e-5f);
    const uint32_t pw[4] = {pr.x, pr.y, pr.z, pr.w};
    float sc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = __uint_as_float((i & 1) ? (pw[i >> 1] & 0xFFFF0000u) : (pw[i >> 1] << 16));
      sc = fmaf(pj[i].z, tanh_mufu(fmaf(rs, a - mean * pj[i].x, pj[i].y)), sc);
    }
    sc = warp_sum(sc);
    if (lane == 0) {
      scores[r] = sc;
      rowstat[r] = make_float2(mean, rs);
    }
  }
}

int pack_pool256_bf16(bci_lstm_s* h, cudaStream_t st) {
  const bci_lstm_weights& w = h->raw;
  pack_attn_w1_kernel<<<ceil_div(256 * 512, 256), 256, 0, st>>>(w.attn_w1, w.ln_w, h->bf16.aw1_bf, 256, 512);
  pack_attn_par_kernel<<<2, 128, 0, st>>>(w.attn_w1, h->bf16.aw1_bf, w.ln_b, w.attn_b1, w.attn_w2, h->bf16.apar, 256, 512);
  BCI_CUDA_OK(cudaMemsetAsync(h->bf16.zero_bias, 0, 256 * sizeof(float), st));
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// seq [T][Bc][512] -> logits / probs / attention; pre [M][256] bf16 and rowstat [M] are workspace
int launch_pool256_bf16(bci_lstm_s* h, const __nv_bfloat16* seq, __nv_bfloat16* pre, float2* rowstat, float* scores, int Bc, int T,
                        float* logits, float* probs, float* attn, cudaStream_t st) {
  const long long M = (long long)Bc * T;
  int rc = launch_proj_gemm_bf16(seq, h->bf16.aw1_bf, h->bf16.zero_bias, pre, (int)M, 256, 512, false, st);
  if (rc) return rc;
  long long blocks = (M + 7) / 8;
  if (blocks > 148ll * 16) blocks = 148ll * 16;
  attn_score256_kernel<<<(unsigned)blocks, 256, 0, st>>>(seq, pre, h->bf16.apar, scores, rowstat, M);
  BCI_LAUNCH_OK();
  const PackedF32& p = h->f32;
  const size_t smem = pool_finish_smem(512, T);
  BCI_REQUIRE(smem <= 48 * 1024, BCI_EINVAL, "bf16 pooling (H=256) supports seq_len <= 2500 (got %d)", T);
  attn_pool_finish_bf16<512><<<ceil_div(Bc, PF_WPC), PF_THREADS, smem, st>>>(seq, scores, rowstat, (long long)Bc, Bc, T, h->cfg.num_classes,
                                                                            p.lnw, p.lnb, p.c0t, p.cb0, p.c3t, p.cb3, p.c6, p.cb6, logits,
                                                                            probs, attn);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int launch_pool_bf16(bci_lstm_s* h, const __nv_bfloat16* seq, float2* stats, float* scores, int Bc, int T, float* logits,
                     float* probs, float* attn, cudaStream_t st) {
  const int M = Bc * T;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16(&tmA, seq, (uint64_t)M, 256, 64, SC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmB, h->bf16.aw1_bf, 128, 256, 64, SC_N);
  if (rc) return rc;
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    BCI_CUDA_OK(cudaFuncSetAttribute(attn_score_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM));
    attr = true;
  }
  // attention.attention.2.bias shifts every score of a window by the same constant and softmax over T is
  // shift invariant, so it cannot change any output (attention weights, context, logits): not applied.
  const float b2 = 0.f;
  const int tiles = ceil_div(M, SC_BM);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  attn_score_bf16<<<grid, SC_THREADS, SC_SMEM, st>>>(tmA, tmB, h->bf16.apar, stats, b2, scores, M, Bc);
  BCI_LAUNCH_OK();
  const PackedF32& p = h->f32;
  const size_t smem = pool_finish_smem(256, T);
  BCI_REQUIRE(smem <= 48 * 1024, BCI_EINVAL, "bf16 pooling supports seq_len <= 4000 (got %d)", T);
  attn_pool_finish_bf16<256><<<ceil_div(Bc, PF_WPC), PF_THREADS, smem, st>>>(seq, scores, stats, 8ll * Bc, Bc, T, h->cfg.num_classes, p.lnw, p.lnb, p.c0t, p.cb0,
                                                       p.c3t, p.cb3, p.c6, p.cb6, logits, probs, attn);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci
