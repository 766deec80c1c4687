// Kernels used by both precision modes: the 61->H input projection (K1) and the fused
// LayerNorm + additive-attention pooling + classifier head (K4/K5).
// Internal activation layout is TIME-MAJOR [T][Bc][feat] so that every per-step access of the
// recurrence is one contiguous slab (the API-level x stays batch-first (B,T,C)).
#pragma once
#include "lstm_handle.cuh"

namespace bci {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------
// K1: z = GELU_erf(LayerNorm_eps1e-5(x W0^T + b0))            04_lstm_model.py:173-178,208
// One warp per (b,t) row; W0^T staged once per CTA in shared memory; x row broadcast by shuffle.
// x (B,T,C) batch-first fp32 (rows b0..b0+Bc of it) -> z [T][Bc][H] time-major.
// ---------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 256;

template <int H, typename OutT>
__global__ void __launch_bounds__(K1_THREADS)
input_proj_kernel(const InputView x, int Bc, int T, int C, const float* __restrict__ w0t,
                  const float* __restrict__ b0, const float* __restrict__ lnw, const float* __restrict__ lnb,
                  OutT* __restrict__ z, int flags) {   // flags: bit 0 = LayerNorm present, bit 1 = round x to bf16 on load
  const int use_ln = flags & 1;
  const bool round_x = (flags & 2) != 0;   // the bf16 engine's contract: bf16 storage of the input changes no result bit
  extern __shared__ __align__(16) float k1_smem[];  // [C][H]
  constexpr int NV = H / 32;          // outputs per lane (4 or 8)
  constexpr int NQ = NV / 4;          // float4 groups per lane
  for (int i = threadIdx.x; i < C * H; i += K1_THREADS) k1_smem[i] = w0t[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long rows = (long long)Bc * T;
  const long long wstride = (long long)gridDim.x * (K1_THREADS / 32);
  float bias[NV], gw[NV], gb[NV];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = q * 128 + lane * 4 + v;
      bias[q * 4 + v] = b0[j]; gw[q * 4 + v] = use_ln ? lnw[j] : 1.f; gb[q * 4 + v] = use_ln ? lnb[j] : 0.f;
    }
  for (long long r = (long long)blockIdx.x * (K1_THREADS / 32) + warp; r < rows; r += wstride) {
    const int b = (int)(r / T), t = (int)(r - (long long)b * T);
    const long long xr = x.elem_off(b) + (long long)t * C;   // first element of the row (windows may overlap / be bf16: InputView)
    float xa = lane < C ? view_load(x, xr + lane) : 0.f;
    float xb = (32 + lane) < C ? view_load(x, xr + 32 + lane) : 0.f;
    if (round_x) { xa = __bfloat162float(__float2bfloat16_rn(xa)); xb = __bfloat162float(__float2bfloat16_rn(xb)); }
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = bias[v];
    for (int c = 0; c < C; ++c) {
      const float xv = __shfl_sync(0xffffffffu, c < 32 ? xa : xb, c & 31);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(k1_smem + c * H + q * 128 + lane * 4);
        acc[q * 4 + 0] = fmaf(xv, w.x, acc[q * 4 + 0]);
        acc[q * 4 + 1] = fmaf(xv, w.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(xv, w.z, acc[q * 4 + 2]);
        acc[q * 4 + 3] = fmaf(xv, w.w, acc[q * 4 + 3]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) s += acc[v];
    float mean = warp_sum(s) * (1.0f / H);
    float q2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) { const float d = acc[v] - mean; q2 = fmaf(d, d, q2); }
    float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / H) + 1e-5f);
    if (!use_ln) { mean = 0.f; rstd = 1.f; }  // nn.Identity instead of LayerNorm (09:191)
    OutT* zr = z + ((long long)t * Bc + b) * H;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float y = use_ln ? (acc[q * 4 + v] - mean) * rstd * gw[q * 4 + v] + gb[q * 4 + v] : acc[q * 4 + v];
        zr[q * 128 + lane * 4 + v] = from_f32<OutT>(gelu_erf(y));
      }
  }
}

// ---------------------------------------------------------------------------------------------
// K4/K5: y = LN(out); s_t = w2.tanh(W1 y_t + b1) + b2; a = softmax_T(s); ctx = sum_t a_t y_t;
// logits = classifier(ctx); probs = softmax(logits).     04_lstm_model.py:112-128,192-204,212-218
// One CTA per window.  Time is processed in chunks of TC rows held (LayerNorm-ed, transposed) in
// shared memory; the softmax-weighted sum is accumulated online (running max / denominator) so
// the sequence is read from HBM exactly once.  Reductions over hidden units are warp shuffles.
// ---------------------------------------------------------------------------------------------
constexpr int K4_THREADS = 256;

template <int H, int ND>
struct PoolCfg {
  static constexpr int D = ND * H;               // LSTM output width
  static constexpr int AH = D / 2;               // attention hidden width (04:115: hidden_size // 2 of the 2H-wide input)
  static constexpr int GROUPS = K4_THREADS / AH; // row groups of the score GEMM: 1, 2 or 4
  static constexpr int RPT = 16;                 // rows (timesteps) per thread in the score GEMM
  static constexpr int TC = GROUPS * RPT;        // timesteps per chunk: 16, 32 or 64
  static constexpr int YS_STRIDE = TC + 4;       // padded, keeps float4 alignment
  static constexpr int WARPS_PER_GROUP = AH / 32;
  // smem floats: ys_t[D][YS_STRIDE] + red[TC][WARPS_PER_GROUP] + chunk_s[TC] + ctx[D] + h1[H] + h2[H/2] + misc
  static constexpr int SMEM_FLOATS = D * YS_STRIDE + TC * WARPS_PER_GROUP + TC + D + H + H / 2 + 8;
};

// use_ln = 0: nn.Identity instead of the final LayerNorm (09:210); use_attn = 0: mean over time instead of attention
// pooling (09:232-234) -- constant scores make the online softmax below exactly that mean.
template <int H, int ND, typename InT>
__global__ void __launch_bounds__(K4_THREADS)
attn_pool_head_kernel(const InT* __restrict__ seq,  // [T][Bc][2H]
                      int Bc, int T, int classes,
                      const float* __restrict__ lnw, const float* __restrict__ lnb,
                      const float* __restrict__ aw1t, const float* __restrict__ ab1,
                      const float* __restrict__ aw2, const float* __restrict__ ab2,
                      const float* __restrict__ c0t, const float* __restrict__ cb0,
                      const float* __restrict__ c3t, const float* __restrict__ cb3,
                      const float* __restrict__ c6, const float* __restrict__ cb6,
                      float* __restrict__ logits,   // [Bc][classes]
                      float* __restrict__ probs,    // [Bc][classes] or null
                      float* __restrict__ attn,     // [Bc][T] or null
                      float* __restrict__ scores_ws, // [Bc][T] scratch for raw scores (needed iff attn)
                      int use_ln, int use_attn,
                      const float* __restrict__ pre = nullptr) {  // optional [T][Bc][AH]: W1 LN(seq) + b1 already computed (tensor cores)
  using Cfg = PoolCfg<H, ND>;
  constexpr int D = Cfg::D, AH = Cfg::AH, TC = Cfg::TC, RPT = Cfg::RPT, YS = Cfg::YS_STRIDE, WPG = Cfg::WARPS_PER_GROUP;
  extern __shared__ __align__(16) float k4_smem[];
  float* ys_t = k4_smem;                 // [D][YS]   LayerNorm-ed chunk, transposed
  float* red = ys_t + D * YS;            // [TC][WPG] partial scores per warp
  float* chunk_s = red + TC * WPG;       // [TC]
  float* ctx_s = chunk_s + TC;           // [D]
  float* h1_s = ctx_s + D;               // [H]
  float* h2_s = h1_s + H;                // [H/2]
  float* misc = h2_s + H / 2;            // [8]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid % AH;                // hidden unit of the score GEMM
  const int grp = tid / AH;              // row group
  const int wig = (tid % AH) >> 5;       // warp index inside the group
  constexpr int DPT = (D + K4_THREADS - 1) / K4_THREADS;  // ctx features per thread (1 or 2; threads >= D idle for D = 128)
  constexpr int EPL = D / 32;            // elements per lane in the LN pass (8 or 16)

  float m_run = -INFINITY, l_run = 0.f;
  float ctx[DPT];
#pragma unroll
  for (int q = 0; q < DPT; ++q) ctx[q] = 0.f;
  const float b1 = use_attn ? ab1[j] : 0.f, w2 = use_attn ? aw2[j] : 0.f, b2 = use_attn ? ab2[0] : 0.f;

  for (int t0 = 0; t0 < T; t0 += TC) {
    const int rows = min(TC, T - t0);
    // ---- LayerNorm each row of the chunk (one warp per row), store transposed -------------
    for (int r = warp; r < TC; r += K4_THREADS / 32) {
      if (r < rows) {
        const InT* src = seq + ((long long)(t0 + r) * Bc + b) * D;
        float v[EPL];
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) { v[e] = to_f32<InT>(src[e * 32 + lane]); s += v[e]; }
        const float mean = warp_sum(s) * (1.0f / D);
        float q2 = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) { const float d = v[e] - mean; q2 = fmaf(d, d, q2); }
        const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const int d = e * 32 + lane;
          ys_t[d * YS + r] = use_ln ? (v[e] - mean) * rstd * lnw[d] + lnb[d] : v[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) ys_t[(e * 32 + lane) * YS + r] = 0.f;
      }
    }
    __syncthreads();
    // ---- scores: u[r][j] = tanh(b1[j] + sum_d y[r][d] W1[j][d]) -----------------------------
    float acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] = b1;
    const float* yrow = ys_t + grp * RPT;
    if (pre) {
      // the score pre-activations come from the split-fp16 tensor-core GEMM (fp32 large-batch path): one coalesced row read each
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int t = t0 + grp * RPT + r;
        acc[r] = t < T ? __ldg(pre + ((long long)t * Bc + b) * AH + j) : 0.f;
      }
    } else if (use_attn) {
#pragma unroll 4
    for (int d = 0; d < D; ++d) {
      const float w = __ldg(aw1t + (long long)d * AH + j);
      const float4* yp = reinterpret_cast<const float4*>(yrow + d * YS);
#pragma unroll
      for (int q = 0; q < RPT / 4; ++q) {
        const float4 y4 = yp[q];
        acc[q * 4 + 0] = fmaf(y4.x, w, acc[q * 4 + 0]);
        acc[q * 4 + 1] = fmaf(y4.y, w, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(y4.z, w, acc[q * 4 + 2]);
        acc[q * 4 + 3] = fmaf(y4.w, w, acc[q * 4 + 3]);
      }
    }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float part = warp_sum(w2 * tanhf(acc[r]));
      if (lane == 0) red[(grp * RPT + r) * WPG + wig] = part;
    }
    __syncthreads();
    if (tid < TC) {
      float s = b2;
#pragma unroll
      for (int w = 0; w < WPG; ++w) s += red[tid * WPG + w];
      chunk_s[tid] = tid < rows ? s : -INFINITY;
      if (attn && tid < rows) scores_ws[(long long)b * T + t0 + tid] = s;
    }
    __syncthreads();
    // ---- online softmax-weighted accumulation ------------------------------------------------
    // (rows beyond the sequence end carry score -inf and y = 0, so they contribute exactly 0)
    float cmax = -INFINITY;
#pragma unroll
    for (int r = 0; r < TC; ++r) cmax = fmaxf(cmax, chunk_s[r]);
    const float m_new = fmaxf(m_run, cmax);
    const float scale = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    l_run *= scale;
#pragma unroll
    for (int q = 0; q < DPT; ++q) ctx[q] *= scale;
#pragma unroll
    for (int r = 0; r < TC; r += 4) {
      const float p0 = expf(chunk_s[r] - m_new), p1 = expf(chunk_s[r + 1] - m_new);
      const float p2 = expf(chunk_s[r + 2] - m_new), p3 = expf(chunk_s[r + 3] - m_new);
      l_run += (p0 + p1) + (p2 + p3);
#pragma unroll
      for (int q = 0; q < DPT; ++q) {
        if (q * K4_THREADS + tid < D) {
          const float4 y4 = *reinterpret_cast<const float4*>(ys_t + (q * K4_THREADS + tid) * YS + r);
          ctx[q] = fmaf(p3, y4.w, fmaf(p2, y4.z, fmaf(p1, y4.y, fmaf(p0, y4.x, ctx[q]))));
        }
      }
    }
    m_run = m_new;
    __syncthreads();  // ys_t / chunk_s reused by the next chunk
  }
  const float inv_l = 1.0f / l_run;
#pragma unroll
  for (int q = 0; q < DPT; ++q)
    if (q * K4_THREADS + tid < D) ctx_s[q * K4_THREADS + tid] = ctx[q] * inv_l;
  if (attn) {
    for (int t = tid; t < T; t += K4_THREADS)
      attn[(long long)b * T + t] = expf(scores_ws[(long long)b * T + t] - m_run) * inv_l;
  }
  __syncthreads();
  // ---- classifier: Linear(2H,H) GELU Linear(H,H/2) GELU Linear(H/2,classes) ----------------
  if (tid < H) {
    float a = cb0[tid];
    for (int d = 0; d < D; ++d) a = fmaf(ctx_s[d], __ldg(c0t + (long long)d * H + tid), a);
    h1_s[tid] = gelu_erf(a);
  }
  __syncthreads();
  if (tid < H / 2) {
    float a = cb3[tid];
    for (int k = 0; k < H; ++k) a = fmaf(h1_s[k], __ldg(c3t + k * (H / 2) + tid), a);
    h2_s[tid] = gelu_erf(a);
  }
  __syncthreads();
  for (int c = warp; c < classes; c += K4_THREADS / 32) {
    float a = 0.f;
    for (int k = lane; k < H / 2; k += 32) a = fmaf(h2_s[k], __ldg(c6 + c * (H / 2) + k), a);
    a = warp_sum(a) + cb6[c];
    if (lane == 0) { logits[(long long)b * classes + c] = a; if (c < 8) misc[c] = a; }
  }
  if (probs) {
    __syncthreads();
    if (tid == 0) {
      // softmax(dim=1); classes <= 8 kept in smem, otherwise re-read from global
      float mx = -INFINITY;
      for (int c = 0; c < classes; ++c) mx = fmaxf(mx, c < 8 ? misc[c] : logits[(long long)b * classes + c]);
      float den = 0.f;
      for (int c = 0; c < classes; ++c) den += expf((c < 8 ? misc[c] : logits[(long long)b * classes + c]) - mx);
      for (int c = 0; c < classes; ++c)
        probs[(long long)b * classes + c] = expf((c < 8 ? misc[c] : logits[(long long)b * classes + c]) - mx) / den;
    }
  }
}

template <int H, int ND, typename InT>
inline int launch_pool_head(const bci_lstm_s* h, const InT* seq, int Bc, int T, float* logits, float* probs, float* attn,
                            float* scores_ws, cudaStream_t st, const float* pre = nullptr) {
  using Cfg = PoolCfg<H, ND>;
  const size_t smem = Cfg::SMEM_FLOATS * sizeof(float);
  static PerDeviceFlag attr_pd;
  bool& attr_set = attr_pd.cur();
  if (!attr_set) {
    BCI_CUDA_OK(cudaFuncSetAttribute(attn_pool_head_kernel<H, ND, InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const PackedF32& p = h->f32;
  attn_pool_head_kernel<H, ND, InT><<<Bc, K4_THREADS, smem, st>>>(seq, Bc, T, h->cfg.num_classes, p.lnw, p.lnb, p.aw1t, p.ab1,
                                                                   p.aw2, p.ab2, p.c0t, p.cb0, p.c3t, p.cb3, p.c6, p.cb6,
                                                                   logits, probs, attn, scores_ws, h->cfg.use_layer_norm,
                                                                   h->cfg.use_attention, pre);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// x (any InputView) -> fp16 (hi, lo) pair rows [T][Bc][64] (time-major, K padded from C to 64 with zeros): A operand of the input
// projection in its split-fp16 GEMM form (fp32 large-batch path).  One thread per (row, pair of channels).
static __global__ void x_pair_rows_kernel(const InputView x, int Bc, int T, int C, __half* __restrict__ x_hi, __half* __restrict__ x_lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)Bc * T;
  if (i >= rows * 32) return;
  const long long r = i >> 5;
  const int k = (int)(i & 31) * 2;
  const int t = (int)(r / Bc), b = (int)(r - (long long)t * Bc);
  const long long src = x.elem_off(b) + (long long)t * C;
  const float v0 = k < C ? view_load(x, src + k) : 0.f, v1 = (k + 1) < C ? view_load(x, src + k + 1) : 0.f;
  const __half2 h2 = __floats2half2_rn(v0, v1);
  const float2 back = __half22float2(h2);
  reinterpret_cast<__half2*>(x_hi)[i] = h2;
  reinterpret_cast<__half2*>(x_lo)[i] = __floats2half2_rn(v0 - back.x, v1 - back.y);
}

// z = GELU_erf(LayerNorm(pre row)) -> fp16 (hi, lo) pair: rows of 128, one warp per row, 4 consecutive features per lane
static __global__ void __launch_bounds__(256)
ln_gelu_pair_rows128_kernel(const float* __restrict__ pre, long long rows, const float* __restrict__ lnw, const float* __restrict__ lnb,
                            int use_ln, __half* __restrict__ z_hi, __half* __restrict__ z_lo) {
  const int lane = threadIdx.x & 31;
  float gw[4], gb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { gw[i] = use_ln ? lnw[lane * 4 + i] : 1.f; gb[i] = use_ln ? lnb[lane * 4 + i] : 0.f; }
  const long long wstride = (long long)gridDim.x * 8;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(pre + r * 128) + lane);
    const float v[4] = {a.x, a.y, a.z, a.w};
    float mean = warp_sum((v[0] + v[1]) + (v[2] + v[3])) * (1.0f / 128);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float dd = v[i] - mean; q2 = fmaf(dd, dd, q2); }
    float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / 128) + 1e-5f);
    if (!use_ln) { mean = 0.f; rstd = 1.f; }
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float y0 = gelu_erf(use_ln ? (v[2 * i] - mean) * rstd * gw[2 * i] + gb[2 * i] : v[2 * i]);
      const float y1 = gelu_erf(use_ln ? (v[2 * i + 1] - mean) * rstd * gw[2 * i + 1] + gb[2 * i + 1] : v[2 * i + 1]);
      const __half2 h2 = __floats2half2_rn(y0, y1);
      const float2 back = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(y0 - back.x, y1 - back.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    *reinterpret_cast<uint2*>(z_hi + r * 128 + lane * 4) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(z_lo + r * 128 + lane * 4) = make_uint2(lo[0], lo[1]);
  }
}

// y = LayerNorm(seq row) -> fp16 (hi, lo) pair, the A operand of the score GEMM in its split-fp16 form.  One warp per row,
// 8 consecutive features per lane (D = 256): 1 KB coalesced in, 2 x 512 B coalesced out.
static __global__ void __launch_bounds__(256)
ln_pair_rows256_kernel(const float* __restrict__ seq, long long rows, const float* __restrict__ lnw, const float* __restrict__ lnb,
                       __half* __restrict__ y_hi, __half* __restrict__ y_lo) {
  const int lane = threadIdx.x & 31;
  float gw[8], gb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { gw[i] = lnw[lane * 8 + i]; gb[i] = lnb[lane * 8 + i]; }
  const long long wstride = (long long)gridDim.x * 8;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    const float4* p = reinterpret_cast<const float4*>(seq + r * 256) + lane * 2;
    const float4 a = __ldg(p), b = __ldg(p + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / 256);
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float dd = v[i] - mean; q2 = fmaf(dd, dd, q2); }
    const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / 256) + 1e-5f);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y0 = (v[2 * i] - mean) * rstd * gw[2 * i] + gb[2 * i];
      const float y1 = (v[2 * i + 1] - mean) * rstd * gw[2 * i + 1] + gb[2 * i + 1];
      const __half2 h2 = __floats2half2_rn(y0, y1);
      const float2 back = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(y0 - back.x, y1 - back.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    *reinterpret_cast<uint4*>(y_hi + r * 256 + lane * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(y_lo + r * 256 + lane * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int H, typename OutT>
inline int launch_input_proj(const bci_lstm_s* h, const InputView& x, int Bc, int T, OutT* z, cudaStream_t st, bool round_x_bf16 = false) {
  const int C = h->cfg.input_size;
  const size_t smem = (size_t)C * H * sizeof(float);
  static PerDeviceFlag attr_pd;
  bool& attr_set = attr_pd.cur();
  if (!attr_set) {
    BCI_CUDA_OK(cudaFuncSetAttribute(input_proj_kernel<H, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long rows = (long long)Bc * T;
  long long blocks = ceil_div64(rows, K1_THREADS / 32);
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  const PackedF32& p = h->f32;
  input_proj_kernel<H, OutT><<<(unsigned)blocks, K1_THREADS, smem, st>>>(x, Bc, T, C, p.w0t, p.b0, p.ln0w, p.ln0b, z,
                                                                         (h->cfg.use_layer_norm ? 1 : 0) | (round_x_bf16 ? 2 : 0));
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci
