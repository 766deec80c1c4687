// Training step of the BiLSTM-attention classifier in fp32: forward with saved activations, full
// backward (BPTT through 3 x 2 LSTM directions, attention pooling, LayerNorms, heads) and a fused
// clip + AdamW update.  Replaces loss.backward() / clip_grad_norm_ / AdamW.step of the reference's
// train loop (04_lstm_model.py:482-507; plain variant 09_sensitivity_analysis.py:297-303) and the
// backward-to-input used by the attribution script (07_explainability.py:242-258).
//
// Structure: every large dense contraction over the flattened row space M = T*Bc runs on the tensor cores in split
// precision (gemm_tf32x3.cu: NT for projections / data gradients, TN with split-K + TMA reduce-add for weight gradients;
// shapes it does not cover -- the 61-channel input projection, the classifier heads, tiny test batches -- fall back to
// the two generic CUDA-core GEMMs below); the serial parts are lstm_rec_f32 (forward, saving gate activations and cell
// states) and lstm_bptt_f32 (its mirror: dh_{t-1} = dG_t . W_hh; training-size batches keep 96 of the 128 W_hh rows in
// shared memory and split the unit rows of the product between the CTA's two thread groups); everything
// row-local (LayerNorm, GELU, softmax over T, head MLP) is small fused kernels.  Dropout (four sites,
// 04:177,186,199,202) uses a stateless hash of (seed, site, element index) so the backward pass
// regenerates the masks instead of storing them.
#include "lstm_shared_kernels.cuh"

namespace bci {

__device__ __forceinline__ float gelu_grad(float x) {
  // d/dx [0.5 x (1 + erf(x/sqrt2))] = Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  return cdf + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// ---- generic GEMMs (bounds-checked, 64x64x16 tiles, 4x4 per thread) -----------------------------
constexpr int TG = 64, TGK = 16, TG_THREADS = 256;

// C[M][N] (ldc) = (accumulate ? C : 0) + A[M][K] (lda) . Bt[K][N] (ldb) + bias[N]
__global__ void __launch_bounds__(TG_THREADS)
gemm_nn_f32(const float* __restrict__ A, int lda, const float* __restrict__ Bt, int ldb, float* __restrict__ C, int ldc, int M, int N,
            int K, const float* __restrict__ bias, int accumulate) {
  __shared__ float As[TGK][TG + 1];
  __shared__ float Bs[TGK][TG + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TG, n0 = blockIdx.x * TG;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TGK) {
    for (int e = tid; e < TG * TGK; e += TG_THREADS) {
      const int m = e / TGK, k = e % TGK;  // A: k fastest (row-major rows)
      As[k][m] = (m0 + m < M && k0 + k < K) ? A[(long long)(m0 + m) * lda + k0 + k] : 0.f;
      const int kk = e / TG, n = e % TG;   // Bt: n fastest
      Bs[kk][n] = (k0 + kk < K && n0 + n < N) ? Bt[(long long)(k0 + kk) * ldb + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TGK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      float* c = C + (long long)m * ldc + n;
      *c = accumulate ? *c + v : v;
    }
  }
}

// C[P][Q] (ldc) += sum_{r < R} A[r][p] (lda) * B[r][q] (ldb); grid.z splits R; C must be zeroed first
__global__ void __launch_bounds__(TG_THREADS)
gemm_tn_f32(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, long long R,
            int P, int Q) {
  __shared__ float As[TGK][TG + 1];
  __shared__ float Bs[TGK][TG + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.y * TG, q0 = blockIdx.x * TG;
  const long long per = (R + gridDim.z - 1) / gridDim.z;
  const long long r_begin = per * blockIdx.z, r_end = (r_begin + per < R) ? r_begin + per : R;
  float acc[4][4] = {};
  for (long long r0 = r_begin; r0 < r_end; r0 += TGK) {
    for (int e = tid; e < TG * TGK; e += TG_THREADS) {
      const int rr = e / TG, c = e % TG;
      const long long r = r0 + rr;
      As[rr][c] = (r < r_end && p0 + c < P) ? A[r * lda + p0 + c] : 0.f;
      Bs[rr][c] = (r < r_end && q0 + c < Q) ? B[r * ldb + q0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TGK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + ty * 4 + i, q = q0 + tx * 4 + j;
      if (p < P && q < Q) atomicAdd(C + (long long)p * ldc + q, acc[i][j]);
    }
}

// Fast path of the same contraction for P % 128 == 0, Q % 128 == 0, lda/ldb % 4 == 0: 128x128x16 tiles, 8x8 per thread,
// 16-byte loads (both operands are row-major in the reduction index, so tiles go to smem without a transpose).
__global__ void __launch_bounds__(256)
gemm_tn_f32_fast(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, long long R) {
  __shared__ __align__(16) float As[2][16][128];
  __shared__ __align__(16) float Bs[2][16][128];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.y * 128, q0 = blockIdx.x * 128;
  const long long per = ((R + gridDim.z - 1) / gridDim.z + 15) / 16 * 16;
  const long long r_begin = per * blockIdx.z, r_end = (r_begin + per < R) ? r_begin + per : R;
  if (r_begin >= r_end) return;
  const int lr = tid >> 5, lc = (tid & 31) * 4;  // rows lr, lr+8 of the 16-row slab; 4 consecutive columns
  float4 ra[2], rb[2];
  auto load = [&](long long r0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long r = r0 + lr + i * 8;
      const bool ok = r < r_end;
      ra[i] = ok ? *reinterpret_cast<const float4*>(A + r * lda + p0 + lc) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = ok ? *reinterpret_cast<const float4*>(B + r * ldb + q0 + lc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float4*>(&As[buf][lr + i * 8][lc]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][lr + i * 8][lc]) = rb[i];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load(r_begin);
  store(0);
  __syncthreads();
  int buf = 0;
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
    const bool more = r0 + 16 < r_end;
    if (more) load(r0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = q0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      atomicAdd(C + (long long)p * ldc + q, acc[i][j]);
    }
  }
}

// out[n] += sum_r A[r][n]   (column sums, atomics per block)
__global__ void colsum_kernel(const float* __restrict__ A, int lda, long long R, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long per = (R + gridDim.y - 1) / gridDim.y;
  const long long r0 = per * blockIdx.y, r1 = (r0 + per < R) ? r0 + per : R;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long long r = r0;
  for (; r + 3 < r1; r += 4) {   // four independent loads in flight per thread
    s0 += A[r * lda + n]; s1 += A[(r + 1) * lda + n]; s2 += A[(r + 2) * lda + n]; s3 += A[(r + 3) * lda + n];
  }
  for (; r < r1; ++r) s0 += A[r * lda + n];
  atomicAdd(out + n, (s0 + s1) + (s2 + s3));
}

static int gemm_nn(const float* A, int lda, const float* Bt, int ldb, float* C, int ldc, int M, int N, int K, const float* bias,
                   int accumulate, cudaStream_t st) {
  if (N % 128 == 0 && K % 16 == 0 && lda == K && ldb == N && ldc == N && ((uintptr_t)A & 15) == 0 && ((uintptr_t)Bt & 15) == 0 &&
      ((uintptr_t)C & 15) == 0)
    return launch_proj_gemm_f32(A, Bt, bias, C, M, N, K, st, accumulate);
  dim3 g(ceil_div(N, TG), ceil_div(M, TG));
  gemm_nn_f32<<<g, TG_THREADS, 0, st>>>(A, lda, Bt, ldb, C, ldc, M, N, K, bias, accumulate);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
static int gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, long long R, int P, int Q, cudaStream_t st) {
  if (R <= 0) return BCI_OK;
  if (P % 128 == 0 && Q % 128 == 0 && lda % 4 == 0 && ldb % 4 == 0 && R >= 1024 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0) {
    const int tiles = (P / 128) * (Q / 128);
    long long splits = (2 * sm_count() + tiles - 1) / tiles;
    if (splits > R / 256) splits = R / 256;
    if (splits < 1) splits = 1;
    gemm_tn_f32_fast<<<dim3(Q / 128, P / 128, (unsigned)splits), 256, 0, st>>>(A, lda, B, ldb, C, ldc, R);
    BCI_LAUNCH_OK();
    return BCI_OK;
  }
  const int tiles = ceil_div(P, TG) * ceil_div(Q, TG);
  int splits = (4 * sm_count() + tiles - 1) / tiles;
  // row ranges of at least 32 rows: the classifier gradients reduce over the batch only (R = 512), where 256-row ranges left a
  // handful of CTAs each walking 16 dependent k-steps (48 us per launch for a few MFLOP)
  const long long max_splits = (R + 31) / 32;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  dim3 g(ceil_div(Q, TG), ceil_div(P, TG), splits);
  gemm_tn_f32<<<g, TG_THREADS, 0, st>>>(A, lda, B, ldb, C, ldc, R, P, Q);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
static int colsum(const float* A, int lda, long long R, int N, float* out, cudaStream_t st) {
  // enough row ranges to fill the machine whatever N is (N = 128: one column block; 1024-row ranges left it at 128 CTAs)
  const int xb = ceil_div(N, 128);
  long long want = (8LL * sm_count() + xb - 1) / xb;
  int ys = (int)((R + 63) / 64 < want ? (R + 63) / 64 : want);
  if (ys > 4096) ys = 4096;
  if (ys < 1) ys = 1;
  colsum_kernel<<<dim3(ceil_div(N, 128), ys), 128, 0, st>>>(A, lda, R, N, out);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---- row-local forward kernels -----------------------------------------------------------------
// K1 (train): one warp per FOUR (t,b) rows.  Saves xT (time-major copy of x, rows padded to 64 floats so that the weight-gradient
// GEMM can read it through a tensor map), xhat0 (normalised pre-activation), rstd0 and z = dropout(GELU(LN(.))).  x (Bc,T,C)
// batch-first.  W0^T sits in shared memory; a lane owns output columns q*128 + 4 lane .. + 3, so one 16-byte shared-memory read per
// input channel feeds 16 FMAs (four rows) -- the first version (one row per warp, weights from global memory, the row copy written
// by lane 0 element by element) took 261 us per 131 072 rows.
constexpr int K1T_ROWS = 4, K1T_XS = 64;
template <int H>
__global__ void __launch_bounds__(256)
inproj_train_fwd(const float* __restrict__ x, int Bc, int T, int C, const float* __restrict__ w0t, const float* __restrict__ b0,
                 const float* __restrict__ lnw, const float* __restrict__ lnb, float* __restrict__ xT, float* __restrict__ xhat,
                 float* __restrict__ rstd_out, float* __restrict__ z, float p_drop, uint64_t seed, int use_ln) {
  extern __shared__ __align__(16) float k1t_w[];   // [C][H]
  constexpr int NQ = H / 128;
  for (int i = threadIdx.x; i < C * H; i += 256) k1t_w[i] = w0t[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)Bc * T;
  const long long groups = (rows + K1T_ROWS - 1) / K1T_ROWS;
  const long long wstride = (long long)gridDim.x * 8;
  float4 bias[NQ], gw[NQ], gb[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int j = q * 128 + lane * 4;
    bias[q] = *reinterpret_cast<const float4*>(b0 + j);
    gw[q] = use_ln ? *reinterpret_cast<const float4*>(lnw + j) : make_float4(1.f, 1.f, 1.f, 1.f);
    gb[q] = use_ln ? *reinterpret_cast<const float4*>(lnb + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long grp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); grp < groups; grp += wstride) {
    float xa[K1T_ROWS], xb[K1T_ROWS];
    float4 acc[K1T_ROWS][NQ];
#pragma unroll
    for (int i = 0; i < K1T_ROWS; ++i) {
      const long long r = grp * K1T_ROWS + i;   // TIME-MAJOR row index
      xa[i] = xb[i] = 0.f;
      if (r < rows) {
        const int t = (int)(r / Bc), b = (int)(r - (long long)t * Bc);
        const float* xr = x + ((long long)b * T + t) * C;
        if (lane < C) xa[i] = xr[lane];
        if (32 + lane < C) xb[i] = xr[32 + lane];
        xT[r * K1T_XS + lane] = xa[i];
        xT[r * K1T_XS + 32 + lane] = xb[i];
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[i][q] = bias[q];
    }
    for (int c = 0; c < C; ++c) {
      float xv[K1T_ROWS];
#pragma unroll
      for (int i = 0; i < K1T_ROWS; ++i) xv[i] = __shfl_sync(0xffffffffu, c < 32 ? xa[i] : xb[i], c & 31);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(k1t_w + c * H + q * 128 + lane * 4);
#pragma unroll
        for (int i = 0; i < K1T_ROWS; ++i) {
          acc[i][q].x = fmaf(xv[i], w.x, acc[i][q].x); acc[i][q].y = fmaf(xv[i], w.y, acc[i][q].y);
          acc[i][q].z = fmaf(xv[i], w.z, acc[i][q].z); acc[i][q].w = fmaf(xv[i], w.w, acc[i][q].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < K1T_ROWS; ++i) {
      const long long r = grp * K1T_ROWS + i;
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) s += (acc[i][q].x + acc[i][q].y) + (acc[i][q].z + acc[i][q].w);
      const float mean = warp_sum(s) * (1.0f / H);
      float q2 = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float d0 = acc[i][q].x - mean, d1 = acc[i][q].y - mean, d2 = acc[i][q].z - mean, d3 = acc[i][q].w - mean;
        q2 = fmaf(d0, d0, q2); q2 = fmaf(d1, d1, q2); q2 = fmaf(d2, d2, q2); q2 = fmaf(d3, d3, q2);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / H) + 1e-5f);
      if (r >= rows) continue;
      if (lane == 0) rstd_out[r] = rstd;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int j = q * 128 + lane * 4;
        const float a4[4] = {acc[i][q].x, acc[i][q].y, acc[i][q].z, acc[i][q].w};
        const float w4[4] = {gw[q].x, gw[q].y, gw[q].z, gw[q].w}, b4[4] = {gb[q].x, gb[q].y, gb[q].z, gb[q].w};
        float xh4[4], z4[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          // without LayerNorm (09:191) xhat holds the raw pre-activation and y = xhat
          xh4[v] = use_ln ? (a4[v] - mean) * rstd : a4[v];
          const float y = use_ln ? fmaf(xh4[v], w4[v], b4[v]) : xh4[v];
          z4[v] = gelu_erf(y) * drop_scale(seed, 0, (uint64_t)r * H + j + v, p_drop);
        }
        *reinterpret_cast<float4*>(xhat + r * H + j) = make_float4(xh4[0], xh4[1], xh4[2], xh4[3]);
        *reinterpret_cast<float4*>(z + r * H + j) = make_float4(z4[0], z4[1], z4[2], z4[3]);
      }
    }
  }
}

// per-warp partial sums of NV columns per lane (column e*32 + lane) -> one atomicAdd per column and BLOCK: every warp adding its own
// partials put 32 768 warps x 512 atomics on 512 addresses (ln_rows_bwd spent 300 us of the step there)
template <int NV>
__device__ __forceinline__ void block_col_atomic_add(float* __restrict__ dst_w, float* __restrict__ dst_b, const float (&gw)[NV],
                                                     const float (&gb)[NV]) {
  __shared__ float red[2][8][NV * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < NV; ++e) { red[0][warp][e * 32 + lane] = gw[e]; red[1][warp][e * 32 + lane] = gb[e]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * NV * 32; i += 256) {
    const int which = i / (NV * 32), col = i - which * NV * 32;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[which][w][col];
    atomicAdd((which ? dst_b : dst_w) + col, s);
  }
}

// dv = LayerNorm-backward(dz * mask * gelu'(y)); accumulates dlnw/dlnb.  One warp per row.
template <int H>
__global__ void __launch_bounds__(256)
inproj_bwd_rows(const float* __restrict__ dz, const float* __restrict__ xhat, const float* __restrict__ rstd_in,
                const float* __restrict__ lnw, const float* __restrict__ lnb, long long rows, float* __restrict__ dv,
                float* __restrict__ dlnw, float* __restrict__ dlnb, float p_drop, uint64_t seed, int use_ln) {
  constexpr int NV = H / 32;
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * 8;
  float gw[NV] = {}, gb[NV] = {};
  if (!use_ln) {  // dv = dz * mask * gelu'(pre-activation)
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int j = v * 32 + lane;
        dv[r * H + j] = dz[r * H + j] * drop_scale(seed, 0, (uint64_t)r * H + j, p_drop) * gelu_grad(xhat[r * H + j]);
      }
    return;
  }
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    float dy[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int j = v * 32 + lane;
      xh[v] = xhat[r * H + j];
      const float y = fmaf(xh[v], lnw[j], lnb[j]);
      dy[v] = dz[r * H + j] * drop_scale(seed, 0, (uint64_t)r * H + j, p_drop) * gelu_grad(y);
      gw[v] = fmaf(dy[v], xh[v], gw[v]);
      gb[v] += dy[v];
      const float dyw = dy[v] * lnw[j];
      s1 += dyw;
      s2 = fmaf(dyw, xh[v], s2);
    }
    s1 = warp_sum(s1) * (1.0f / H);
    s2 = warp_sum(s2) * (1.0f / H);
    const float rstd = rstd_in[r];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int j = v * 32 + lane;
      dv[r * H + j] = rstd * (dy[v] * lnw[j] - s1 - xh[v] * s2);
    }
  }
  block_col_atomic_add<NV>(dlnw, dlnb, gw, gb);
}

// final LayerNorm forward over rows of width D: xhat, rstd, Y = xhat*w + b
template <int D>
__global__ void __launch_bounds__(256)
ln_rows_fwd(const float* __restrict__ x, long long rows, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ xhat,
            float* __restrict__ rstd_out, float* __restrict__ y, int use_ln) {
  constexpr int NV = D / 32;
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * 8;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < NV; ++e) { v[e] = x[r * D + e * 32 + lane]; s += v[e]; }
    const float mean = warp_sum(s) * (1.0f / D);
    float q2 = 0.f;
#pragma unroll
    for (int e = 0; e < NV; ++e) { const float d = v[e] - mean; q2 = fmaf(d, d, q2); }
    const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.0f / D) + 1e-5f);
    if (lane == 0) rstd_out[r] = rstd;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int d = e * 32 + lane;
      const float xh = use_ln ? (v[e] - mean) * rstd : v[e];   // nn.Identity (09:210): Y = out
      xhat[r * D + d] = xh;
      y[r * D + d] = use_ln ? fmaf(xh, w[d], b[d]) : xh;
    }
  }
}

// dx = LayerNorm-backward(dy) (dy already includes both the pooled-context and the score paths)
template <int D>
__global__ void __launch_bounds__(256)
ln_rows_bwd(const float* __restrict__ dy, const float* __restrict__ xhat, const float* __restrict__ rstd_in, const float* __restrict__ w,
            long long rows, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int use_ln) {
  constexpr int NV = D / 32;
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * 8;
  float gw[NV] = {}, gb[NV] = {};
  if (!use_ln) {
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride)
#pragma unroll
      for (int e = 0; e < NV; ++e) dx[r * D + e * 32 + lane] = dy[r * D + e * 32 + lane];
    return;
  }
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    float g[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int d = e * 32 + lane;
      g[e] = dy[r * D + d];
      xh[e] = xhat[r * D + d];
      gw[e] = fmaf(g[e], xh[e], gw[e]);
      gb[e] += g[e];
      const float gwv = g[e] * w[d];
      s1 += gwv;
      s2 = fmaf(gwv, xh[e], s2);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
    const float rstd = rstd_in[r];
#pragma unroll
    for (int e = 0; e < NV; ++e) {
      const int d = e * 32 + lane;
      dx[r * D + d] = rstd * (g[e] * w[d] - s1 - xh[e] * s2);
    }
  }
  block_col_atomic_add<NV>(dw, db, gw, gb);
}

// ---- attention pooling (train): one CTA per window -------------------------------------------------
// forward: s_t = b2 + sum_j w2_j tanh(PRE[t][b][j]); a = softmax_T(s); ctx = sum_t a_t Y[t][b][:]
// D = LSTM output width, AH = D/2 = attention hidden width; pre == nullptr: mean pooling (09:232-234, a_t = 1/T)
__global__ void __launch_bounds__(256)
attn_train_fwd(const float* __restrict__ pre, const float* __restrict__ y, int Bc, int T, int D, int AH, const float* __restrict__ w2,
               const float* __restrict__ b2, float* __restrict__ attn, float* __restrict__ ctx) {
  extern __shared__ float at_smem[];  // [T] scores -> weights
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // four time steps per warp iteration, all their loads issued before the first tanh: the rows of one window are 512 KB apart, and
  // one dependent load per step made this loop (and the context sum below) pure DRAM latency
  const int nq = AH / 32;   // <= 8
  float w2r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) w2r[q] = (pre && q < nq) ? w2[lane + 32 * q] : 0.f;
  for (int t0 = warp * 4; t0 < T; t0 += 32) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        v[u][q] = (pre && t0 + u < T && q < nq) ? pre[((long long)(t0 + u) * Bc + b) * AH + lane + 32 * q] : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fmaf(w2r[q], tanhf(v[u][q]), s);
      s = warp_sum(s) + (pre ? b2[0] : 0.f);
      if (lane == 0 && t0 + u < T) at_smem[t0 + u] = pre ? s : 0.f;
    }
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < T; t += 256) m = fmaxf(m, at_smem[t]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int t = tid; t < T; t += 256) { const float e = expf(at_smem[t] - m); at_smem[t] = e; l += e; }
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
  for (int w = 0; w < 8; ++w) l += red[w];
  const float inv = 1.0f / l;
  for (int t = tid; t < T; t += 256) { const float a = at_smem[t] * inv; at_smem[t] = a; attn[(long long)b * T + t] = a; }
  __syncthreads();
  for (int d = tid; d < D; d += 256) {
    float c = 0.f;
    int t = 0;
    for (; t + 8 <= T; t += 8) {
      float yv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) yv[u] = y[((long long)(t + u) * Bc + b) * D + d];
#pragma unroll
      for (int u = 0; u < 8; ++u) c = fmaf(at_smem[t + u], yv[u], c);
    }
    for (; t < T; ++t) c = fmaf(at_smem[t], y[((long long)t * Bc + b) * D + d], c);
    ctx[(long long)b * D + d] = c;
  }
}

// backward: from dctx -> dY (context path only), dPRE, dw2, db2.  `pre` is left intact so the same saved forward can be
// back-propagated repeatedly (07_explainability.py:252 calls backward(retain_graph=True) once per sample).
__global__ void __launch_bounds__(256)
attn_train_bwd(const float* __restrict__ pre, float* __restrict__ dpre, const float* __restrict__ y, const float* __restrict__ attn, const float* __restrict__ dctx,
               int Bc, int T, int D, int AH, const float* __restrict__ w2, float* __restrict__ dY, float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ float ab_smem[];  // [T] da -> ds ; [D] dctx
  float* ds = ab_smem;
  float* dc = ab_smem + T;
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < D; d += 256) dc[d] = dctx[(long long)b * D + d];
  __syncthreads();
  // da_t = dctx . Y_t ; dY_t = a_t * dctx
  const int nd = D / 32;   // <= 16
  for (int t0 = warp * 2; t0 < T; t0 += 16) {   // two time steps per warp iteration, loads first
    float yv[2][16];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int q = 0; q < 16; ++q) yv[u][q] = (t0 + u < T && q < nd) ? y[((long long)(t0 + u) * Bc + b) * D + lane + 32 * q] : 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t0 + u >= T) break;
      const long long row = (long long)(t0 + u) * Bc + b;
      const float a = attn[(long long)b * T + t0 + u];
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (q < nd) {
          const float dcv = dc[lane + 32 * q];
          s = fmaf(dcv, yv[u][q], s);
          dY[row * D + lane + 32 * q] = a * dcv;
        }
      }
      s = warp_sum(s);
      if (lane == 0) ds[t0 + u] = s;
    }
  }
  if (!pre) return;  // mean pooling: a_t = 1/T is a constant, dY = dctx / T is all there is
  __syncthreads();
  float dot = 0.f;
  for (int t = tid; t < T; t += 256) dot = fmaf(attn[(long long)b * T + t], ds[t], dot);
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = 0.f;
  for (int w = 0; w < 8; ++w) dot += red[w];
  __syncthreads();
  float sb2 = 0.f;
  for (int t = tid; t < T; t += 256) { const float v = attn[(long long)b * T + t] * (ds[t] - dot); ds[t] = v; sb2 += v; }
  sb2 = warp_sum(sb2);
  if (lane == 0) red[warp] = sb2;
  __syncthreads();
  if (tid == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; atomicAdd(db2, s); }
  __syncthreads();   // ds is final
  // dPRE[t][j] = ds_t w2_j (1 - u^2), u = tanh(pre);  dw2_j += sum_t ds_t u.  A warp per time step (two in flight), a lane per 32-strided
  // column: the first version gave every column j to ONE thread that walked the 256 steps with dependent loads 512 KB apart
  const int nq = AH / 32;   // <= 8
  float gacc[8], w2r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { gacc[q] = 0.f; w2r[q] = q < nq ? w2[lane + 32 * q] : 0.f; }
  for (int t0 = warp * 2; t0 < T; t0 += 16) {
    float pv[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int q = 0; q < 8; ++q) pv[u][q] = (t0 + u < T && q < nq) ? pre[((long long)(t0 + u) * Bc + b) * AH + lane + 32 * q] : 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t0 + u >= T) break;
      const float dst = ds[t0 + u];
      const long long o = ((long long)(t0 + u) * Bc + b) * AH;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < nq) {
          const float uu = tanhf(pv[u][q]);
          gacc[q] = fmaf(dst, uu, gacc[q]);
          dpre[o + lane + 32 * q] = dst * w2r[q] * (1.0f - uu * uu);
        }
      }
    }
  }
  float* gred = ab_smem + T + D;   // [8 warps][AH]
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < nq) gred[warp * AH + lane + 32 * q] = gacc[q];
  __syncthreads();
  for (int j = tid; j < AH; j += 256) {
    float g = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) g += gred[w * AH + j];
    atomicAdd(dw2 + j, g);
  }
}

// ---- classifier head (train): one CTA per window ------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(H)
head_train_fwd(const float* __restrict__ ctx, int classes, const float* __restrict__ c0t, const float* __restrict__ cb0,
               const float* __restrict__ c3t, const float* __restrict__ cb3, const float* __restrict__ c6, const float* __restrict__ cb6,
               float* __restrict__ pre1, float* __restrict__ h1d, float* __restrict__ pre2, float* __restrict__ h2d,
               float* __restrict__ logits, float* __restrict__ probs, float p_drop, uint64_t seed, int D) {
  __shared__ float cs[2 * H], h1[H], h2[H / 2], lg[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int d = tid; d < D; d += H) cs[d] = ctx[(long long)b * D + d];
  __syncthreads();
  if (tid < H) {
    float a = cb0[tid];
    for (int d = 0; d < D; ++d) a = fmaf(cs[d], c0t[(long long)d * H + tid], a);
    pre1[(long long)b * H + tid] = a;
    const float v = gelu_erf(a) * drop_scale(seed, 2, (uint64_t)b * H + tid, p_drop);
    h1[tid] = v;
    h1d[(long long)b * H + tid] = v;
  }
  __syncthreads();
  if (tid < H / 2) {
    float a = cb3[tid];
    for (int k = 0; k < H; ++k) a = fmaf(h1[k], c3t[k * (H / 2) + tid], a);
    pre2[(long long)b * (H / 2) + tid] = a;
    const float v = gelu_erf(a) * drop_scale(seed, 3, (uint64_t)b * (H / 2) + tid, p_drop);
    h2[tid] = v;
    h2d[(long long)b * (H / 2) + tid] = v;
  }
  __syncthreads();
  if (tid < classes) {
    float a = cb6[tid];
    for (int k = 0; k < H / 2; ++k) a = fmaf(h2[k], c6[tid * (H / 2) + k], a);
    logits[(long long)b * classes + tid] = a;
    lg[tid] = a;
  }
  __syncthreads();
  if (probs && tid == 0) {
    float mx = -INFINITY, den = 0.f;
    for (int c = 0; c < classes; ++c) mx = fmaxf(mx, lg[c]);
    for (int c = 0; c < classes; ++c) den += expf(lg[c] - mx);
    for (int c = 0; c < classes; ++c) probs[(long long)b * classes + c] = expf(lg[c] - mx) / den;
  }
}

// dlogits -> dpre2, dpre1, dctx (weight gradients are TN GEMMs over the batch afterwards)
template <int H>
__global__ void __launch_bounds__(H)
head_train_bwd(const float* __restrict__ dlogits, int classes, const float* __restrict__ pre1, const float* __restrict__ pre2,
               const float* __restrict__ w6 /*[cls][H/2]*/, const float* __restrict__ w3 /*[H/2][H]*/, const float* __restrict__ w0 /*[H][2H]*/,
               float* __restrict__ dpre1, float* __restrict__ dpre2, float* __restrict__ dctx, float p_drop, uint64_t seed, int D) {
  __shared__ float dl[8], d2[H / 2], d1[H];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < classes) dl[tid] = dlogits[(long long)b * classes + tid];
  __syncthreads();
  if (tid < H / 2) {
    float g = 0.f;
    for (int c = 0; c < classes; ++c) g = fmaf(dl[c], w6[c * (H / 2) + tid], g);
    g *= drop_scale(seed, 3, (uint64_t)b * (H / 2) + tid, p_drop) * gelu_grad(pre2[(long long)b * (H / 2) + tid]);
    d2[tid] = g;
    dpre2[(long long)b * (H / 2) + tid] = g;
  }
  __syncthreads();
  if (tid < H) {
    float g = 0.f;
    for (int k = 0; k < H / 2; ++k) g = fmaf(d2[k], w3[k * H + tid], g);
    g *= drop_scale(seed, 2, (uint64_t)b * H + tid, p_drop) * gelu_grad(pre1[(long long)b * H + tid]);
    d1[tid] = g;
    dpre1[(long long)b * H + tid] = g;
  }
  __syncthreads();
  for (int d = tid; d < D; d += H) {
    float g = 0.f;
    for (int j = 0; j < H; ++j) g = fmaf(d1[j], w0[j * D + d], g);
    dctx[(long long)b * D + d] = g;
  }
}

// ---- BPTT through one LSTM layer ---------------------------------------------------------------------------
// grid = (window tiles, 2 directions), 256 threads, thread = (hidden unit j, group of 16 windows): the mirror of
// lstm_rec_f32.  Walks the direction's time order backwards; dG_t (gate-interleaved) goes to global for the weight /
// input GEMMs and, transposed, to shared memory for dh_{t-1} = dG_t . W_hh.
constexpr int BP_THREADS = 256;

// RES_U > 0: the first RES_U unit rows of the packed W_hh stay in shared memory behind the dG tile (see lstm_rec_f32's RES_K).
template <int H, int BP_WPT, int RES_U = 0>
__global__ void __launch_bounds__(BP_THREADS, 1)
lstm_bptt_f32(const float* __restrict__ dout,    // [T][Bc][ND*H]  grad wrt the layer output
              const float* __restrict__ gates,   // [T][Bc][ND][H][4] post-activation i,f,g,o
              const float* __restrict__ csave,   // [T][Bc][ND][H]
              const float* __restrict__ whh_bf,  // [unit][j][4 gates] (float4 per (unit, j)), forward direction
              const float* __restrict__ whh_br,  // reverse direction
              float* __restrict__ dG,            // [T][Bc][ND][H][4]
              float* __restrict__ dG_lo,         // optional: tf32 remainder of dG for the split-precision GEMMs (gemm_tf32x3.cu)
              int Bc, int T, int ND) {
  constexpr int GROUPS = BP_THREADS / H, MT = GROUPS * BP_WPT, GS = MT + 4;
  extern __shared__ __align__(16) float bp_smem[];  // [4H][GS]
  const int tid = threadIdx.x, j = tid % H, grp = tid / H;
  const int dir = blockIdx.y;
  const int b_base = blockIdx.x * MT + grp * BP_WPT;
  const float4* __restrict__ W = reinterpret_cast<const float4*>(dir ? whh_br : whh_bf);
  float dh_rec[BP_WPT], dc[BP_WPT];
#pragma unroll
  for (int w = 0; w < BP_WPT; ++w) { dh_rec[w] = 0.f; dc[w] = 0.f; }
  float4* wres = reinterpret_cast<float4*>(bp_smem + 4 * H * GS);  // [RES_U][H]
  if (RES_U > 0)
    for (int i = tid; i < RES_U * H; i += BP_THREADS) wres[i] = __ldg(W + i);
  // USPLIT (the training-size variant): the two thread groups of a CTA do not split the tile's 8 windows but the 128 unit rows of
  // the product dh[j][w] = sum_u dG[u][.][w] W[u][.][j] -- each thread covers all 8 windows for half of the rows (16-row blocks
  // alternate between the groups, so both have 48 resident and 16 streamed rows) and the halves are exchanged through 4 KB of
  // shared memory.  Per thread and step that is 64 weight reads + 512 broadcast dG reads instead of 128 + 512, a quarter fewer
  // shared-memory wavefronts on a kernel ncu showed at 79 % of that pipe.
  constexpr bool USPLIT = RES_U > 0 && GROUPS == 2 && BP_WPT == 4;
  static_assert(!USPLIT || RES_U == 96, "the u-split variant is written for 96 resident rows of 128");
  float4* xch = reinterpret_cast<float4*>(wres + RES_U * H);  // [2 destination groups][H]: partial dh of the other group's windows
  float own[BP_WPT];
  if (USPLIT) {
#pragma unroll
    for (int w = 0; w < BP_WPT; ++w) own[w] = 0.f;
    xch[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
  }

  // latency-bound variant (RES_U > 0): the saved activations of step s-1 are requested while the recurrent product of step s
  // runs (the elementwise part at the top of a step otherwise waits ~1 us for them); c(s-1) is step s's cprev, already here
  constexpr bool PRE = RES_U > 0;
  float4 pg[PRE ? BP_WPT : 1];
  float pc[PRE ? BP_WPT : 1], pcp[PRE ? BP_WPT : 1], pdo[PRE ? BP_WPT : 1];
  auto fetch = [&](int s, float4* g4, float* cc, float* cp, float* dd) {
    const int t = dir ? (T - 1 - s) : s;
    const int tp = dir ? (t + 1) : (t - 1);
#pragma unroll
    for (int w = 0; w < BP_WPT; ++w) {
      const int b = b_base + w;
      if (b < Bc) {
        const long long row = (long long)t * Bc + b;
        g4[w] = __ldg(reinterpret_cast<const float4*>(gates) + (row * ND + dir) * H + j);
        if (cc) cc[w] = __ldg(csave + (row * ND + dir) * H + j);
        cp[w] = (s > 0) ? __ldg(csave + (((long long)tp * Bc + b) * ND + dir) * H + j) : 0.f;
        dd[w] = __ldg(dout + row * (ND * H) + dir * H + j);
      }
    }
  };
  if (PRE) fetch(T - 1, pg, pc, pcp, pdo);
  for (int s = T - 1; s >= 0; --s) {
    const int t = dir ? (T - 1 - s) : s;               // time index of forward step s
    const int tp = dir ? (t + 1) : (t - 1);            // time index of forward step s-1 (previous state)
    float* dgs = bp_smem + (j * 4) * GS + grp * BP_WPT;
    if (USPLIT) {  // this group's windows: own half of the unit rows + the other group's half (written before the last barrier)
      const float4 o = xch[grp * H + j];
      dh_rec[0] = own[0] + o.x; dh_rec[1] = own[1] + o.y; dh_rec[2] = own[2] + o.z; dh_rec[3] = own[3] + o.w;
    }
    // rows that are not resident: requested now, consumed after the resident part of the product
    constexpr int TAIL_U = RES_U > 0 ? (USPLIT ? 16 : H - RES_U) : 1;
    float4 wt[TAIL_U];
    if (RES_U > 0) {
#pragma unroll
      for (int uu = 0; uu < TAIL_U; ++uu) wt[uu] = __ldg(W + (long long)(RES_U + (USPLIT ? grp * 16 : 0) + uu) * H + j);
    }
#pragma unroll
    for (int w = 0; w < BP_WPT; ++w) {
      const int b = b_base + w;
      float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < Bc) {
        const long long row = (long long)t * Bc + b;
        const float4 g = PRE ? pg[w] : reinterpret_cast<const float4*>(gates)[(row * ND + dir) * H + j];
        const float c = PRE ? pc[w] : csave[(row * ND + dir) * H + j];
        const float cprev = PRE ? pcp[w] : ((s > 0) ? csave[(((long long)tp * Bc + b) * ND + dir) * H + j] : 0.f);
        const float dh = (PRE ? pdo[w] : dout[row * (ND * H) + dir * H + j]) + dh_rec[w];
        const float tc = rec_tanh(c);  // the forward's own tanh(c) (h = o tanh(c))
        const float dct = fmaf(dh * g.w, 1.0f - tc * tc, dc[w]);
        dg.x = dct * g.z * g.x * (1.0f - g.x);          // d pre_i
        dg.y = dct * cprev * g.y * (1.0f - g.y);        // d pre_f
        dg.z = dct * g.x * (1.0f - g.z * g.z);          // d pre_g
        dg.w = dh * tc * g.w * (1.0f - g.w);            // d pre_o
        dc[w] = dct * g.y;
        reinterpret_cast<float4*>(dG)[(row * ND + dir) * H + j] = dg;
        if (dG_lo) {  // x - trunc_tf32(x), rounded to tf32: what split_tf32_kernel would produce in a separate pass over dG
          auto lo = [](float x) {
            const float rem = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            uint32_t r;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
            return __uint_as_float(r);
          };
          reinterpret_cast<float4*>(dG_lo)[(row * ND + dir) * H + j] = make_float4(lo(dg.x), lo(dg.y), lo(dg.z), lo(dg.w));
        }
      }
      dgs[0 * GS + w] = dg.x; dgs[1 * GS + w] = dg.y; dgs[2 * GS + w] = dg.z; dgs[3 * GS + w] = dg.w;
    }
    __syncthreads();
    float4 ng[PRE ? BP_WPT : 1];
    float ncp[PRE ? BP_WPT : 1], ndo[PRE ? BP_WPT : 1];
    if (PRE && s > 0) fetch(s - 1, ng, nullptr, ncp, ndo);
    // dh_rec[w] (unit j) = sum_n dG[w][n] * W_hh_b[n][j]
    float acc[BP_WPT];
#pragma unroll
    for (int w = 0; w < BP_WPT; ++w) acc[w] = 0.f;
    const float* gsrc = bp_smem + grp * BP_WPT;
    // 32 independent 16-byte weight loads (= 128 rows of W_hh) in flight per thread: with training batches the machine is not
    // full and the L2 round trips of this stream ARE the step time (the first version had 16 four-byte loads in flight:
    // 32 round trips per step, 31 us; this one pays H/32 round trips)
    auto fma_unit = [&](int u, const float4 w4) {
      const float wg[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int gsel = 0; gsel < 4; ++gsel) {
        const float4* gp = reinterpret_cast<const float4*>(gsrc + (u * 4 + gsel) * GS);
#pragma unroll
        for (int q = 0; q < BP_WPT / 4; ++q) {
          const float4 g4 = gp[q];
          acc[q * 4 + 0] = fmaf(g4.x, wg[gsel], acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(g4.y, wg[gsel], acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(g4.z, wg[gsel], acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(g4.w, wg[gsel], acc[q * 4 + 3]);
        }
      }
    };
    if (USPLIT) {
      float a8[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) a8[w] = 0.f;
      auto fma8 = [&](int u, const float4 w4) {
        const float wg[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int gsel = 0; gsel < 4; ++gsel) {
          const float4* gp = reinterpret_cast<const float4*>(bp_smem + (u * 4 + gsel) * GS);  // all 8 windows of the tile
          const float4 g0 = gp[0], g1 = gp[1];
          a8[0] = fmaf(g0.x, wg[gsel], a8[0]); a8[1] = fmaf(g0.y, wg[gsel], a8[1]);
          a8[2] = fmaf(g0.z, wg[gsel], a8[2]); a8[3] = fmaf(g0.w, wg[gsel], a8[3]);
          a8[4] = fmaf(g1.x, wg[gsel], a8[4]); a8[5] = fmaf(g1.y, wg[gsel], a8[5]);
          a8[6] = fmaf(g1.z, wg[gsel], a8[6]); a8[7] = fmaf(g1.w, wg[gsel], a8[7]);
        }
      };
#pragma unroll
      for (int blk = 0; blk < 3; ++blk) {  // resident 16-row blocks 2 blk + grp (< 6)
        const int u0 = (2 * blk + grp) * 16;
#pragma unroll 8
        for (int i = 0; i < 16; ++i) fma8(u0 + i, wres[(u0 + i) * H + j]);
      }
#pragma unroll
      for (int uu = 0; uu < 16; ++uu) fma8(RES_U + grp * 16 + uu, wt[uu]);  // streamed block 6 + grp
      // the other group's windows go to shared memory, this group's stay in registers until the halves meet after the barrier
      if (grp == 0) {
        own[0] = a8[0]; own[1] = a8[1]; own[2] = a8[2]; own[3] = a8[3];
        xch[1 * H + j] = make_float4(a8[4], a8[5], a8[6], a8[7]);
      } else {
        own[0] = a8[4]; own[1] = a8[5]; own[2] = a8[6]; own[3] = a8[7];
        xch[0 * H + j] = make_float4(a8[0], a8[1], a8[2], a8[3]);
      }
    } else if (RES_U > 0) {
#pragma unroll 8
      for (int u = 0; u < RES_U; ++u) fma_unit(u, wres[u * H + j]);
#pragma unroll
      for (int uu = 0; uu < TAIL_U; ++uu) fma_unit(RES_U + uu, wt[uu]);
    } else {
      for (int u0 = 0; u0 < H; u0 += 32) {
        float4 wv[32];
#pragma unroll
        for (int uu = 0; uu < 32; ++uu) wv[uu] = __ldg(W + (long long)(u0 + uu) * H + j);
#pragma unroll
        for (int uu = 0; uu < 32; ++uu) fma_unit(u0 + uu, wv[uu]);
      }
    }
    if (!USPLIT) {
#pragma unroll
      for (int w = 0; w < BP_WPT; ++w) dh_rec[w] = acc[w];
    }
    if (PRE && s > 0) {
#pragma unroll
      for (int w = 0; w < BP_WPT; ++w) { pc[w] = pcp[w]; pg[w] = ng[w]; pcp[w] = ncp[w]; pdo[w] = ndo[w]; }
    }
    __syncthreads();
  }
}

// ---- small elementwise helpers ---------------------------------------------------------------------------------
// dst = dropout(src); lo (optional) = tf32 remainder of dst for the split-precision GEMM that consumes it next
__global__ void scale_mask_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n, float p, uint64_t seed, uint32_t site,
                                  float* __restrict__ lo = nullptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i] * drop_scale(seed, site, (uint64_t)i, p);
  dst[i] = v;
  if (lo) {
    const float rem = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(rem));
    lo[i] = __uint_as_float(r);
  }
}
// interleaved rows (n = unit*4 + gate, optionally + dir*4H) -> reference gate-major rows; dst (4H, K)
__global__ void unpack_gate_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int K, int row0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int gate = row / H, unit = row - gate * H;
  dst[i] = src[(long long)(row0 + unit * 4 + gate) * K + k];
}
__global__ void unpack_bias_kernel(const float* __restrict__ src, float* __restrict__ d_ih, float* __restrict__ d_hh, int H, int row0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H) return;
  const int gate = i / H, unit = i - gate * H;
  const float v = src[row0 + unit * 4 + gate];
  d_ih[i] = v;
  d_hh[i] = v;
}
// dst[r][c] = src[r][c] for c < cols (row strides lds / ldd)
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  dst[r * ldd + c] = src[r * lds + c];
}
// dxT [T][Bc][C] -> dx (Bc,T,C)
__global__ void untranspose_x_kernel(const float* __restrict__ src, float* __restrict__ dst, int Bc, int T, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Bc * T * C) return;
  const int c = (int)(i % C);
  const long long bt = i / C;
  const int t = (int)(bt % T), b = (int)(bt / T);
  dst[i] = src[((long long)t * Bc + b) * C + c];
}

// ---- workspace -------------------------------------------------------------------------------------------------------
struct TrainWs {
  float *xT, *xhat0, *rstd0, *z;
  float *gates[BCI_MAX_LAYERS], *cst[BCI_MAX_LAYERS], *out[BCI_MAX_LAYERS], *outd[BCI_MAX_LAYERS];
  float *G;                 // forward: projected inputs; backward: dG of layers l = L-1, L-3, ...
  float *G2;                // backward: dG of layers L-2, L-4, ... (the weight-gradient GEMMs of layer l read dG(l) on a side stream
                            // while BPTT of layer l-1 writes the other buffer)
  float *tmpW2;             // scratch of the side stream
  float *xhatF, *rstdF, *Y, *PRE, *attn, *ctx, *pre1, *h1d, *pre2, *h2d;
  float *dbias[2];          // bias gradients of the tensor-core BPTT (summed inside the kernel), one per dG buffer
  float *dA, *dB;           // [M][2H] gradient ping-pong
  float *dctx, *dpre1, *dpre2, *tmpW, *hdr;
  // tf32 remainders for the split-precision tcgen05 GEMMs: layer input (main stream), dG (one per dG buffer), and the side
  // stream's copies of the layer input / output
  float *lo_in, *lo_G, *lo_G2, *lo_in2, *lo_out2;
  size_t total;
};
// mode: 1 = the forward ran the mixed step (its saved gate activations are fp16: the backward must be the mixed one too), 0 = fp32-parity
struct TrainHeader { float dropout; uint32_t valid; uint64_t seed; int batch, T, mode; };

static void carve_train(const bci_lstm_config& c, int B, int T, float p_drop, char* base, TrainWs& w) {
  const size_t H = c.hidden_size, D = feat_width(c), C = c.input_size, M = (size_t)B * T;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = reinterpret_cast<float*>(base + off); off += align_up(n * sizeof(float), 256); return p; };
  w.hdr = take(64);
  w.xT = take(M * K1T_XS); w.xhat0 = take(M * H); w.rstd0 = take(M); w.z = take(M * H);
  for (int l = 0; l < c.num_layers; ++l) {
    w.gates[l] = take(M * 4 * D); w.cst[l] = take(M * D); w.out[l] = take(M * D);
    w.outd[l] = (p_drop > 0.f && l < c.num_layers - 1) ? take(M * D) : w.out[l];
  }
  w.G = take(M * 4 * D);
  w.G2 = take(M * 4 * D);
  w.dbias[0] = take(4 * D); w.dbias[1] = take(4 * D);
  w.xhatF = take(M * D); w.rstdF = take(M); w.Y = take(M * D); w.PRE = take(M * (D / 2));
  w.attn = take((size_t)B * T); w.ctx = take(B * D); w.pre1 = take(B * H); w.h1d = take(B * H);
  w.pre2 = take(B * (H / 2)); w.h2d = take(B * (H / 2));
  w.dA = take(M * D); w.dB = take(M * D);
  w.dctx = take(B * D); w.dpre1 = take(B * H); w.dpre2 = take(B * (H / 2));
  w.tmpW = take(4 * D * (D > H ? D : H) + 1024);
  w.tmpW2 = take(4 * D * (D > H ? D : H) + 1024);
  if (tf32x3_enabled()) {
    w.lo_in = take(M * D); w.lo_G = take(M * 4 * D); w.lo_G2 = take(M * 4 * D); w.lo_in2 = take(M * D); w.lo_out2 = take(M * D);
  } else {
    w.lo_in = w.lo_G = w.lo_G2 = w.lo_in2 = w.lo_out2 = nullptr;
  }
  w.total = off;
}

size_t lstm_workspace_train(const bci_lstm_config& c, int batch, int T) {
  TrainWs w;
  carve_train(c, batch > 0 ? batch : 1, T, 0.5f, nullptr, w);  // worst case: dropout copies present
  return w.total;
}

template <int H, int ND>
static int forward_train_t(bci_lstm_s* h, const float* x, int B, int T, float p_drop, uint64_t seed, float* logits, float* probs,
                           float* attn, TrainWs& w, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const PackedF32& p = h->f32;
  constexpr int D = ND * H, AH = D / 2;
  const int C = c.input_size, use_ln = c.use_layer_norm;
  const long long M = (long long)B * T;
  const int rb = (int)((M + 7) / 8 < 4096 ? (M + 7) / 8 : 4096);
  // mixed-precision step: the recurrences run on the tensor cores with 16-bit operands (lstm_rec_swap.cu)
  const bool mixed = h->last_mode == 1;   // decided by lstm_forward_train (recorded in the workspace header for the backward)
  // fp32-parity step: the same kernel in its split-precision form (three fp16 product chains, fp32-grade); BCI_TRAIN_REC=simt keeps
  // the CUDA-core recurrence of round 1
  const bool split_fwd = !mixed && swap_rec_enabled() && rec_swap_ok(H, w.G, 4 * D);
  if (mixed || split_fwd) {
    int rc = pack_swap_operands(h, st);
    if (rc) return rc;
  }
  {
    static PerDeviceFlag k1_attr_pd;
    bool& k1_attr = k1_attr_pd.cur();
    if (!k1_attr) {
      BCI_CUDA_OK(cudaFuncSetAttribute(inproj_train_fwd<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * H * (int)sizeof(float)));
      k1_attr = true;
    }
  }
  const int k1_blocks = (int)((M + 31) / 32 < 4LL * sm_count() ? (M + 31) / 32 : 4LL * sm_count());
  inproj_train_fwd<H><<<k1_blocks, 256, (size_t)C * H * sizeof(float), st>>>(x, B, T, C, p.w0t, p.b0, p.ln0w, p.ln0b, w.xT, w.xhat0, w.rstd0, w.z, p_drop * 0.5f, seed, use_ln);
  BCI_LAUNCH_OK();
  const float* in = w.z;
  bool lo_ready = false;  // w.lo_in already holds the remainder of `in` (written by the dropout kernel of the layer below)
  for (int l = 0; l < c.num_layers; ++l) {
    const int K = layer_in_width(c, l);
    int rc;
    bool g_half = false;
    if (tf32x3_nt_ok(in, K, p.wih_b[l], K, w.G, 4 * D, (int)M, 4 * D, K)) {
      // mixed mode: one TF32 pass over the raw operands, no remainders, and G written as fp16 (the recurrence that reads it is
      // HBM-bound) when the CTA-pair kernel takes the shape
      g_half = mixed && gemm_tf32_half_ok(in, K, p.wih_b[l], K, w.G, 4 * D, (int)M, 4 * D, K);
      if (!mixed && !lo_ready && (rc = split_tf32(in, nullptr, w.lo_in, M * K, st))) return rc;
      if (g_half)
        rc = gemm_tf32_nt_half(in, K, p.wih_b[l], K, p.bias[l], reinterpret_cast<__half*>(w.G), 4 * D, (int)M, 4 * D, K, st);
      else
        rc = gemm_tf32x3_nt(in, mixed ? nullptr : w.lo_in, K, p.wih_b[l], mixed ? nullptr : p.wih_b_lo[l], K, p.bias[l], w.G, 4 * D, (int)M,
                            4 * D, K, 0, st);
    } else {
      rc = launch_proj_gemm_f32(in, p.wih_t[l], p.bias[l], w.G, (int)M, 4 * D, K, st);
    }
    if (rc) return rc;
    // tensor-core recurrences write the dropped copy of the layer output themselves (no separate pass)
    const bool fused_drop = (mixed || split_fwd) && w.outd[l] != w.out[l];
    const SwapDropout fdrop{w.outd[l], mixed ? nullptr : w.lo_in, p_drop, seed, (uint32_t)(16 + l)};
    if (mixed && H == 256) rc = launch_rec_swap256_fwd(ND, w.G, 4 * D, p.whh_sw_f[l], w.out[l], w.gates[l], w.cst[l], D, B, T, st, fused_drop ? &fdrop : nullptr, g_half);
    else if (mixed || split_fwd) rc = launch_rec_swap_fwd(ND, w.G, 4 * D, p.whh_sw_f[l], w.out[l], w.gates[l], w.cst[l], D, B, T, split_fwd, st, fused_drop ? &fdrop : nullptr, g_half);
    else rc = launch_rec_f32(H, ND, w.G, p.whh_t[l][0], p.whh_t[l][1], w.out[l], w.gates[l], w.cst[l], B, T, st);
    if (rc) return rc;
    lo_ready = fused_drop && !mixed && w.lo_in != nullptr;
    if (w.outd[l] != w.out[l] && !fused_drop) {
      scale_mask_kernel<<<(unsigned)ceil_div64(M * D, 256), 256, 0, st>>>(w.out[l], w.outd[l], M * D, p_drop, seed, 16 + l,
                                                                          mixed ? nullptr : w.lo_in);
      BCI_LAUNCH_OK();
      lo_ready = w.lo_in != nullptr && !mixed;
    }
    in = w.outd[l];
  }
  const float* seq = w.out[c.num_layers - 1];
  ln_rows_fwd<D><<<rb, 256, 0, st>>>(seq, M, p.lnw, p.lnb, w.xhatF, w.rstdF, w.Y, use_ln);
  BCI_LAUNCH_OK();
  int rc = BCI_OK;
  if (c.use_attention) {
    if (tf32x3_nt_ok(w.Y, D, p.aw1, D, w.PRE, AH, (int)M, AH, D)) {
      if (!mixed && (rc = split_tf32(w.Y, nullptr, w.lo_in, M * D, st))) return rc;
      rc = gemm_tf32x3_nt(w.Y, mixed ? nullptr : w.lo_in, D, p.aw1, mixed ? nullptr : p.aw1_lo, D, p.ab1, w.PRE, AH, (int)M, AH, D, 0, st);
    } else {
      rc = gemm_nn(w.Y, D, p.aw1t, AH, w.PRE, AH, (int)M, AH, D, p.ab1, 0, st);
    }
    if (rc) return rc;
  }
  attn_train_fwd<<<B, 256, T * sizeof(float), st>>>(c.use_attention ? w.PRE : nullptr, w.Y, B, T, D, AH, p.aw2, p.ab2, w.attn, w.ctx);
  BCI_LAUNCH_OK();
  if (attn) BCI_CUDA_OK(cudaMemcpyAsync(attn, w.attn, (size_t)B * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  head_train_fwd<H><<<B, H, 0, st>>>(w.ctx, c.num_classes, p.c0t, p.cb0, p.c3t, p.cb3, p.c6, p.cb6, w.pre1, w.h1d, w.pre2, w.h2d,
                                       logits, probs, p_drop, seed, D);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int lstm_forward_train(bci_lstm_s* h, const float* x, int batch, int T, float dropout, uint64_t seed, float* logits, float* probs,
                       float* attn, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  BCI_REQUIRE(batch <= 4 * max_chunk(c, 1), BCI_EINVAL, "bci_lstm_forward(train=1): batch %d exceeds the training limit %d", batch,
              4 * max_chunk(c, 1));
  TrainWs w;
  carve_train(c, batch, T, dropout, (char*)ws, w);
  BCI_REQUIRE(ws_bytes >= w.total, BCI_ENOMEM, "bci_lstm_forward(train=1): workspace %zu < %zu bytes", ws_bytes, w.total);
  BCI_REQUIRE(T * sizeof(float) + 2 * c.hidden_size * sizeof(float) <= 40 * 1024, BCI_EINVAL, "training supports seq_len <= 8192");
  const int G4 = 4 * feat_width(c);
  const int mode = (h->train_mode == BCI_TRAIN_MIXED && (rec_swap_ok(c.hidden_size, w.G, G4) || rec_swap256_ok(c.hidden_size, w.G, G4))) ? 1 : 0;
  TrainHeader hd{dropout, 0xB200C0DEu, seed, batch, T, mode};
  BCI_CUDA_OK(cudaMemcpyAsync(w.hdr, &hd, sizeof(hd), cudaMemcpyHostToDevice, st));
  h->last_train_ws = ws; h->last_dropout = dropout; h->last_seed = seed; h->last_batch = batch; h->last_T = T; h->last_mode = mode;
  if (c.bidirectional)
    return c.hidden_size == 128 ? forward_train_t<128, 2>(h, x, batch, T, dropout, seed, logits, probs, attn, w, st)
                                : forward_train_t<256, 2>(h, x, batch, T, dropout, seed, logits, probs, attn, w, st);
  return c.hidden_size == 128 ? forward_train_t<128, 1>(h, x, batch, T, dropout, seed, logits, probs, attn, w, st)
                              : forward_train_t<256, 1>(h, x, batch, T, dropout, seed, logits, probs, attn, w, st);
}

template <int H, int ND>
static int backward_t(bci_lstm_s* h, const float* dlogits, int B, int T, float p_drop, uint64_t seed, int fwd_mode, float* dx,
                      const bci_lstm_grads* g, TrainWs& w, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const PackedF32& p = h->f32;
  const bci_lstm_weights& raw = h->raw;
  constexpr int D = ND * H, AH = D / 2, G4 = 4 * D;  // G4: gate columns of all directions
  const int C = c.input_size, L = c.num_layers, cls = c.num_classes, use_ln = c.use_layer_norm;
  const long long M = (long long)B * T;
  const int rb = (int)((M + 7) / 8 < 8LL * sm_count() ? (M + 7) / 8 : 8LL * sm_count());   // row kernels that end in column atomics: few, long blocks
  auto zero = [&](float* ptr, size_t n) { return cudaMemsetAsync(ptr, 0, n * sizeof(float), st); };
  auto zero_on = [&](cudaStream_t s2, float* ptr, size_t n) { return cudaMemsetAsync(ptr, 0, n * sizeof(float), s2); };
  int rc;
  // the mode of the forward that filled this workspace (its saved gate activations are fp16 in the mixed mode)
  const bool mixed = fwd_mode == 1;
  if (mixed && (rc = pack_swap_operands(h, st))) return rc;   // (no-op unless the weights were reloaded since the forward)
  const bool split_bwd = !mixed && swap_rec_enabled() && rec_swap_ok(H, w.G, G4) && !h->sw_stale;   // fp32-parity BPTT on the tensor cores
  // ---- head ----
  head_train_bwd<H><<<B, H, 0, st>>>(dlogits, cls, w.pre1, w.pre2, raw.cls_w6, raw.cls_w3, raw.cls_w0, w.dpre1, w.dpre2, w.dctx,
                                       p_drop, seed, D);
  BCI_LAUNCH_OK();
  BCI_CUDA_OK(zero(g->cls_w6, (size_t)cls * (H / 2))); BCI_CUDA_OK(zero(g->cls_b6, cls));
  BCI_CUDA_OK(zero(g->cls_w3, (size_t)(H / 2) * H));   BCI_CUDA_OK(zero(g->cls_b3, H / 2));
  BCI_CUDA_OK(zero(g->cls_w0, (size_t)H * D));         BCI_CUDA_OK(zero(g->cls_b0, H));
  if ((rc = gemm_tn(dlogits, cls, w.h2d, H / 2, g->cls_w6, H / 2, B, cls, H / 2, st))) return rc;
  if ((rc = colsum(dlogits, cls, B, cls, g->cls_b6, st))) return rc;
  if ((rc = gemm_tn(w.dpre2, H / 2, w.h1d, H, g->cls_w3, H, B, H / 2, H, st))) return rc;
  if ((rc = colsum(w.dpre2, H / 2, B, H / 2, g->cls_b3, st))) return rc;
  if ((rc = gemm_tn(w.dpre1, H, w.ctx, D, g->cls_w0, D, B, H, D, st))) return rc;
  if ((rc = colsum(w.dpre1, H, B, H, g->cls_b0, st))) return rc;
  // ---- attention pooling ----
  float* dPRE = w.dB;  // [M][AH] lives in dB until the LayerNorm backward below overwrites it (no longer needed then)
  if (c.use_attention) {
    BCI_CUDA_OK(zero(g->attn_w2, AH)); BCI_CUDA_OK(zero(g->attn_b2, 1));
    BCI_CUDA_OK(zero(g->attn_w1, (size_t)AH * D)); BCI_CUDA_OK(zero(g->attn_b1, AH));
  }
  if (use_ln) { BCI_CUDA_OK(zero(g->ln_w, D)); BCI_CUDA_OK(zero(g->ln_b, D)); }
  attn_train_bwd<<<B, 256, (T + D + 8 * AH) * sizeof(float), st>>>(c.use_attention ? w.PRE : nullptr, dPRE, w.Y, w.attn, w.dctx, B, T, D, AH, p.aw2,
                                                          w.dA /*dY*/, g->attn_w2, g->attn_b2);
  BCI_LAUNCH_OK();
  if (c.use_attention) {
    // dY += dPRE . W1 ; dW1 = dPRE^T Y ; db1 = colsum(dPRE)
    if (tf32x3_nt_ok(dPRE, AH, p.aw1t, AH, w.dA, D, (int)M, D, AH) && tf32x3_tn_ok(dPRE, AH, w.Y, D, g->attn_w1, D, M, AH, D)) {
      // lo_in / lo_out2 are free here: the forward is over and the side stream starts after the top layer's BPTT
      if (!mixed && (rc = split_tf32(dPRE, nullptr, w.lo_in, M * AH, st))) return rc;
      if (!mixed && (rc = split_tf32(w.Y, nullptr, w.lo_out2, M * D, st))) return rc;
      if ((rc = gemm_tf32x3_nt(dPRE, mixed ? nullptr : w.lo_in, AH, p.aw1t, mixed ? nullptr : p.aw1t_lo, AH, nullptr, w.dA, D, (int)M, D, AH, 1, st))) return rc;
      if ((rc = gemm_tf32x3_tn(dPRE, mixed ? nullptr : w.lo_in, AH, w.Y, mixed ? nullptr : w.lo_out2, D, g->attn_w1, D, M, AH, D, st))) return rc;
    } else {
      if ((rc = gemm_nn(dPRE, AH, raw.attn_w1, D, w.dA, D, (int)M, D, AH, nullptr, 1, st))) return rc;
      if ((rc = gemm_tn(dPRE, AH, w.Y, D, g->attn_w1, D, M, AH, D, st))) return rc;
    }
    if ((rc = colsum(dPRE, AH, M, AH, g->attn_b1, st))) return rc;
  }
  // final LayerNorm backward: dA (dY) -> dB (grad wrt the last LSTM layer's output)
  ln_rows_bwd<D><<<rb, 256, 0, st>>>(w.dA, w.xhatF, w.rstdF, p.lnw, M, w.dB, g->ln_w, g->ln_b, use_ln);
  BCI_LAUNCH_OK();
  float* dcur = w.dB;   // grad wrt out[l]
  float* dnext = w.dA;  // scratch for grad wrt the layer input
  // ---- LSTM layers, top down ----
  // 16 windows per thread unless that leaves most SMs idle (typical training batches): then 8
  bool small = ND * ceil_div(B, (BP_THREADS / H) * 16) < sm_count();
  bool tiny = ND * ceil_div(B, (BP_THREADS / H) * 8) < sm_count();
  small_batch_policy(H, ND, B, BP_THREADS / H, small, tiny);
  const int MT = (BP_THREADS / H) * (tiny ? 4 : small ? 8 : 16);
  const size_t bp_smem = (size_t)4 * H * (MT + 4) * sizeof(float);
  // training batches at H = 128: 96 of the 128 unit rows of W_hh resident in shared memory behind the dG tile
  constexpr int BP_RES = H == 128 ? 96 : 0;
  constexpr size_t bp_res_bytes = (size_t)(BP_RES + 2) * H * sizeof(float4);  // resident rows + the u-split exchange buffer
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    if (H == 128)
      BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_f32<H, 4, BP_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)4 * H * ((BP_THREADS / H) * 4 + 4) * sizeof(float) + bp_res_bytes)));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_f32<H, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)4 * H * ((BP_THREADS / H) * 16 + 4) * sizeof(float))));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_f32<H, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)4 * H * ((BP_THREADS / H) * 4 + 4) * sizeof(float))));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_bptt_f32<H, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((size_t)4 * H * ((BP_THREADS / H) * 8 + 4) * sizeof(float))));
    attr = true;
  }
  // side stream: everything that only produces weight gradients of layer l (it needs dG(l), the layer's input and output, all
  // final by then) runs beside the BPTT recurrence of layer l-1, which at training batch sizes leaves most of the machine idle
  if (!h->side_ready) {
    BCI_CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    BCI_CUDA_OK(cudaEventCreateWithFlags(&h->ev_dg, cudaEventDisableTiming));
    BCI_CUDA_OK(cudaEventCreateWithFlags(&h->ev_side[0], cudaEventDisableTiming));
    BCI_CUDA_OK(cudaEventCreateWithFlags(&h->ev_side[1], cudaEventDisableTiming));
    BCI_CUDA_OK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->side_ready = true;
  }
  cudaStream_t sd = h->side;
  int used[2] = {0, 0};
  for (int l = L - 1; l >= 0; --l) {
    const int K = layer_in_width(c, l);
    const float* in = (l == 0) ? w.z : w.outd[l - 1];
    const int gb = (L - 1 - l) & 1;
    float* dGl = gb ? w.G2 : w.G;
    if (used[gb]) BCI_CUDA_OK(cudaStreamWaitEvent(st, h->ev_side[gb], 0));  // the side stream has finished reading this buffer
    // split-precision tensor-core GEMMs for this layer's gradients when the shapes allow (training batches do); BPTT then also
    // writes the tf32 remainder of dG (no separate pass over the 0.5 GB array)
    float* dGl_lo = gb ? w.lo_G2 : w.lo_G;
    const bool tc = tf32x3_tn_ok(dGl, G4, in, K, w.tmpW2, K, M, G4, K) && tf32x3_tn_ok(dGl, G4, w.out[l], D, w.tmpW2, H, M - B, 4 * H, H) &&
                    tf32x3_nt_ok(dGl, G4, p.wih_t[l], G4, dnext, K, (int)M, K, G4);
    if (mixed || split_bwd) BCI_CUDA_OK(zero(w.dbias[gb], (size_t)G4));
    // dcur is the gradient wrt this layer's DROPPED output when a dropout site follows it: the tensor-core BPTT kernels apply the mask
    const bool drop_here = (mixed || split_bwd) && l < L - 1 && w.outd[l] != w.out[l];
    const SwapDropout bdrop{nullptr, nullptr, drop_here ? p_drop : 0.f, seed, (uint32_t)(16 + l)};
    if (mixed && H == 256) {
      if ((rc = launch_bptt_swap256(ND, dcur, w.gates[l], w.cst[l], p.whh_sw_b[l], dGl, w.dbias[gb], G4, D, B, T, st, &bdrop))) return rc;
    } else if (mixed) {
      if ((rc = launch_bptt_swap(ND, dcur, w.gates[l], w.cst[l], p.whh_sw_b[l], dGl, nullptr, w.dbias[gb], G4, D, B, T, false, st, &bdrop))) return rc;
    } else if (split_bwd) {
      if ((rc = launch_bptt_swap(ND, dcur, w.gates[l], w.cst[l], p.whh_sw_b16[l], dGl, tc ? dGl_lo : nullptr, w.dbias[gb], G4, D, B, T, true, st, &bdrop))) return rc;
    } else if (tiny && H == 128)
      lstm_bptt_f32<H, 4, BP_RES><<<dim3(ceil_div(B, MT), ND), BP_THREADS, bp_smem + bp_res_bytes, st>>>(dcur, w.gates[l], w.cst[l], p.whh_b[l][0], p.whh_b[l][1], dGl, tc ? dGl_lo : nullptr, B, T, ND);
    else if (tiny)
      lstm_bptt_f32<H, 4><<<dim3(ceil_div(B, MT), ND), BP_THREADS, bp_smem, st>>>(dcur, w.gates[l], w.cst[l], p.whh_b[l][0], p.whh_b[l][1], dGl, tc ? dGl_lo : nullptr, B, T, ND);
    else if (small)
      lstm_bptt_f32<H, 8><<<dim3(ceil_div(B, MT), ND), BP_THREADS, bp_smem, st>>>(dcur, w.gates[l], w.cst[l], p.whh_b[l][0], p.whh_b[l][1], dGl, tc ? dGl_lo : nullptr, B, T, ND);
    else
      lstm_bptt_f32<H, 16><<<dim3(ceil_div(B, MT), ND), BP_THREADS, bp_smem, st>>>(dcur, w.gates[l], w.cst[l], p.whh_b[l][0], p.whh_b[l][1], dGl, tc ? dGl_lo : nullptr, B, T, ND);
    BCI_LAUNCH_OK();
    BCI_CUDA_OK(cudaEventRecord(h->ev_dg, st));
    BCI_CUDA_OK(cudaStreamWaitEvent(sd, h->ev_dg, 0));
    // dW_ih (all directions at once, interleaved rows) = dG^T . in
    if (tc) {
      if (!mixed && (rc = split_tf32(in, nullptr, w.lo_in2, M * K, sd))) return rc;
      if ((rc = gemm_tf32x3_tn(dGl, mixed ? nullptr : dGl_lo, G4, in, mixed ? nullptr : w.lo_in2, K, w.tmpW2, K, M, G4, K, sd))) return rc;
      if (!mixed && (rc = split_tf32(w.out[l], nullptr, w.lo_out2, M * D, sd))) return rc;
    } else {
      BCI_CUDA_OK(zero_on(sd, w.tmpW2, (size_t)G4 * K));
      if ((rc = gemm_tn(dGl, G4, in, K, w.tmpW2, K, M, G4, K, sd))) return rc;
    }
    for (int d = 0; d < ND; ++d) {
      unpack_gate_rows_kernel<<<(unsigned)ceil_div64((long long)4 * H * K, 256), 256, 0, sd>>>(w.tmpW2, g->w_ih[l][d], H, K, d * 4 * H);
      BCI_LAUNCH_OK();
    }
    // dW_hh[d] = dG[:, d]^T . h_prev, h_prev(t) = out[t-1] (forward) / out[t+1] (reverse): a row shift by Bc rows
    for (int d = 0; d < ND; ++d) {
      const long long R = M - B;
      const long long offA = d * 4 * H + (d == 0 ? (long long)B * G4 : 0), offB = d * H + (d == 0 ? 0 : (long long)B * D);
      const float* Ad = dGl + offA;
      const float* Bd = w.out[l] + offB;
      if (tc) {
        if ((rc = gemm_tf32x3_tn(Ad, mixed ? nullptr : dGl_lo + offA, G4, Bd, mixed ? nullptr : w.lo_out2 + offB, D, w.tmpW2, H, R, 4 * H, H, sd))) return rc;
      } else {
        BCI_CUDA_OK(zero_on(sd, w.tmpW2, (size_t)4 * H * H));
        if ((rc = gemm_tn(Ad, G4, Bd, D, w.tmpW2, H, R, 4 * H, H, sd))) return rc;
      }
      unpack_gate_rows_kernel<<<(unsigned)ceil_div64((long long)4 * H * H, 256), 256, 0, sd>>>(w.tmpW2, g->w_hh[l][d], H, H, 0);
      BCI_LAUNCH_OK();
    }
    // biases
    const float* bsrc = w.tmpW2;
    if (mixed || split_bwd) {
      bsrc = w.dbias[gb];   // summed by the BPTT kernel itself: no pass over dG
    } else {
      BCI_CUDA_OK(zero_on(sd, w.tmpW2, (size_t)G4));
      if ((rc = colsum(dGl, G4, M, G4, w.tmpW2, sd))) return rc;
    }
    for (int d = 0; d < ND; ++d) {
      unpack_bias_kernel<<<ceil_div(4 * H, 256), 256, 0, sd>>>(bsrc, g->b_ih[l][d], g->b_hh[l][d], H, d * 4 * H);
      BCI_LAUNCH_OK();
    }
    BCI_CUDA_OK(cudaEventRecord(h->ev_side[gb], sd));
    used[gb] = 1;
    // grad wrt the layer input: dnext [M][K] = dG . wih_b
    if (tc) {
      if ((rc = gemm_tf32x3_nt(dGl, mixed ? nullptr : dGl_lo, G4, p.wih_t[l], mixed ? nullptr : p.wih_t_lo[l], G4, nullptr, dnext, K, (int)M, K, G4, 0, st))) return rc;
    } else if ((rc = gemm_nn(dGl, G4, p.wih_b[l], K, dnext, K, (int)M, K, G4, nullptr, 0, st))) {
      return rc;
    }
    if (l > 0 && w.outd[l - 1] != w.out[l - 1] && !(mixed || split_bwd)) {   // (the tensor-core BPTT of layer l-1 applies this mask itself)
      scale_mask_kernel<<<(unsigned)ceil_div64(M * K, 256), 256, 0, st>>>(dnext, dnext, M * K, p_drop, seed, 16 + (l - 1));
      BCI_LAUNCH_OK();
    }
    float* tsw = dcur; dcur = dnext; dnext = tsw;
  }
  // ---- input projection: dcur = dz [M][H] ----
  if (use_ln) { BCI_CUDA_OK(zero(g->input_ln_w, H)); BCI_CUDA_OK(zero(g->input_ln_b, H)); }
  BCI_CUDA_OK(zero(g->input_proj_w, (size_t)H * C)); BCI_CUDA_OK(zero(g->input_proj_b, H));
  inproj_bwd_rows<H><<<rb, 256, 0, st>>>(dcur, w.xhat0, w.rstd0, p.ln0w, p.ln0b, M, dnext /*dv*/, g->input_ln_w, g->input_ln_b,
                                         p_drop * 0.5f, seed, use_ln);
  BCI_LAUNCH_OK();
  // dW0 (H, C) = dv^T . xT: on the tensor cores as a 128-column product -- xT rows are 64 floats (C <= 64, zero-padded), the TMA unit
  // zero-fills the rest of the tile -- into scratch, then the C valid columns are copied out (the gradient's rows are C floats apart:
  // no tensor map can describe them)
  if (tf32x3_tn_ok(dnext, H, w.xT, K1T_XS, w.tmpW, 128, M, H, 128)) {
    const bool sp = !mixed;
    // remainders into buffers the main stream owns and no longer needs (the final LayerNorm's xhat and output: the side stream may
    // still be reading the dG remainders)
    if (sp && (rc = split_tf32(dnext, nullptr, w.xhatF, M * H, st))) return rc;
    if (sp && (rc = split_tf32(w.xT, nullptr, w.Y, M * K1T_XS, st))) return rc;
    if ((rc = gemm_tf32x3_tn(dnext, sp ? w.xhatF : nullptr, H, w.xT, sp ? w.Y : nullptr, K1T_XS, w.tmpW, 128, M, H, 128, st, 0, K1T_XS))) return rc;
    copy_cols_kernel<<<ceil_div(H * C, 256), 256, 0, st>>>(w.tmpW, 128, g->input_proj_w, C, H, C);
    BCI_LAUNCH_OK();
  } else if ((rc = gemm_tn(dnext, H, w.xT, K1T_XS, g->input_proj_w, C, M, H, C, st))) {
    return rc;
  }
  if ((rc = colsum(dnext, H, M, H, g->input_proj_b, st))) return rc;
  if (dx) {
    // dxT [M][C] = dv . W0 (H x C); reuse dcur as scratch
    if ((rc = gemm_nn(dnext, H, raw.input_proj_w, C, dcur, C, (int)M, C, H, nullptr, 0, st))) return rc;
    untranspose_x_kernel<<<(unsigned)ceil_div64(M * C, 256), 256, 0, st>>>(dcur, dx, B, T, C);
    BCI_LAUNCH_OK();
  }
  // the caller's stream owns every gradient again
  BCI_CUDA_OK(cudaEventRecord(h->ev_join, sd));
  BCI_CUDA_OK(cudaStreamWaitEvent(st, h->ev_join, 0));
  return BCI_OK;
}

int lstm_backward_impl(bci_lstm_s* h, const float* x, const float* dlogits, int batch, int T, float* dx, const bci_lstm_grads* g,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)x;
  const bci_lstm_config& c = h->cfg;
  // the forward left its configuration in the workspace header
  TrainHeader hd;
  if (ws == h->last_train_ws && batch == h->last_batch && T == h->last_T) {
    // the usual case -- backward right after its forward: no stream synchronisation inside the training step
    hd = TrainHeader{h->last_dropout, 0xB200C0DEu, h->last_seed, batch, T, h->last_mode};
  } else {  // an older forward's workspace (07:252 keeps several alive): read what that forward left there
    BCI_CUDA_OK(cudaMemcpyAsync(&hd, ws, sizeof(hd), cudaMemcpyDeviceToHost, st));
    BCI_CUDA_OK(cudaStreamSynchronize(st));
  }
  BCI_REQUIRE(hd.valid == 0xB200C0DEu && hd.batch == batch && hd.T == T, BCI_ESTATE,
              "bci_lstm_backward: workspace does not hold a train=1 forward of this shape");
  TrainWs w;
  carve_train(c, batch, T, hd.dropout, (char*)ws, w);
  BCI_REQUIRE(ws_bytes >= w.total, BCI_ENOMEM, "bci_lstm_backward: workspace %zu < %zu bytes", ws_bytes, w.total);
  // every module the configuration has needs its gradient pointers
  BCI_REQUIRE(g->input_proj_w && g->input_proj_b && g->cls_w0 && g->cls_b0 && g->cls_w3 && g->cls_b3 && g->cls_w6 && g->cls_b6,
              BCI_EINVAL, "bci_lstm_backward: a projection / classifier gradient pointer is NULL");
  if (c.use_layer_norm)
    BCI_REQUIRE(g->input_ln_w && g->input_ln_b && g->ln_w && g->ln_b, BCI_EINVAL, "bci_lstm_backward: a LayerNorm gradient pointer is NULL");
  if (c.use_attention)
    BCI_REQUIRE(g->attn_w1 && g->attn_b1 && g->attn_w2 && g->attn_b2, BCI_EINVAL, "bci_lstm_backward: an attention gradient pointer is NULL");
  for (int l = 0; l < c.num_layers; ++l)
    for (int d = 0; d < num_dirs(c); ++d)
      BCI_REQUIRE(g->w_ih[l][d] && g->w_hh[l][d] && g->b_ih[l][d] && g->b_hh[l][d], BCI_EINVAL,
                  "bci_lstm_backward: LSTM gradient pointer NULL (layer %d dir %d)", l, d);
  if (c.bidirectional)
    return c.hidden_size == 128 ? backward_t<128, 2>(h, dlogits, batch, T, hd.dropout, hd.seed, hd.mode, dx, g, w, st)
                                : backward_t<256, 2>(h, dlogits, batch, T, hd.dropout, hd.seed, hd.mode, dx, g, w, st);
  return c.hidden_size == 128 ? backward_t<128, 1>(h, dlogits, batch, T, hd.dropout, hd.seed, hd.mode, dx, g, w, st)
                              : backward_t<256, 1>(h, dlogits, batch, T, hd.dropout, hd.seed, hd.mode, dx, g, w, st);
}

// ---- fused clip + AdamW ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(out, t);
  }
}
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float grad_scale, float max_norm,
                             float* __restrict__ norm_scratch) {
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1; AdamW (decoupled decay) as torch.optim.AdamW
  const float total = sqrtf(norm_scratch[0]) * fabsf(grad_scale);
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  if (blockIdx.x == 0 && threadIdx.x == 0) norm_scratch[1] = total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// weighted cross-entropy (mean over the batch's class weights) and its gradient wrt the logits: one block, every reduction in a
// fixed order (the loss is bit-reproducible from run to run); B is a training batch (hundreds of windows), classes <= 8
__global__ void __launch_bounds__(256)
ce_loss_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ cw, int B, int classes,
                    float loss_scale, float* __restrict__ loss_out, float* __restrict__ dlogits) {
  __shared__ float red_w[8], red_l[8];
  __shared__ float tot_w;
  float sw = 0.f, sl = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int y = (int)labels[i];
    float mx = -INFINITY;
    for (int c = 0; c < classes; ++c) mx = fmaxf(mx, logits[(long long)i * classes + c]);
    float den = 0.f;
    for (int c = 0; c < classes; ++c) den += expf(logits[(long long)i * classes + c] - mx);
    const float w = cw ? cw[y] : 1.f;
    sw += w;
    sl += w * (logf(den) + mx - logits[(long long)i * classes + y]);
  }
  sw = warp_sum(sw); sl = warp_sum(sl);
  if ((threadIdx.x & 31) == 0) { red_w[threadIdx.x >> 5] = sw; red_l[threadIdx.x >> 5] = sl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red_w[w]; b += red_l[w]; }
    tot_w = a;
    if (loss_out) *loss_out = loss_scale * b / a;
  }
  __syncthreads();
  const float inv = loss_scale / tot_w;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int y = (int)labels[i];
    float mx = -INFINITY;
    for (int c = 0; c < classes; ++c) mx = fmaxf(mx, logits[(long long)i * classes + c]);
    float den = 0.f;
    for (int c = 0; c < classes; ++c) den += expf(logits[(long long)i * classes + c] - mx);
    const float w = (cw ? cw[y] : 1.f) * inv;
    for (int c = 0; c < classes; ++c)
      dlogits[(long long)i * classes + c] = w * (expf(logits[(long long)i * classes + c] - mx) / den - (c == y ? 1.f : 0.f));
  }
}

__global__ void grad_accumulate_kernel(float* __restrict__ acc, const float* __restrict__ g, long long n, int first) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[i] = first ? g[i] : acc[i] + g[i];
}

}  // namespace bci

extern "C" int bci_ce_loss_grad(const float* logits, const int64_t* labels, const float* class_weight, int32_t batch, int32_t classes,
                                float loss_scale, float* loss_out, float* dlogits, void* stream) {
  using namespace bci;
  BCI_REQUIRE(logits && labels && dlogits && batch >= 1 && classes >= 1 && classes <= 8, BCI_EINVAL, "bci_ce_loss_grad: bad arguments");
  ce_loss_grad_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, reinterpret_cast<const long long*>(labels), class_weight, batch, classes,
                                                         loss_scale, loss_out, dlogits);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

extern "C" int bci_grad_accumulate(float* acc, const float* g, int64_t n, int32_t first, void* stream) {
  using namespace bci;
  BCI_REQUIRE(acc && g && n >= 0, BCI_EINVAL, "bci_grad_accumulate: bad arguments");
  if (n == 0) return BCI_OK;
  const int blocks = (int)((n + 1023) / 1024 < 4 * sm_count() ? (n + 1023) / 1024 : 4 * sm_count());
  grad_accumulate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(acc, g, n, first);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

extern "C" int bci_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int32_t step, float grad_scale, float max_norm, float* norm_scratch, void* stream) {
  bci::NvtxRange nvtx_range("bci_adamw_step");
  using namespace bci;
  BCI_REQUIRE(p && g && m && v && norm_scratch && n >= 0 && step >= 1, BCI_EINVAL, "bci_adamw_step: bad arguments");
  if (n == 0) return BCI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BCI_CUDA_OK(cudaMemsetAsync(norm_scratch, 0, 2 * sizeof(float), st));
  const int blocks = (int)((n + 1023) / 1024 < 4 * sm_count() ? (n + 1023) / 1024 : 4 * sm_count());
  sumsq_kernel<<<blocks, 256, 0, st>>>(g, n, norm_scratch);
  BCI_LAUNCH_OK();
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale, max_norm, norm_scratch);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// diagnostics: the stateless dropout mask (common.cuh: drop_scale) of elements [0, n) of one site, as the kernels evaluate it
__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, float p, uint64_t seed, uint32_t site) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bci::drop_scale(seed, site, (uint64_t)i, p);
}
extern "C" int bci_selftest_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint32_t site, void* stream) {
  using namespace bci;
  BCI_REQUIRE(out && n >= 0 && p >= 0.f && p < 1.f, BCI_EINVAL, "bci_selftest_dropout_mask: bad arguments");
  if (n == 0) return BCI_OK;
  dropout_mask_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(out, n, p, seed, site);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

