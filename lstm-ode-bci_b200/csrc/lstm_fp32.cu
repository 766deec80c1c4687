// fp32 (parity) mode of the BiLSTM forward: fp32 arithmetic everywhere (see rec_sigmoid / rec_tanh for the gate activations), so
// logits/probabilities match the reference's fp32 path to <= 1e-5 (north_star).  Plain TF32 / bf16 tensor-core math
// (10- / 8-bit mantissas) cannot meet that bound over 3 x 256 dependent steps: the recurrence stays on the FMA pipe and the
// time-parallel GEMMs use TF32 in split precision; the bf16 tcgen05 path lives in lstm_bf16*.cu.
//
//   K2  G[T*Bc][8H] = in[T*Bc][K] . W_ih^T (both directions) + (b_ih + b_hh): the split-precision tcgen05 GEMM of
//       gemm_tf32x3.cu (three TF32 MMAs per product, fp32-grade); proj_gemm_f32 below is the CUDA-core version it
//       replaced (BCI_FP32_GEMM=simt, and shapes with fewer than 128 rows)
//   K3  lstm_rec_f32  : persistent over the whole sequence per (window tile, direction):
//                       gates = G_t + h W_hh^T, sigma/tanh, cell update, h -> smem for step t+1
// Reference: nn.LSTM inside EnhancedLSTMModel (04_lstm_model.py:181-188,211).
#include "lstm_shared_kernels.cuh"
#include <cstdlib>

namespace bci {

// ---------------------------------------------------------------------------------------------
// K2: SGEMM with bias, C[M][N] = A[M][K] Bt[K][N] + bias[N].  128x128x16 tiles, 8x8 per thread,
// register-prefetched global loads, conflict-free split fragments.
// ---------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128, GK = 16, GEMM_THREADS = 256;

__global__ void __launch_bounds__(GEMM_THREADS)
proj_gemm_f32(const float* __restrict__ A, const float* __restrict__ Bt, const float* __restrict__ bias,
              float* __restrict__ C, int M, int N, int K, int accumulate) {
  __shared__ __align__(16) float As[2][GK][GM + 4];
  __shared__ __align__(16) float Bs[2][GK][GN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  // global load mapping: A tile 128 rows x 16 k: thread loads 2 float4 along k
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;  // rows a_row and a_row+64
  const int b_row = tid >> 5, b_nq = (tid & 31) * 4; // k rows b_row and b_row+8
  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = m0 + a_row + i * 64;
      ra[i] = (m < M) ? *reinterpret_cast<const float4*>(A + (long long)m * K + k0 + a_kq) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = *reinterpret_cast<const float4*>(Bt + (long long)(k0 + b_row + i * 8) * N + n0 + b_nq);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = a_row + i * 64;
      As[buf][a_kq + 0][m] = ra[i].x; As[buf][a_kq + 1][m] = ra[i].y;
      As[buf][a_kq + 2][m] = ra[i].z; As[buf][a_kq + 3][m] = ra[i].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_row + i * 8][b_nq]) = rb[i];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int nk = K / GK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
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
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4.., 64+tx*4..}
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 bia0 = bias ? *reinterpret_cast<const float4*>(bias + n0 + tx * 4) : zero4;
  const float4 bia1 = bias ? *reinterpret_cast<const float4*>(bias + n0 + 64 + tx * 4) : zero4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m < M) {
      float4* c0 = reinterpret_cast<float4*>(C + (long long)m * N + n0 + tx * 4);
      float4* c1 = reinterpret_cast<float4*>(C + (long long)m * N + n0 + 64 + tx * 4);
      const float4 p0 = accumulate ? *c0 : zero4, p1 = accumulate ? *c1 : zero4;
      *c0 = make_float4(acc[i][0] + bia0.x + p0.x, acc[i][1] + bia0.y + p0.y, acc[i][2] + bia0.z + p0.z, acc[i][3] + bia0.w + p0.w);
      *c1 = make_float4(acc[i][4] + bia1.x + p1.x, acc[i][5] + bia1.y + p1.y, acc[i][6] + bia1.z + p1.z, acc[i][7] + bia1.w + p1.w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3: recurrence.  grid = (window tiles, 2 directions); 256 threads; thread = (hidden unit j,
// group of 16 windows).  h_{t-1} lives transposed in shared memory ([k][window], broadcast
// float4 reads); W_hh^T (gate-interleaved, [k][4H]) streams through L1/L2 as coalesced float4;
// the cell state stays in registers for the whole sequence.
// ---------------------------------------------------------------------------------------------
constexpr int REC_THREADS = 256;
constexpr int REC_WPT = 16;  // windows per thread (inference tiles); small training batches use 8 to fill more SMs

// PF = W_hh rows (float4 each) in flight per thread.  Large batches run two CTAs per SM and are FMA-bound: PF = 8 (128-register
// budget).  Small (training) batches cannot fill the machine: there the L2 round trips of the weight stream are the step time,
// so the <.., 8, 32> variant runs one CTA per SM with 32 loads in flight (4 round trips per step instead of 16).
//
// RES_K > 0 (training batches, H = 128): the first RES_K rows of W_hh^T (96 of 128 = 192 KB) are copied into shared memory
// once and stay there for the whole sequence; only the remaining rows stream from L2, prefetched into registers at the top
// of the step.  With 512 windows the kernel was bound by exactly that stream (128 CTAs x 256 KB per step = 33 MB through
// L2 per step, ~10 us); the resident copy cuts it to a quarter.
template <int H, int REC_WPT = 16, int PF = 8, int RES_K = 0>
__global__ void __launch_bounds__(REC_THREADS, PF > 8 ? 1 : 2)
lstm_rec_f32(const float* __restrict__ G,      // [T][Bc][ND][H][4]
             const float* __restrict__ whh_f,  // [H][H][4] forward direction
             const float* __restrict__ whh_r,  // reverse direction
             float* __restrict__ out,          // [T][Bc][ND*H]
             float* __restrict__ gates_save,   // optional [T][Bc][ND][H][4] post-activation (train)
             float* __restrict__ c_save,       // optional [T][Bc][ND][H] (train)
             int Bc, int T, int ND) {
  constexpr int GROUPS = REC_THREADS / H;
  constexpr int MT = GROUPS * REC_WPT;
  constexpr int HS = MT + 4;  // padded row: conflict-free float4 stores, 16 B aligned
  __shared__ __align__(16) float hs[2][H][HS];
  const int tid = threadIdx.x;
  const int j = tid % H, grp = tid / H;
  const int dir = blockIdx.y;
  const int b_base = blockIdx.x * MT + grp * REC_WPT;
  const float4* __restrict__ W = reinterpret_cast<const float4*>(dir ? whh_r : whh_f);

  extern __shared__ __align__(16) float4 rec_wres[];  // [RES_K][H] (RES_K > 0 only)
  if (RES_K > 0)
    for (int i = tid; i < RES_K * H; i += REC_THREADS) rec_wres[i] = __ldg(W + i);
  for (int i = tid; i < 2 * H * HS; i += REC_THREADS) (&hs[0][0][0])[i] = 0.f;
  float c[REC_WPT];
#pragma unroll
  for (int w = 0; w < REC_WPT; ++w) c[w] = 0.f;
  __syncthreads();

  int cur = 0;
  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    float4 acc[REC_WPT];
    const float4* Gt = reinterpret_cast<const float4*>(G) + (((long long)t * Bc) * ND + dir) * H + j;
    // latency-bound variant (RES_K > 0): the projected inputs are added AFTER the recurrent product, so their loads are in
    // flight during the whole k loop instead of heading the dependent FMA chains (~1 us of exposed latency per step)
    constexpr bool LATE_G = RES_K > 0;
    float4 gin[LATE_G ? REC_WPT : 1];
#pragma unroll
    for (int w = 0; w < REC_WPT; ++w) {
      const int b = b_base + w;
      acc[w] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (LATE_G) gin[w] = (b < Bc) ? __ldg(Gt + (long long)b * ND * H) : make_float4(0.f, 0.f, 0.f, 0.f);
      else if (b < Bc) asm volatile("prefetch.global.L1 [%0];" ::"l"(Gt + (long long)b * ND * H));  // no registers to park it in: the
                                                                                             // load after the k loop then hits L1
    }
    const float* hcur = &hs[cur][0][grp * REC_WPT];
    // W_hh^T streams from L2 (it does not fit beside h in one SM's smem in fp32): 8 independent 16-byte loads are issued
    // per thread before they are consumed, otherwise each k iteration exposes a full L2 round trip (the first version,
    // unrolled by 2, spent ~36 k cycles per step on exactly that).
    auto fma_row = [&](int k, const float4 w4) {
      const float4* hp = reinterpret_cast<const float4*>(hcur + k * HS);
#pragma unroll
      for (int q = 0; q < REC_WPT / 4; ++q) {
        const float4 h4 = hp[q];
        const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float4& a = acc[q * 4 + e];
          a.x = fmaf(hv[e], w4.x, a.x); a.y = fmaf(hv[e], w4.y, a.y);
          a.z = fmaf(hv[e], w4.z, a.z); a.w = fmaf(hv[e], w4.w, a.w);
        }
      }
    };
    if (RES_K > 0) {
      constexpr int TAIL = H - RES_K;
      float4 wt[TAIL > 0 ? TAIL : 1];
#pragma unroll
      for (int kk = 0; kk < TAIL; ++kk) wt[kk] = __ldg(W + (long long)(RES_K + kk) * H + j);  // in flight during the resident part
#pragma unroll 8
      for (int k = 0; k < RES_K; ++k) fma_row(k, rec_wres[k * H + j]);
#pragma unroll
      for (int kk = 0; kk < TAIL; ++kk) fma_row(RES_K + kk, wt[kk]);
    } else {
      for (int k0 = 0; k0 < H; k0 += PF) {
        float4 w8[PF];
#pragma unroll
        for (int kk = 0; kk < PF; ++kk) w8[kk] = __ldg(W + (long long)(k0 + kk) * H + j);
#pragma unroll
        for (int kk = 0; kk < PF; ++kk) fma_row(k0 + kk, w8[kk]);
      }
    }
    // every variant adds the projected input AFTER the recurrent sum (same k order everywhere), so a window's result does not depend on
    // which variant -- i.e. which batch size -- computed it; the throughput variants load it here (the other CTA of the SM covers the latency)
#pragma unroll
    for (int w = 0; w < REC_WPT; ++w) {
      const int b = b_base + w;
      const float4 gv = LATE_G ? gin[w] : ((b < Bc) ? __ldg(Gt + (long long)b * ND * H) : make_float4(0.f, 0.f, 0.f, 0.f));
      acc[w].x += gv.x; acc[w].y += gv.y; acc[w].z += gv.z; acc[w].w += gv.w;
    }
    float4* hnext = reinterpret_cast<float4*>(&hs[cur ^ 1][j][grp * REC_WPT]);
    float hq[4];
#pragma unroll
    for (int w = 0; w < REC_WPT; ++w) {
      const float ig = rec_sigmoid(acc[w].x), fg = rec_sigmoid(acc[w].y);
      const float gg = rec_tanh(acc[w].z), og = rec_sigmoid(acc[w].w);
      c[w] = fmaf(fg, c[w], ig * gg);
      const float hv = og * rec_tanh(c[w]);
      hq[w & 3] = hv;
      if ((w & 3) == 3) hnext[w >> 2] = make_float4(hq[0], hq[1], hq[2], hq[3]);
      const int b = b_base + w;
      if (b < Bc) {
        const long long row = (long long)t * Bc + b;
        out[row * (ND * H) + dir * H + j] = hv;
        if (gates_save) {
          reinterpret_cast<float4*>(gates_save)[(row * ND + dir) * H + j] = make_float4(ig, fg, gg, og);
          c_save[(row * ND + dir) * H + j] = c[w];
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
}

int launch_proj_gemm_f32(const float* A, const float* Bt, const float* bias, float* C, int M, int N, int K, cudaStream_t st,
                         int accumulate) {
  BCI_REQUIRE(N % GN == 0 && K % GK == 0, BCI_EINVAL, "proj_gemm_f32: N %% 128 and K %% 16 required (N=%d K=%d)", N, K);
  dim3 gg(N / GN, ceil_div(M, GM));
  proj_gemm_f32<<<gg, GEMM_THREADS, 0, st>>>(A, Bt, bias, C, M, N, K, accumulate);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// Windows per thread of the small-batch recurrences (shared with BPTT): the smallest tile that still leaves SMs idle wins.
// BCI_REC_WPT=4|8|16 overrides it (experiment, 512-window training step: H = 128 16.8 / 32.9 ms with 4 / 8 windows per thread;
// H = 256 54.7 / 60.8 / 102 ms with 4 / 8 / 16 -- even where 4-window tiles need 1.7 waves and stream 1 MB of W_hh per CTA and
// step, the smaller tile is faster: the step is latency-bound, not L2-bandwidth-bound).
void small_batch_policy(int H, int ND, int Bc, int groups, bool& small, bool& tiny) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BCI_REC_WPT");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 4) { small = true; tiny = true; return; }
  if (forced == 8) { small = true; tiny = false; return; }
  if (forced == 16) { small = false; tiny = false; return; }
  (void)H; (void)ND; (void)Bc; (void)groups;
}

int launch_rec_f32(int H, int ND, const float* G, const float* whh_f, const float* whh_r, float* out, float* gates, float* csave,
                   int Bc, int T, cudaStream_t st) {
  // tiles of 16 windows per thread unless that leaves most SMs idle (training batches): then 8
  const int groups = REC_THREADS / H;
  bool small = ND * ceil_div(Bc, groups * 16) < sm_count();
  bool tiny = ND * ceil_div(Bc, groups * 8) < sm_count();  // even 8-window groups leave SMs idle (training: 512 windows)
  small_batch_policy(H, ND, Bc, groups, small, tiny);
  if (H == 128) {
    if (tiny) {
      constexpr int RES = 96;
      constexpr size_t res_bytes = (size_t)RES * 128 * sizeof(float4);
      static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
      if (!attr) {
        BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_f32<128, 4, 32, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_bytes));
        attr = true;
      }
      lstm_rec_f32<128, 4, 32, RES><<<dim3(ceil_div(Bc, 2 * 4), ND), REC_THREADS, res_bytes, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
    }
    else if (small) lstm_rec_f32<128, 8, 32><<<dim3(ceil_div(Bc, 2 * 8), ND), REC_THREADS, 0, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
    else lstm_rec_f32<128, 16><<<dim3(ceil_div(Bc, 2 * 16), ND), REC_THREADS, 0, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
  } else {
    if (tiny) lstm_rec_f32<256, 4, 32><<<dim3(ceil_div(Bc, 4), ND), REC_THREADS, 0, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
    else if (small) lstm_rec_f32<256, 8, 32><<<dim3(ceil_div(Bc, 8), ND), REC_THREADS, 0, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
    else lstm_rec_f32<256, 16><<<dim3(ceil_div(Bc, 16), ND), REC_THREADS, 0, st>>>(G, whh_f, whh_r, out, gates, csave, Bc, T, ND);
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---------------------------------------------------------------------------------------------
// weight packing (runs at bci_lstm_load_weights)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // dst[c][r] = src[r][c]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
  dst[(long long)c * rows + r] = src[i];
}

// src (4H, K) gate-major rows (i,f,g,o blocks) -> dst [K][ld] at column col0 + unit*4 + gate
__global__ void pack_gates_t_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int K, int ld, int col0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int gate = row / H, unit = row - gate * H;
  dst[(long long)k * ld + col0 + unit * 4 + gate] = src[i];
}

// W_hh (4H, H) gate-major rows -> dst [unit][j][gate]: one float4 = the four gate rows of `unit` at input column j (BPTT reads
// dh_{t-1}[j] = sum_n dG[n] W_hh[n][j] with coalesced 16-byte loads over j)
__global__ void pack_whh_b4_kernel(const float* __restrict__ src, float* __restrict__ dst, int H) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * H) return;
  const int row = (int)(i / H), j = (int)(i - (long long)row * H);
  const int gate = row / H, unit = row - gate * H;
  dst[((long long)unit * H + j) * 4 + gate] = src[i];
}

// src (4H, K) gate-major rows -> dst [row0 + unit*4 + gate][K]
__global__ void pack_gates_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int K, int row0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)4 * H * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int gate = row / H, unit = row - gate * H;
  dst[(long long)(row0 + unit * 4 + gate) * K + k] = src[i];
}

__global__ void pack_bias_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ dst, int H, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H) return;
  const int gate = i / H, unit = i - gate * H;
  dst[col0 + unit * 4 + gate] = bih[i] + bhh[i];
}

__global__ void copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

static inline unsigned nblk(long long n) { return (unsigned)ceil_div64(n, 256); }

// input_proj.0.weight (H, C) -> [part][H][64] fp16 pair x 16, K zero-padded
__global__ void pack_w0_f16_kernel(const float* __restrict__ w0, __half* __restrict__ dst, int H, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * 64) return;
  const int j = i >> 6, k = i & 63;
  const float v = k < C ? w0[j * C + k] * F16X3_WSCALE : 0.f;
  const __half hi = __float2half_rn(v);
  dst[i] = hi;
  dst[H * 64 + i] = __float2half_rn(v - __half2float(hi));
}

// the fp16-split operand copies of the large-batch inference path, from the fp32 layouts packed by lstm_pack_f32 (same stream)
static int pack_f16_operands(bci_lstm_s* h, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const bci_lstm_weights& w = h->raw;
  PackedF32& p = h->f32;
  const int H = c.hidden_size, C = c.input_size, ND = num_dirs(c), D = ND * H, AH = D / 2;
  int rc = 0;
  pack_w0_f16_kernel<<<ceil_div(H * 64, 256), 256, 0, st>>>(w.input_proj_w, p.w0_16, H, C);
  BCI_LAUNCH_OK();
  for (int l = 0; l < c.num_layers && !rc; ++l) {
    const int K = layer_in_width(c, l);
    for (int d = 0; d < ND && !rc; ++d) rc = pack_whh_f16x3(w.w_hh[l][d], p.whh16[l] + (size_t)d * 2 * 4 * H * H, H, st);
    if (!rc) rc = split_f16(p.wih_b[l], p.wih16[l], p.wih16[l] + (size_t)ND * 4 * H * K, (long long)ND * 4 * H * K, F16X3_WSCALE, st);
  }
  if (!rc && c.use_attention) rc = split_f16(p.aw1, p.aw1_16, p.aw1_16 + (size_t)AH * D, (long long)AH * D, F16X3_WSCALE, st);
  if (!rc) h->f16_stale = false;
  return rc;
}

int lstm_pack_f32(bci_lstm_s* h, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const bci_lstm_weights& w = h->raw;
  PackedF32& p = h->f32;
  const int H = c.hidden_size, C = c.input_size, ND = num_dirs(c), D = ND * H, AH = D / 2;
  transpose_kernel<<<nblk((long long)H * C), 256, 0, st>>>(w.input_proj_w, p.w0t, H, C);
  h->f16_stale = true;
  h->sw_stale = true;
  copy_kernel<<<nblk(H), 256, 0, st>>>(w.input_proj_b, p.b0, H);
  if (c.use_layer_norm) {
    copy_kernel<<<nblk(H), 256, 0, st>>>(w.input_ln_w, p.ln0w, H);
    copy_kernel<<<nblk(H), 256, 0, st>>>(w.input_ln_b, p.ln0b, H);
  }
  for (int l = 0; l < c.num_layers; ++l) {
    const int K = layer_in_width(c, l);
    for (int d = 0; d < ND; ++d) {
      pack_gates_t_kernel<<<nblk((long long)4 * H * K), 256, 0, st>>>(w.w_ih[l][d], p.wih_t[l], H, K, ND * 4 * H, d * 4 * H);
      pack_gates_t_kernel<<<nblk((long long)4 * H * H), 256, 0, st>>>(w.w_hh[l][d], p.whh_t[l][d], H, H, 4 * H, 0);
      pack_bias_kernel<<<nblk(4 * H), 256, 0, st>>>(w.b_ih[l][d], w.b_hh[l][d], p.bias[l], H, d * 4 * H);
      pack_gates_rows_kernel<<<nblk((long long)4 * H * K), 256, 0, st>>>(w.w_ih[l][d], p.wih_b[l], H, K, d * 4 * H);
      pack_whh_b4_kernel<<<nblk((long long)4 * H * H), 256, 0, st>>>(w.w_hh[l][d], p.whh_b[l][d], H);
    }
    int rc = split_tf32(p.wih_b[l], nullptr, p.wih_b_lo[l], (long long)ND * 4 * H * K, st);
    if (!rc) rc = split_tf32(p.wih_t[l], nullptr, p.wih_t_lo[l], (long long)ND * 4 * H * K, st);
    if (rc) return rc;
  }
  if (c.use_layer_norm) {
    copy_kernel<<<nblk(D), 256, 0, st>>>(w.ln_w, p.lnw, D);
    copy_kernel<<<nblk(D), 256, 0, st>>>(w.ln_b, p.lnb, D);
  }
  if (c.use_attention) {
    transpose_kernel<<<nblk((long long)AH * D), 256, 0, st>>>(w.attn_w1, p.aw1t, AH, D);
    copy_kernel<<<nblk(AH * D), 256, 0, st>>>(w.attn_w1, p.aw1, AH * D);
    int rc = split_tf32(p.aw1, nullptr, p.aw1_lo, (long long)AH * D, st);
    if (!rc) rc = split_tf32(p.aw1t, nullptr, p.aw1t_lo, (long long)AH * D, st);
    if (rc) return rc;
    copy_kernel<<<nblk(AH), 256, 0, st>>>(w.attn_b1, p.ab1, AH);
    copy_kernel<<<nblk(AH), 256, 0, st>>>(w.attn_w2, p.aw2, AH);
    copy_kernel<<<1, 256, 0, st>>>(w.attn_b2, p.ab2, 1);
  }
  transpose_kernel<<<nblk((long long)H * D), 256, 0, st>>>(w.cls_w0, p.c0t, H, D);
  copy_kernel<<<nblk(H), 256, 0, st>>>(w.cls_b0, p.cb0, H);
  transpose_kernel<<<nblk((long long)(H / 2) * H), 256, 0, st>>>(w.cls_w3, p.c3t, H / 2, H);
  copy_kernel<<<nblk(H / 2), 256, 0, st>>>(w.cls_b3, p.cb3, H / 2);
  copy_kernel<<<nblk(c.num_classes * (H / 2)), 256, 0, st>>>(w.cls_w6, p.c6, c.num_classes * (H / 2));
  copy_kernel<<<1, 256, 0, st>>>(w.cls_b6, p.cb6, c.num_classes);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

size_t lstm_store_bytes_f32(const bci_lstm_config& c) {
  const size_t H = c.hidden_size, C = c.input_size, D = 2 * H;
  size_t n = C * H + 3 * H;
  for (int l = 0; l < c.num_layers; ++l) n += 4 * ((size_t)layer_in_width(c, l) * 8 * H) + 2 * (2 * H * 4 * H) + 8 * H + 2 * 4 * H * H + (size_t)layer_in_width(c, l) * 8 * H + 5 * (4 * H * H);
  n += 2 * D + 5 * D * H + 64 * H + H + H + 4 + D * H + H + H * (H / 2) + H / 2 + (size_t)c.num_classes * (H / 2) + c.num_classes + 64;
  return align_up(n * sizeof(float) + 256 * 80, 256);
}

void lstm_carve_f32(bci_lstm_s* h, char* base) {
  const bci_lstm_config& c = h->cfg;
  const size_t H = c.hidden_size, C = c.input_size, D = 2 * H;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = reinterpret_cast<float*>(base + off); off += align_up(n * sizeof(float), 256); return p; };
  PackedF32& p = h->f32;
  p.w0t = take(C * H); p.b0 = take(H); p.ln0w = take(H); p.ln0b = take(H);
  for (int l = 0; l < c.num_layers; ++l) {
    p.wih_t[l] = take((size_t)layer_in_width(c, l) * 8 * H);
    p.bias[l] = take(8 * H);
    p.whh_t[l][0] = take(H * 4 * H);
    p.whh_t[l][1] = take(H * 4 * H);
    p.wih_b[l] = take((size_t)layer_in_width(c, l) * 8 * H);
    p.whh_b[l][0] = take(H * 4 * H);
    p.whh_b[l][1] = take(H * 4 * H);
    p.wih_b_lo[l] = take((size_t)layer_in_width(c, l) * 8 * H);
    p.wih_t_lo[l] = take((size_t)layer_in_width(c, l) * 8 * H);
    p.whh16[l] = reinterpret_cast<__half*>(take(2 * 4 * H * H));   // 2 directions x 2 parts x 4H x H halves
    p.wih16[l] = reinterpret_cast<__half*>(take((size_t)layer_in_width(c, l) * 8 * H));   // 2 parts x 8H x K_l halves
    p.whh_sw_f[l] = reinterpret_cast<__half*>(take(2 * 4 * H * H));      // 2 directions x 2 parts x 4H x H halves
    p.whh_sw_b[l] = reinterpret_cast<__nv_bfloat16*>(take(4 * H * H));   // 2 directions x H x 4H bf16
    p.whh_sw_b16[l] = reinterpret_cast<__half*>(take(2 * 4 * H * H));    // 2 directions x 2 parts x H x 4H halves
  }
  p.lnw = take(D); p.lnb = take(D); p.aw1t = take(D * H); p.ab1 = take(H); p.aw2 = take(H); p.ab2 = take(4);
  p.aw1 = take(D * H); p.aw1_lo = take(D * H); p.aw1t_lo = take(D * H);
  p.aw1_16 = reinterpret_cast<__half*>(take(D * H));   // 2 parts x (D/2) x D halves
  p.w0_16 = reinterpret_cast<__half*>(take(64 * H));   // 2 parts x H x 64 halves
  p.c0t = take(D * H); p.cb0 = take(H); p.c3t = take(H * (H / 2)); p.cb3 = take(H / 2);
  p.c6 = take((size_t)c.num_classes * (H / 2)); p.cb6 = take(c.num_classes);
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
static size_t chunk_bytes_f32(const bci_lstm_config& c, int Bc, int T) {
  const size_t H = c.hidden_size, D = feat_width(c), rows = (size_t)Bc * T;
  return align_up(rows * H * 4, 256) + align_up(rows * 4 * D * 4, 256) + 3 * align_up(rows * D * 4, 256) +
         align_up((size_t)Bc * T * 4, 256);
}

size_t lstm_chunk_bytes_f32(const bci_lstm_config& c, int Bc, int T) { return chunk_bytes_f32(c, Bc, T); }

size_t lstm_workspace_fp32(const bci_lstm_config& c, int batch, int T) {
  const int Bc = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  return chunk_bytes_f32(c, Bc > 0 ? Bc : 1, T);
}

template <int H, int ND>
static int forward_chunk_f32(bci_lstm_s* h, const InputView& x, int Bc, int T, float* logits, float* probs, float* attn,
                             char* ws, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const size_t rows = (size_t)Bc * T;
  constexpr int D = ND * H;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = ws + off; off += align_up(bytes, 256); return p; };
  float* z = reinterpret_cast<float*>(take(rows * H * 4));
  float* g = reinterpret_cast<float*>(take(rows * 4 * D * 4));
  float* o0 = reinterpret_cast<float*>(take(rows * D * 4));
  float* o1 = reinterpret_cast<float*>(take(rows * D * 4));
  float* scores = reinterpret_cast<float*>(take(rows * 4));
  float* in_lo = reinterpret_cast<float*>(take(rows * D * 4));  // tf32 remainder of the layer input (split-precision GEMM)
  h->prof.mark(-1, st);
  int rc = 0;
  const float* in = z;
  float* outs[2] = {o0, o1};
  // Large batches (H = 128): the whole LSTM stack on the tensor cores in split FP16 precision -- projections as three fp16 MMA
  // chains per product (gemm_f16x3_nt: twice the rate of the 3 x TF32 form), recurrences on CTA pairs (lstm_fp32_tc.cu).  The layer
  // input travels as an fp16 (hi, lo) pair: z is split once, and every recurrence writes h_t directly in that form for the next
  // layer's projection (the pair it computes for its own MMA operand), so there is no split pass between layers; only the last
  // layer writes fp32 for the pooling kernel.  The pair buffers reuse the space of the tf32 remainder array.
  __half* in_hi16 = reinterpret_cast<__half*>(in_lo);
  __half* in_lo16 = in_hi16 + rows * D;
  const bool fast = h->infer_fast;   // small batch of the bf16 engine: reduced-precision forms of the same kernels
  const bool tc = !fast && H == 128 && tc_rec_ok(H, ND, Bc, g, 4 * D, o0, D) && c.input_size <= 64 &&
                  f16x3_nt_ok(in_hi16, H, h->f32.wih16[0], H, g, 4 * D, (int)rows, 4 * D, H);
  if (tc && h->f16_stale && (rc = pack_f16_operands(h, st))) return rc;
  if (tc) {
    // K1 on the tensor cores too: x -> fp16 pair rows (K padded to 64, in the z buffer) -> split-fp16 GEMM (+ b0, into the G
    // buffer) -> LayerNorm + erf-GELU row kernel that writes z directly as the pair the first projection reads
    __half* x_hi = reinterpret_cast<__half*>(z);
    __half* x_lo = x_hi + rows * 64;
    x_pair_rows_kernel<<<(unsigned)ceil_div64((long long)rows * 32, 256), 256, 0, st>>>(x, Bc, T, c.input_size, x_hi, x_lo);
    BCI_LAUNCH_OK();
    if ((rc = gemm_f16x3_nt(x_hi, x_lo, 64, h->f32.w0_16, h->f32.w0_16 + (size_t)H * 64, 64, h->f32.b0, g, H, (int)rows, H, 64,
                            1.0f / F16X3_WSCALE, st)))
      return rc;
    ln_gelu_pair_rows128_kernel<<<(unsigned)(rows / 8 < 4096 ? (rows + 7) / 8 : 4096), 256, 0, st>>>(
        g, (long long)rows, h->f32.ln0w, h->f32.ln0b, c.use_layer_norm, in_hi16, in_lo16);
    BCI_LAUNCH_OK();
  } else {
    if ((rc = launch_input_proj<H, float>(h, x, Bc, T, z, st, fast))) return rc;
  }
  h->prof.mark(0, st);
  for (int l = 0; l < c.num_layers; ++l) {
    const int K = layer_in_width(c, l);
    const int M = (int)rows, N = 4 * D;
    if (tc) {
      const bool last = l == c.num_layers - 1;
      const __half* w16 = h->f32.wih16[l];
      if ((rc = gemm_f16x3_nt(in_hi16, in_lo16, K, w16, w16 + (size_t)N * K, K, h->f32.bias[l], g, N, M, N, K, 1.0f / F16X3_WSCALE, st)))
        return rc;
      h->prof.mark(1, st);
      float* o = outs[l & 1];
      if ((rc = launch_rec_f16x3(ND, g, N, h->f32.whh16[l], last ? o : nullptr, last ? nullptr : in_hi16, last ? nullptr : in_lo16,
                                 nullptr, nullptr, D, Bc, T, st)))
        return rc;
      h->prof.mark(2, st);
      in = o;
      continue;
    }
    if (tf32x3_nt_ok(in, K, h->f32.wih_b[l], K, g, N, M, N, K)) {
      // G = in . W_ih^T on the tensor cores in split precision (3 x TF32, fp32-grade)
      if (!fast && (rc = split_tf32(in, nullptr, in_lo, (long long)M * K, st))) return rc;
      if ((rc = gemm_tf32x3_nt(in, fast ? nullptr : in_lo, K, h->f32.wih_b[l], fast ? nullptr : h->f32.wih_b_lo[l], K, h->f32.bias[l], g, N,
                               M, N, K, 0, st)))
        return rc;
    } else {
      dim3 gg(N / GN, ceil_div(M, GM));
      proj_gemm_f32<<<gg, GEMM_THREADS, 0, st>>>(in, h->f32.wih_t[l], h->f32.bias[l], g, M, N, K, 0);
      BCI_LAUNCH_OK();
    }
    h->prof.mark(1, st);
    float* o = outs[l & 1];
    // the tile size follows the batch (launch_rec_f32): small batches -- single windows of predict_trajectory, the reference's 256 /
    // 512-window passes -- use the latency-oriented variants with W_hh resident in shared memory instead of leaving most SMs idle
    // H = 128: the swapped tensor-core recurrence in its split (fp32-grade) form -- W_hh resident in tensor memory, 8 windows per
    // CTA (lstm_rec_swap.cu; 1.96 us per step against 4.9 on the CUDA cores at 512 windows)
    if (H == 128 && swap_rec_enabled() && rec_swap_ok(H, g, N)) {
      if ((rc = pack_swap_operands(h, st))) return rc;
      if ((rc = launch_rec_swap_fwd(ND, g, N, h->f32.whh_sw_f[l], o, nullptr, nullptr, D, Bc, T, !fast, st))) return rc;
    } else if ((rc = launch_rec_f32(H, ND, g, h->f32.whh_t[l][0], h->f32.whh_t[l][1], o, nullptr, nullptr, Bc, T, st))) {
      return rc;
    }
    h->prof.mark(2, st);
    in = o;
  }
  if (tc && ND == 2 && c.use_layer_norm && c.use_attention && f16x3_nt_ok(in_hi16, D, h->f32.aw1_16, D, g, D / 2, (int)rows, D / 2, D)) {
    // attention scores on the tensor cores as well: y = LN(out) as an fp16 pair (the pair buffers are free again), then
    // pre = y . W1^T + b1 by the split-fp16 GEMM into the (now dead) G buffer; the pooling kernel reads pre instead of running
    // its CUDA-core score product (7.1 -> ~3 ms per 9472 windows)
    ln_pair_rows256_kernel<<<(unsigned)(rows / 8 < 4096 ? (rows + 7) / 8 : 4096), 256, 0, st>>>(in, (long long)rows, h->f32.lnw, h->f32.lnb,
                                                                                              in_hi16, in_lo16);
    BCI_LAUNCH_OK();
    if ((rc = gemm_f16x3_nt(in_hi16, in_lo16, D, h->f32.aw1_16, h->f32.aw1_16 + (size_t)(D / 2) * D, D, h->f32.ab1, g, D / 2, (int)rows,
                            D / 2, D, 1.0f / F16X3_WSCALE, st)))
      return rc;
    rc = launch_pool_head<H, ND, float>(h, in, Bc, T, logits, probs, attn, scores, st, g);
  } else {
    rc = launch_pool_head<H, ND, float>(h, in, Bc, T, logits, probs, attn, scores, st);
  }
  h->prof.mark(3, st);
  return rc;
}

int lstm_forward_fp32(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const int chunk = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  BCI_REQUIRE(ws_bytes >= chunk_bytes_f32(c, chunk, T), BCI_ENOMEM, "bci_lstm_forward: workspace %zu < %zu bytes", ws_bytes,
              chunk_bytes_f32(c, chunk, T));
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int Bc = (batch - b0) < chunk ? (batch - b0) : chunk;
    const InputView xb = chunk_view(x, b0);
    float* lg = logits + (size_t)b0 * c.num_classes;
    float* pr = probs ? probs + (size_t)b0 * c.num_classes : nullptr;
    float* at = attn ? attn + (size_t)b0 * T : nullptr;
    int rc;
    if (c.bidirectional)
      rc = c.hidden_size == 128 ? forward_chunk_f32<128, 2>(h, xb, Bc, T, lg, pr, at, (char*)ws, st)
                                : forward_chunk_f32<256, 2>(h, xb, Bc, T, lg, pr, at, (char*)ws, st);
    else
      rc = c.hidden_size == 128 ? forward_chunk_f32<128, 1>(h, xb, Bc, T, lg, pr, at, (char*)ws, st)
                                : forward_chunk_f32<256, 1>(h, xb, Bc, T, lg, pr, at, (char*)ws, st);
    if (rc) return rc;
  }
  return BCI_OK;
}

}  // namespace bci
