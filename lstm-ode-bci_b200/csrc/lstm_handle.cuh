// Handle + packed-weight layout shared by the LSTM translation units.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace bci {

// Packed fp32 weights (all device pointers into one allocation owned by the handle).
// Gate-interleaved column order everywhere: column n = dir*4H + unit*4 + gate (gate = i,f,g,o)
// so that one thread owning hidden unit `unit` reads its four gates as one float4.
struct PackedF32 {
  float* w0t;    // [C][H]      input_proj.0.weight^T
  float* b0;     // [H]
  float* ln0w;   // [H]
  float* ln0b;   // [H]
  float* wih_t[BCI_MAX_LAYERS];     // [K_l][ND*4H] all directions, gate-interleaved (ND = 1 or 2 directions)
  float* bias[BCI_MAX_LAYERS];      // [ND*4H]     b_ih + b_hh, same order
  float* whh_t[BCI_MAX_LAYERS][2];  // [H][4H]     per direction, gate-interleaved
  // row-major copies with gate-interleaved ROWS (n = dir*4H + unit*4 + gate), used by the backward pass:
  float* wih_b[BCI_MAX_LAYERS];     // [ND*4H][K_l] din = dG . wih_b
  float* whh_b[BCI_MAX_LAYERS][2];  // [H unit][H j][4 gates]  dh_{t-1}[j] = sum_(unit,gate) dG_t . whh_b
  // split-precision fp16 operand of the tensor-core recurrence (lstm_fp32_tc.cu, H = 128): [ND][part hi/lo][n' = unit*4 + gate][k],
  // values scaled by 16
  __half* whh16[BCI_MAX_LAYERS];
  // operands of the swapped tensor-core recurrences of the mixed-precision training step (lstm_rec_swap.cu, H = 128):
  // whh_sw_f [ND][part hi/lo][4H][H] fp16 of 16 w (PyTorch row order); the transpose [H j][4H k = gate*H + unit] as bf16
  // (whh_sw_b [ND][H][4H]) and as an fp16 pair of 16 w (whh_sw_b16 [ND][part][H][4H])
  __half* whh_sw_f[BCI_MAX_LAYERS];
  __nv_bfloat16* whh_sw_b[BCI_MAX_LAYERS];
  __half* whh_sw_b16[BCI_MAX_LAYERS];
  // ... and of the projection GEMM in its fp16-split form (gemm_tf32x3.cu, F16): [part hi/lo][ND*4H gate-interleaved rows][K_l], x 16
  __half* wih16[BCI_MAX_LAYERS];
  __half* w0_16;   // input_proj.0.weight (H, C) zero-padded to K = 64, fp16 (hi, lo) pair x 16
  __half* aw1_16;  // attention.0.weight (D/2, D) as an fp16 (hi, lo) pair x 16: B operand of the score GEMM (fp32 large-batch path)
  // tf32 remainders (x - tf32(x)) of wih_b / wih_t: second operand of the split-precision tcgen05 GEMMs (gemm_tf32x3.cu)
  float* wih_b_lo[BCI_MAX_LAYERS];
  float* wih_t_lo[BCI_MAX_LAYERS];
  float* aw1;      // [D/2][D]    attention.0.weight (own copy: the split pair below must describe the same bits)
  float* aw1_lo;   // remainder of aw1
  float* aw1t_lo;  // remainder of aw1t
  float* lnw;    // [D]         D = ND*H
  float* lnb;    // [D]
  float* aw1t;   // [D][D/2]    attention.0.weight^T
  float* ab1;    // [D/2]
  float* aw2;    // [D/2]
  float* ab2;    // [1]
  float* c0t;    // [D][H]      classifier.0.weight^T
  float* cb0;    // [H]
  float* c3t;    // [H][H/2]    classifier.3.weight^T
  float* cb3;    // [H/2]
  float* c6;     // [classes][H/2]
  float* cb6;    // [classes]
};

// bf16 operand copies for the tcgen05 path (K-major rows, PyTorch (out,in) orientation kept):
//   wih_bf[l]    [8H][K_l]  row n' = dir*4H + perm(unit,gate)   (B operand of the projection GEMM)
//   whh_bf[l][d] [4H][H]    row n' = perm(unit,gate)            (B operand of the recurrence MMA)
//   bias_p[l]    [8H] fp32 in the same permuted order
// perm(unit,gate) = (unit/8)*32 + gate*8 + unit%8: a 32-column TMEM slab holds i,f,g,o of 8 units.
struct PackedBF16 {
  __nv_bfloat16* wih_bf[BCI_MAX_LAYERS];
  __nv_bfloat16* whh_bf[BCI_MAX_LAYERS][2];
  float* bias_p[BCI_MAX_LAYERS];
  // fused cluster kernel (lstm_bf16_fused.cu): W_ih rows and biases in the SAME perm_T order as whh_bf
  __nv_bfloat16* wih_t_bf[BCI_MAX_LAYERS];  // [2][4H][K_l]
  float* bias_t[BCI_MAX_LAYERS];            // [2][4H]
  // H = 256 (lstm_bf16_h256.cu): rows / columns in perm_256 order
  __nv_bfloat16* wih256[BCI_MAX_LAYERS];    // [2*1024][K_l]  B operand of the projection GEMM
  __nv_bfloat16* whh256[BCI_MAX_LAYERS];    // [2][1024][256] resident operand of the cluster recurrence
  float* bias256[BCI_MAX_LAYERS];           // [2*1024]
  float* zero_bias;                         // [256] zeros (bias operand of the H = 256 score GEMM)
  // attention scores on tensor cores with LayerNorm folded in (lstm_bf16_pool.cu):
  //   aw1_bf [H][2H] = bf16(W1[j][d] * ln_w[d]);  apar[j] = {s_j = sum_d aw1_bf[j][d], c_j = b1_j + sum_d ln_b[d] W1[j][d], w2_j, 0}
  __nv_bfloat16* aw1_bf;
  float4* apar;
  float pool_smax;  // HOST value: sum_j |attention.2.weight_j| >= |score| (bound used by the single-pass pooling kernel)
  float pool_par[3][128];  // HOST copies of apar's {s_j}, {c_j}, {w2_j}: passed to that kernel as a constant-bank argument
  // input projection on tensor cores (lstm_bf16_inproj.cu): w0_bf [H][64] = bf16(input_proj.0.weight), K zero-padded
  // from C to 64; par0[j] = {b0_j, ln_w_j, ln_b_j, 0}
  __nv_bfloat16* w0_bf;
  float4* par0;
};

}  // namespace bci

namespace bci {
// BCI_NVTX=1: an NVTX range per forward ("bci_lstm_forward") with a mark at the end of each phase, for `ncu --nvtx --nvtx-include`
// / timeline tools (tracing).  nvtx3 is header-only; without an attached tool the calls are no-ops.
inline void nvtx_phase(int ph) {
  if (!nvtx_on()) return;
  static const char* const names[4] = {"end input_proj", "end proj_gemm", "end recurrence", "end pool_head"};
  if (ph < 0) {
    nvtxRangePushA("bci_lstm_forward");
  } else {
    nvtxMarkA(names[ph & 3]);
    if (ph == 3) nvtxRangePop();
  }
}
struct Profiler {
  static constexpr int MAX_EV = 2048;
  bool enabled = false;
  int n = 0;
  cudaEvent_t ev[MAX_EV];
  int phase[MAX_EV];      // phase that ENDS at ev[i] (-1 for the opening event of a forward)
  bool created = false;
  // record the end of `ph` (or the start marker with ph = -1)
  void mark(int ph, cudaStream_t st) {
    nvtx_phase(ph);
    if (!enabled || n >= MAX_EV) return;
    if (!created) { for (int i = 0; i < MAX_EV; ++i) cudaEventCreate(&ev[i]); created = true; }
    cudaEventRecord(ev[n], st);
    phase[n++] = ph;
  }
};
}  // namespace bci

struct bci_lstm_s {
  bci_lstm_config cfg;
  int device;
  bool loaded;
  void* store;         // one cudaMalloc
  size_t store_bytes;
  bci::PackedF32 f32;
  bci::PackedBF16 bf16;
  // raw (unpacked) weight pointers of the last load_weights (caller-owned; used by backward)
  bci_lstm_weights raw;
  bci::Profiler prof;
  // side stream of the backward pass (weight-gradient GEMMs of layer l overlap the BPTT recurrence of layer l-1)
  cudaStream_t side;
  cudaEvent_t ev_dg, ev_side[2], ev_join;
  bool side_ready;
  // the fp16-split operand copies (whh16 / wih16 / aw1_16 / w0_16) serve the large-batch fp32 INFERENCE path only: they are packed on
  // its first use after a load, so a training loop that reloads the weights every step does not pay for them
  bool f16_stale;
  // training precision (bci_lstm_set_train_mode): BCI_TRAIN_FP32 (parity) or BCI_TRAIN_MIXED; the mixed step's 16-bit recurrent
  // operands are packed on its first forward after a load
  int train_mode;
  bool sw_stale;
  // set by the bf16 engine around a small-batch forward it hands to the fp32 path's small-batch code in its fast form (one fp16
  // product chain in the swapped recurrence, single-pass TF32 projections): see lstm_forward_bf16
  bool infer_fast;
  // the last train=1 forward (workspace + header): a backward on the same workspace needs no device->host read of the header
  void* last_train_ws;
  float last_dropout;
  uint64_t last_seed;
  int last_batch, last_T, last_mode;
};

namespace bci {

// Where the windows of a forward come from (bci_lstm_input of the C ABI + the index of the chunk's first window).  Window w starts at
// element  (w / wpr) * rstride + (w % wpr) * wstride  (wpr == 0: w * wstride) and is seq_len x C contiguous elements from there:
// packed (B,T,C) windows have wstride = T*C; overlapping windows of a (samples, C) recording have wstride = step*C (02:157-180).
struct InputView {
  const void* data;
  int dtype;              // BCI_IN_F32 | BCI_IN_BF16
  int wpr;                // windows per run (recording); 0 = one run
  long long wstride;      // elements
  long long rstride;      // elements
  long long first;        // global index of window 0 of this chunk
  __host__ __device__ long long elem_off(long long b) const {
    const long long w = first + b;
    if (wpr > 0) { const long long r = w / wpr; return r * rstride + (w - r * wpr) * wstride; }
    return w * wstride;
  }
  __host__ __device__ int esize() const { return dtype == BCI_IN_BF16 ? 2 : 4; }
};
inline InputView packed_view(const float* x, int T, int C) { return InputView{x, BCI_IN_F32, 0, (long long)T * C, 0, 0}; }
inline InputView chunk_view(InputView v, long long b0) { v.first += b0; return v; }
#ifdef __CUDACC__
__device__ __forceinline__ float view_load(const InputView& v, long long e) {
  return v.dtype == BCI_IN_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(v.data)[e])
                                : reinterpret_cast<const float*>(v.data)[e];
}
#endif

inline int num_dirs(const bci_lstm_config& c) { return c.bidirectional ? 2 : 1; }
inline int feat_width(const bci_lstm_config& c) { return num_dirs(c) * c.hidden_size; }          // D: LSTM output width
inline int attn_width(const bci_lstm_config& c) { return feat_width(c) / 2; }                    // attention hidden width
inline int layer_in_width(const bci_lstm_config& c, int l) { return l == 0 ? c.hidden_size : feat_width(c); }

// chunking policy: windows processed per internal pass (bounds the workspace)
int bf16_chunk_windows();  // lstm_bf16.cu: depends on the recurrence path in use (split K2+K3 or fused cluster kernel)
int h256_max_clusters();   // lstm_bf16_h256.cu: co-resident 4-CTA clusters of the H = 256 recurrence
int tc_max_clusters();     // lstm_fp32_tc.cu: co-resident CTA pairs of the fp32 tensor-core recurrence
inline int max_chunk(const bci_lstm_config& c, int train) {
  if (c.precision == BCI_PRECISION_BF16)
    return train ? 2048 : (c.hidden_size == 256 ? h256_max_clusters() * 2 * 128 : bf16_chunk_windows());
  const int base = train ? 512 : 2048;
  if (!train && c.hidden_size == 128 && tc_max_clusters() > 0) {
    // fp32 inference, H = 128: one full wave of the pair recurrence (lstm_fp32_tc.cu): clusters x 256 windows / 2 directions
    const int wave = tc_max_clusters() * 128;
    return wave > base ? wave : base;
  }
  return c.hidden_size > 128 ? base / 2 : base;
}

struct FwdWorkspace {
  // fp32 path
  float* z;       // [T][Bc][H]
  float* g;       // [T][Bc][8H]
  float* out[2];  // [T][Bc][2H] ping-pong
  size_t total;
};

// fp32 forward entry points (lstm_fp32.cu)
int lstm_forward_fp32(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn,
                      void* ws, size_t ws_bytes, cudaStream_t st);
size_t lstm_workspace_fp32(const bci_lstm_config& c, int batch, int T);
size_t lstm_chunk_bytes_f32(const bci_lstm_config& c, int Bc, int T);
int launch_proj_gemm_f32(const float* A, const float* Bt, const float* bias, float* C, int M, int N, int K, cudaStream_t st,
                         int accumulate = 0);
int launch_rec_f32(int H, int ND, const float* G, const float* whh_f, const float* whh_r, float* out, float* gates, float* csave,
                   int Bc, int T, cudaStream_t st);
void small_batch_policy(int H, int ND, int Bc, int groups, bool& small, bool& tiny);
// fp32-grade recurrence on the tensor cores (lstm_fp32_tc.cu)
int pack_whh_f16x3(const float* w_hh, __half* dst, int H, cudaStream_t st);
bool tc_rec_ok(int H, int ND, int Bc, const void* G, int ldg, const void* out, int D);
int tc_max_clusters();
int launch_rec_f16x3(int ND, const float* G, int ldg, const __half* whh16, float* out, __half* out_hi16, __half* out_lo16, float* gates,
                     float* csave, int D, int Bc, int T, cudaStream_t st);
// swapped (weights-as-A) tensor-core recurrences of the mixed-precision training step (lstm_rec_swap.cu)
// dropout between LSTM layers folded into the recurrence kernels (forward: dropped copy of h_t written next to h_t; BPTT: the same
// mask applied to the incoming gradient)
struct SwapDropout {
  float* outd;
  float* outd_lo;
  float p;
  uint64_t seed;
  uint32_t site;
};
int pack_whh_swap(const float* w_hh, __half* fwd, __nv_bfloat16* bwd, __half* bwd16, int H, cudaStream_t st);
bool rec_swap_ok(int H, const void* G, int ldg);
bool swap_rec_enabled();
// hidden_size 256, mixed precision: the same recurrences on CTA pairs (each CTA owns 128 units; h / dG halves exchanged through DSMEM)
bool rec_swap256_ok(int H, const void* G, int ldg);
int launch_rec_swap256_fwd(int ND, const float* G, int ldg, const __half* whh, float* out, float* gates, float* csave, int D, int Bc, int T,
                           cudaStream_t st, const SwapDropout* drop = nullptr, bool g_half = false);
int launch_bptt_swap256(int ND, const float* dout, const float* gates, const float* csave, const __nv_bfloat16* whhT, float* dG, float* dbias,
                        int ldg, int D, int Bc, int T, cudaStream_t st, const SwapDropout* drop = nullptr);
int pack_swap_operands(bci_lstm_s* h, cudaStream_t st);
int launch_rec_swap_fwd(int ND, const float* G, int ldg, const __half* whh, float* out, float* gates, float* csave, int D, int Bc, int T,
                        bool split, cudaStream_t st, const SwapDropout* drop = nullptr, bool g_half = false);
int launch_bptt_swap(int ND, const float* dout, const float* gates, const float* csave, const void* whhT, float* dG, float* dG_lo,
                     float* dbias, int ldg, int D, int Bc, int T, bool split, cudaStream_t st, const SwapDropout* drop = nullptr);
constexpr float F16X3_WSCALE = 16.0f;   // weights of the fp16-split paths are stored x 16 (keeps their lo parts out of fp16's subnormals)
int split_f16(const float* x, __half* hi, __half* lo, long long n, float scale, cudaStream_t st);
bool f16x3_nt_ok(const void* A_hi, int lda, const void* W_hi, int ldw, const void* C, int ldc, int M, int N, int K);
int gemm_f16x3_nt(const __half* A_hi, const __half* A_lo, int lda, const __half* W_hi, const __half* W_lo, int ldw, const float* bias,
                  float* C, int ldc, int M, int N, int K, float out_scale, cudaStream_t st);
// split-precision tcgen05 GEMMs of the fp32 path (gemm_tf32x3.cu)
bool tf32x3_enabled();
bool tf32x3_nt_ok(const void* A, int lda, const void* W, int ldw, const void* C, int ldc, int M, int N, int K);
bool tf32x3_tn_ok(const void* A, int lda, const void* B, int ldb, const void* C, int ldc, long long R, int P, int Q);
int split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st);
int gemm_tf32x3_nt(const float* A_hi, const float* A_lo, int lda, const float* W_hi, const float* W_lo, int ldw, const float* bias,
                   float* C, int ldc, int M, int N, int K, int accumulate, cudaStream_t st);
bool gemm_tf32_half_ok(const void* A, int lda, const void* W, int ldw, const void* C, int ldc, int M, int N, int K);
int gemm_tf32_nt_half(const float* A, int lda, const float* W, int ldw, const float* bias, __half* C, int ldc, int M, int N, int K,
                      cudaStream_t st);
int gemm_tf32x3_tn(const float* A_hi, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb, float* C, int ldc,
                   long long R, int P, int Q, cudaStream_t st, int force_splits = 0, int q_valid = 0);
// bf16 / tcgen05 forward (lstm_bf16.cu)
int lstm_forward_bf16(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn,
                      void* ws, size_t ws_bytes, cudaStream_t st);
size_t lstm_workspace_bf16(const bci_lstm_config& c, int batch, int T);
int lstm_pack_bf16(bci_lstm_s* h, cudaStream_t st);
size_t lstm_store_bytes_bf16(const bci_lstm_config& c);
void lstm_carve_bf16(bci_lstm_s* h, char* base);
int pack_pool_bf16(bci_lstm_s* h, cudaStream_t st);
int pack_inproj_bf16(bci_lstm_s* h, cudaStream_t st);
bool input_proj_bf16_ok(const bci_lstm_s* h, const InputView& x, int T);
int launch_input_proj_bf16(bci_lstm_s* h, const InputView& x, int Bc, int T, __nv_bfloat16* z, cudaStream_t st);
int launch_pool_bf16(bci_lstm_s* h, const __nv_bfloat16* seq, float2* stats, float* scores, int Bc, int T, float* logits,
                     float* probs, float* attn, cudaStream_t st);
// single-pass pooling (lstm_bf16_pool_stream.cu)
bool pool_stream_ok(const bci_lstm_s* h, int Bc, int T);
int launch_pool_stream_bf16(bci_lstm_s* h, const __nv_bfloat16* seq, const float2* stats, float* ctx_ws, int Bc, int T, float* logits,
                            float* probs, float* attn, cudaStream_t st);
int launch_fused_rec_bf16(const __nv_bfloat16* in, const __nv_bfloat16* wih, const __nv_bfloat16* whh_f, const __nv_bfloat16* whh_r,
                          const float* bias, __nv_bfloat16* out, float2* stats, int Bc, int T, int Kin, cudaStream_t st);
int fused_max_clusters();  // co-resident 4-CTA clusters of the fused kernel on this device (0 if it cannot run)
// H = 256 bf16 path (lstm_bf16_h256.cu)
int launch_rec256_bf16(const __nv_bfloat16* G, const __nv_bfloat16* whh, __nv_bfloat16* out, int Bc, int T, cudaStream_t st);
int pack_h256_bf16(bci_lstm_s* h, cudaStream_t st);
int pack_pool256_bf16(bci_lstm_s* h, cudaStream_t st);
int launch_pool256_bf16(bci_lstm_s* h, const __nv_bfloat16* seq, __nv_bfloat16* pre, float2* rowstat, float* scores, int Bc, int T,
                        float* logits, float* probs, float* attn, cudaStream_t st);
// CTA-pair form (lstm_bf16_gemm_pair.cu); returns 1 when the shape is not for it
int launch_proj_gemm_bf16_pair(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, __nv_bfloat16* C, int M, int N, int K,
                               bool blocked, cudaStream_t st);
int launch_proj_gemm_bf16(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, __nv_bfloat16* C, int M, int N, int K,
                          bool blocked, cudaStream_t st);

}  // namespace bci
