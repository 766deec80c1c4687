/*
 * bci_b200.h -- C ABI of the B200-native hot path of LSTM-ODE-BCI.
 *
 * The reference (khurrameycon/LSTM-ODE-BCI) is pure Python and has no FFI of its own; its
 * de-facto boundary for this path is the Python call contract of three objects
 * (SURVEY.md §8 b).  Each entry point below names the reference call it replaces
 * (file:line under the reference tree).  INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add to bind them.
 *
 * Conventions
 *   - every function returns 0 on success, a negative BCI_E* code otherwise;
 *     bci_last_error() returns a thread-local message for the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller unless a name ends in
 *     `_host`; the library allocates only the per-handle packed weight copies.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - one handle may be used by one stream at a time (thread-compatible, not thread-safe).
 *   - no C++/torch types cross this boundary.
 */
#ifndef BCI_B200_H
#define BCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCI_ABI_VERSION 5   /* 5: + bci_permute_channels (additive) */
#define BCI_MAX_LAYERS 4

enum {
  BCI_OK = 0,
  BCI_EINVAL = -1,      /* bad argument / unsupported configuration */
  BCI_ECUDA = -2,       /* CUDA runtime / driver error */
  BCI_ENOMEM = -3,      /* workspace too small or allocation failed */
  BCI_ESTATE = -4,      /* call order violated (e.g. forward before load_weights) */
  BCI_EUNSUPPORTED = -5 /* device is not sm_100 */
};

enum { BCI_PRECISION_FP32 = 0, BCI_PRECISION_BF16 = 1 };

int bci_abi_version(void);
const char* bci_last_error(void);
/* Device capability probe: writes SM count, returns BCI_EUNSUPPORTED unless CC 10.x. */
int bci_device_check(int device, int* sm_count);

/* ------------------------------------------------------------------------------------------
 * BiLSTM + attention pooling  (EnhancedLSTMModel, 04_lstm_model.py:153-222)
 * ---------------------------------------------------------------------------------------- */
typedef struct bci_lstm_s* bci_lstm_t;

typedef struct {
  int32_t input_size;    /* C: EEG channels (61)                     04:163 input_size        */
  int32_t hidden_size;   /* H: 128 or 256                            04:163,876-877           */
  int32_t num_layers;    /* 1..BCI_MAX_LAYERS (3)                    04:163                   */
  int32_t num_classes;   /* 2                                        04:164                   */
  int32_t bidirectional; /* 1 (04:164) or 0 (AblationLSTMModel, 09:176-240)                   */
  int32_t precision;     /* BCI_PRECISION_*; BF16 requires the full model (bidirectional, attention,
                            LayerNorm, hidden_size 128)                                       */
  int32_t use_attention; /* 1: attention pooling (04:112-128); 0: mean over time (09:232-234) */
  int32_t use_layer_norm;/* 1: LayerNorm in input_proj and after the LSTM; 0: nn.Identity (09:191,210) */
} bci_lstm_config;

/* Device pointers to fp32 parameters in the reference state-dict layout:
 * PyTorch (out,in) row-major, gate row order i,f,g,o.  [layer][0]=forward, [layer][1]=reverse.
 * With D = H * (bidirectional ? 2 : 1): w_ih of layers >= 1 is (4H, D), ln_* (D), attn_w1 (D/2, D),
 * attn_b1 (D/2), attn_w2 (1, D/2), cls_w0 (H, D).  Pointers of absent modules (reverse direction when
 * unidirectional, attn_* without attention, *ln_* without LayerNorm) are ignored and may be NULL. */
typedef struct {
  const float* input_proj_w;  /* input_proj.0.weight (H,C)   */
  const float* input_proj_b;  /* input_proj.0.bias   (H)     */
  const float* input_ln_w;    /* input_proj.1.weight (H)     */
  const float* input_ln_b;    /* input_proj.1.bias   (H)     */
  const float* w_ih[BCI_MAX_LAYERS][2]; /* lstm.weight_ih_l{k}[_reverse] (4H, H or 2H) */
  const float* w_hh[BCI_MAX_LAYERS][2]; /* lstm.weight_hh_l{k}[_reverse] (4H, H)       */
  const float* b_ih[BCI_MAX_LAYERS][2]; /* lstm.bias_ih_l{k}[_reverse]   (4H)          */
  const float* b_hh[BCI_MAX_LAYERS][2]; /* lstm.bias_hh_l{k}[_reverse]   (4H)          */
  const float* ln_w;          /* layer_norm.weight (2H) */
  const float* ln_b;          /* layer_norm.bias   (2H) */
  const float* attn_w1;       /* attention.attention.0.weight (H,2H) */
  const float* attn_b1;       /* attention.attention.0.bias   (H)    */
  const float* attn_w2;       /* attention.attention.2.weight (1,H)  */
  const float* attn_b2;       /* attention.attention.2.bias   (1)    */
  const float* cls_w0;        /* classifier.0.weight (H,2H)   */
  const float* cls_b0;        /* classifier.0.bias   (H)      */
  const float* cls_w3;        /* classifier.3.weight (H/2,H)  */
  const float* cls_b3;        /* classifier.3.bias   (H/2)    */
  const float* cls_w6;        /* classifier.6.weight (classes,H/2) */
  const float* cls_b6;        /* classifier.6.bias   (classes)     */
} bci_lstm_weights;

/* replaces EnhancedLSTMModel.__init__ (04:163-204): allocates the packed weight store. */
int bci_lstm_create(const bci_lstm_config* cfg, bci_lstm_t* out);
int bci_lstm_destroy(bci_lstm_t h);

/* replaces load_state_dict / .to(device) (06:416-430): repacks the fp32 parameters into the
 * kernels' layouts (transposed fp32 copies; bf16 swizzle-ready copies in BF16 precision).
 * Re-callable after every optimizer step. */
int bci_lstm_load_weights(bci_lstm_t h, const bci_lstm_weights* w, void* stream);

/* Precision of the TRAINING step (bci_lstm_forward(train=1) + bci_lstm_backward), independent of the handle's inference precision.
 *   BCI_TRAIN_FP32  (default) fp32-parity step: gradients within 2e-4 of each tensor's max-abs of torch autograd on the reference
 *                   module in fp32 (04:482-507 without autocast).
 *   BCI_TRAIN_MIXED the analogue of the reference's own GPU training mode, autocast + GradScaler (04:486-490, 499-503): the two
 *                   recurrences of every layer run on the tensor cores with 16-bit operands (forward fp16, BPTT bf16 -- bf16 has
 *                   fp32's exponent range, so no loss scaling is needed) and the large GEMMs in single-pass TF32; accumulators,
 *                   gates, cell states, LayerNorm, softmax, loss and the optimizer stay fp32.  Full-width variants only
 *                   (hidden_size 128: one CTA per 8 windows; 256: CTA pairs). */
enum { BCI_TRAIN_FP32 = 0, BCI_TRAIN_MIXED = 1 };
int bci_lstm_set_train_mode(bci_lstm_t h, int32_t mode);

/* Optional per-phase device timing for bench.py's roofline (SURVEY.md §8 d).  When enabled, forward
 * records CUDA events on the caller's stream between its phases; bci_lstm_get_profile synchronises
 * those events and returns the accumulated milliseconds and launch counts per phase since the last
 * call, then resets them.  Phases: 0 input projection (K1), 1 W_ih projection GEMM (K2),
 * 2 recurrence (K3), 3 LayerNorm+attention pooling+head (K4/K5). */
#define BCI_PROF_PHASES 4
int bci_lstm_set_profiling(bci_lstm_t h, int32_t enable);
int bci_lstm_get_profile(bci_lstm_t h, float ms[BCI_PROF_PHASES], int32_t launches[BCI_PROF_PHASES]);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches claim). */
int64_t bci_launch_count(void);

/* windows processed per internal pass of the inference forward (one full wave of the recurrence kernel on this device:
 * 16 896 = 33 clusters x 4 tiles x 128 for the fused bf16 path on a 148-SM B200); batches that are multiples of it leave no
 * SM idle.  bench.py sizes its step with it. */
int bci_lstm_chunk_windows(bci_lstm_t h, int32_t* windows);

/* bytes of caller-provided scratch needed by forward (train=0) or forward+backward (train=1) */
int bci_lstm_workspace_bytes(bci_lstm_t h, int32_t batch, int32_t seq_len, int32_t train, size_t* bytes);

/* replaces EnhancedLSTMModel.forward(x, return_attention) (04:206-222) followed by the callers'
 * softmax(dim=1) (04:613, 06:232,351, 08:207, 10:231).
 *   x      (B,T,C) fp32 contiguous, batch-first
 *   logits (B,classes) fp32            required
 *   probs  (B,classes) fp32            optional (NULL to skip); column 0 = P(open), 1 = P(closed)
 *   attn   (B,T) fp32                  optional
 *   train  0: eval (dropout off).  1: additionally keeps the activations bci_lstm_backward needs
 *          in the workspace; dropout probabilities are taken from `dropout` (0 = none; the four
 *          sites of 04:177,186,199,202 use dropout/2, dropout, dropout, dropout). */
int bci_lstm_forward(bci_lstm_t h, const float* x, int32_t batch, int32_t seq_len, int32_t train,
                     float dropout, uint64_t seed, float* logits, float* probs, float* attn,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Where the windows of an inference forward live.  The reference materialises every window: X (N,256,61) float32 written by
 * create_sequences (02_preprocessing.py:157-180) and read back by 04/06/08/10.  Those windows overlap by 50 % (step 128 of 256
 * samples, 02:50-51,170), so the same samples cross PCIe and HBM twice.  A view lets the forward read the windows IN PLACE from
 * (samples, C) recordings -- window w of a run starts `window_stride` elements after window w-1 and is seq_len*C contiguous
 * elements long -- and in bf16 as well as fp32 (the bf16 tensor-core mode rounds x to bf16 on load, so bf16 storage changes no
 * result bit there and halves the bytes again).
 *   packed windows, the reference's X:            windows_per_run = 0, window_stride = seq_len*C
 *   R recordings of S samples, (R,S,C) row-major: windows_per_run = (S - seq_len)/step + 1, window_stride = step*C, run_stride = S*C */
enum { BCI_IN_F32 = 0, BCI_IN_BF16 = 1 };
typedef struct {
  const void* data;         /* DEVICE pointer to the first element of window 0                         */
  int32_t dtype;            /* BCI_IN_F32 | BCI_IN_BF16                                                */
  int32_t windows_per_run;  /* windows cut from one run (recording); 0 = all windows form a single run */
  int64_t window_stride;    /* ELEMENTS between the starts of consecutive windows of a run             */
  int64_t run_stride;       /* ELEMENTS between the starts of consecutive runs (ignored for a single run) */
  int64_t first_window;     /* index (in the run/window numbering above) of the call's window 0: lets a caller walk a resident
                               buffer in pass-sized pieces that cross run boundaries                   */
} bci_lstm_input;
/* bci_lstm_forward (train = 0) over a view: window b reads elements [off(b), off(b) + seq_len*C) with
 * off(b) = (w / windows_per_run) * run_stride + (w % windows_per_run) * window_stride, w = first_window + b.  Same outputs, workspace and precision
 * modes as bci_lstm_forward; every window start should be 16-byte aligned for the TMA-fed input projection (otherwise the
 * CUDA-core projection kernel runs). */
int bci_lstm_forward_view(bci_lstm_t h, const bci_lstm_input* in, int32_t batch, int32_t seq_len, float* logits, float* probs,
                          float* attn, void* workspace, size_t workspace_bytes, void* stream);

/* Gradient pointers, same layout/shape as bci_lstm_weights; every pointer of a module the configuration has
 * must be non-NULL.  Gradients are OVERWRITTEN (not accumulated). */
typedef struct {
  float* input_proj_w; float* input_proj_b; float* input_ln_w; float* input_ln_b;
  float* w_ih[BCI_MAX_LAYERS][2]; float* w_hh[BCI_MAX_LAYERS][2];
  float* b_ih[BCI_MAX_LAYERS][2]; float* b_hh[BCI_MAX_LAYERS][2];
  float* ln_w; float* ln_b; float* attn_w1; float* attn_b1; float* attn_w2; float* attn_b2;
  float* cls_w0; float* cls_b0; float* cls_w3; float* cls_b3; float* cls_w6; float* cls_b6;
} bci_lstm_grads;

/* replaces loss.backward() through the model (04:490-494; 07:242-258 for dx): BPTT from
 * dlogits (B,classes).  Must follow a train=1 forward on the same workspace.  dx (B,T,C)
 * optional. */
int bci_lstm_backward(bci_lstm_t h, const float* x, const float* dlogits, int32_t batch, int32_t seq_len,
                      float* dx, const bci_lstm_grads* grads, void* workspace, size_t workspace_bytes,
                      void* stream);

/* replaces criterion(outputs, y) / accumulation_steps and its backward to the logits (04:456-458,486-494: CrossEntropyLoss with
 * class weights, mean reduction = sum_i w[y_i] nll_i / sum_i w[y_i]):
 *   logits (B,classes) fp32, labels (B) int64, class_weight (classes) fp32 or NULL (all ones),
 *   loss_scale multiplies both outputs (1/accumulation_steps; a GradScaler scale would go here too),
 *   loss_out (1) fp32 receives the scaled loss, dlogits (B,classes) fp32 its gradient.  One block, fixed-order reductions. */
int bci_ce_loss_grad(const float* logits, const int64_t* labels, const float* class_weight, int32_t batch, int32_t classes,
                     float loss_scale, float* loss_out, float* dlogits, void* stream);
/* gradient accumulation over micro-batches (04:489,497: accumulation_steps = 4): acc = (first ? 0 : acc) + g over n floats */
int bci_grad_accumulate(float* acc, const float* g, int64_t n, int32_t first, void* stream);

/* replaces clip_grad_norm_ + AdamW.step on a flat fp32 bucket (04:497-507, 04:438): p, g, m, v
 * are flat arrays of n floats; `grad_scale` multiplies g first (1/world_size after an NCCL
 * sum all-reduce, 1/loss_scale for AMP); if max_norm > 0 the scaled gradient is clipped to that
 * global L2 norm.  `norm_scratch` is 2 floats of device scratch; the pre-clip norm is left in
 * norm_scratch[1]. */
int bci_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   float max_norm, float* norm_scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel optimizer step over peer memory (BASELINE config 3; SURVEY.md §8 b `bci_fused_step`, §8 e).
 * The reference trains on one device (04:482-507); this is its loop body for N ranks: the gradient
 * all-reduce is fused into the load phase of the clip + AdamW kernels -- each rank reads every peer's
 * gradient bucket directly over NVLink (CUDA IPC peer pointers, system-scope flags), sums in rank order and
 * updates its replica.  Results are bit-identical on all ranks.
 *   1. every rank: bci_comm_create -> bci_comm_export (BCI_COMM_HANDLE_BYTES host bytes)
 *   2. exchange the handles (torch.distributed all_gather_object / any host channel), rank order
 *   3. every rank: bci_comm_connect(all handles)
 *   4. per step: write local gradients into bci_comm_bucket (e.g. bci_lstm_backward with grads pointing into
 *      it), then bci_fused_step on the same stream.  All ranks must call it the same number of times.
 * ---------------------------------------------------------------------------------------- */
typedef struct bci_comm_s* bci_comm_t;
#define BCI_COMM_HANDLE_BYTES 128
int bci_comm_create(int32_t rank, int32_t world, int64_t n_floats, bci_comm_t* out);
int bci_comm_export(bci_comm_t c, void* handle_host);
int bci_comm_connect(bci_comm_t c, const void* handles_host /* world x BCI_COMM_HANDLE_BYTES */);
/* device pointer of this rank's gradient bucket (n_floats fp32, owned by the communicator) */
int bci_comm_bucket(bci_comm_t c, float** bucket, int64_t* n_floats);
int bci_comm_destroy(bci_comm_t c);
/* replaces all-reduce(mean) + clip_grad_norm_(max_norm) + AdamW.step (04:497-507) on flat p/m/v of n_floats;
 * norm_out (optional, 2 device floats): [0] squared norm of the summed gradient, [1] pre-clip norm of the mean */
int bci_fused_step(bci_comm_t c, float* p, float* m, float* v, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int32_t step, float max_norm, float* norm_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Three-state A/P/F ODE ensemble with probabilistic rate coupling
 * ---------------------------------------------------------------------------------------- */
enum { BCI_ODE_RK4 = 0, BCI_ODE_RK45 = 1 };
/* REF06: clamp max(0,.) in the RHS, y0 /= sum(y0), t = linspace(0,t_end,n_points), then
 *        clip[0,1] + row renormalisation   (CognitiveStateODE.solve, 06:174-180 == 05:137-169)
 * REF08: raw RHS, y0 as given, no post-processing (predict_trajectory, 08:149-153; pass
 *        t_end = n_steps*dt and n_points = n_steps+1) */
enum { BCI_ODE_STYLE_REF06 = 0, BCI_ODE_STYLE_REF08 = 1 };
/* where the initial state comes from */
enum {
  BCI_Y0_GIVEN = 0,          /* y0 (3,N) SoA                                               */
  BCI_Y0_FROM_PROBS_06 = 1,  /* thresholds on P(closed)/P(open): 06:377-382 == 10:250-255   */
  BCI_Y0_FROM_PCLOSED_08 = 2 /* prob_to_ode_state(P(closed)): 08:215-234                    */
};
enum { BCI_OUT_F32 = 0, BCI_OUT_F64 = 1 };

typedef struct {
  int32_t mode;        /* BCI_ODE_RK4 | BCI_ODE_RK45 */
  int32_t style;       /* BCI_ODE_STYLE_* */
  int32_t y0_mode;     /* BCI_Y0_* */
  int32_t coupling;    /* 1: modulate_ode_rates (06:236-264): k_af,k_pf *= 1+alpha*P(closed);
                          k_fa,k_pa *= 1+alpha*P(open); all six floored at 0.001.  0: rates used as is */
  int64_t n;           /* trajectories */
  float base_rates[6]; /* k_ap,k_af,k_pa,k_pf,k_fa,k_fp used when rates == NULL */
  float alpha;         /* coupling strength used when alpha_arr == NULL (06:204) */
  const float* rates;     /* optional (6,N) SoA per-trajectory base rates */
  const float* alpha_arr; /* optional (N) */
  const float* p_open;    /* (N), required if coupling or BCI_Y0_FROM_PROBS_06 */
  const float* p_closed;  /* (N), required if coupling or y0_mode != GIVEN */
  const float* y0;        /* (3,N) SoA, required if BCI_Y0_GIVEN */
  double t_end;
  int32_t n_points;    /* >= 2 */
  int32_t substeps;    /* RK4: equal steps per output interval; 0 = per trajectory so that
                          0.01*(h*lambda)^4 <= 2e-7 (lambda = largest total outflow rate) */
  double rtol, atol;   /* RK45: scipy.solve_ivp defaults 1e-3 / 1e-6 (05:158-163) */
  int32_t out_dtype;   /* BCI_OUT_F32 | BCI_OUT_F64: element type of traj / final_state */
  void* traj;          /* optional (N,n_points,3) */
  void* final_state;   /* optional (N,3): last row of the trajectory (10:271-273) */
  int32_t* n_steps;    /* optional (N): RK45 accepted+rejected steps / RK4 steps taken */
} bci_ode_args;

/* replaces the per-sample host loop of predict_batch step 2 (06:372-401),
 * get_three_state_probabilities step 2 (10:245-273), multistep_forecast (08:264-282) and
 * CognitiveStateODE.solve / predict_trajectory themselves: one launch, one thread per trajectory. */
int bci_ode_solve(const bci_ode_args* args, void* stream);

/* replaces CognitiveStateODE.solve_with_modulation (05_ode_model.py:171-196): rates that vary with time.  The reference
 * calls a Python `modulation_func(t, params)` from inside LSODA; here the caller samples it once at the stage times of a
 * fixed-step RK4 -- node m at t0 + m*h/2, h = t_span / (n_points-1) / substeps, M = 2*substeps*(n_points-1) + 1 nodes,
 * rate order k_ap,k_af,k_pa,k_pf,k_fa,k_fp -- and N trajectories integrate in one launch (fp64).
 *   rate_nodes  (M,6) fp64 shared by all trajectories, or (M,6,N) when per_trajectory != 0
 *   y0          (3,N) fp64 SoA;  style REF06: y0 /= sum, max(0,.) clamp, clip + renormalise (05:184-194); REF08: raw
 *   traj        optional (N,n_points,3) fp64;  final_state optional (N,3) fp64 */
typedef struct {
  int64_t n;
  int32_t style;
  int32_t n_points;
  int32_t substeps;
  int32_t per_trajectory;
  double t_span;
  const double* rate_nodes;
  const double* y0;
  double* traj;
  double* final_state;
} bci_ode_mod_args;
int bci_ode_solve_modulated(const bci_ode_mod_args* args, void* stream);

/* replaces the read-outs that follow the solve:
 *   pred06[i]  = traj_last[i].F > 0.5                          (06:396-401)
 *   cls10[i]   = 2 if F > .5 else 0 if A > .5 else 1           (10:282-288)
 * from final_state (N,3) fp32; either output may be NULL. */
int bci_ode_classify(const float* final_state, int64_t n, int32_t* pred06, int32_t* cls10, void* stream);
/* forecast read-out clip(F_h + 0.5 P_h, 0, 1) at `n_h` horizons (08:273-279) from traj
 * (N,n_points,3) fp32 -> out (N,n_h) fp32.  horizons_host is a HOST array. */
int bci_ode_forecast_readout(const float* traj, int64_t n, int32_t n_points, const int32_t* horizons_host,
                             int32_t n_h, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Raw recording -> model windows (SURVEY.md §8 f row 4; 02_preprocessing.py:114-180)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_recordings;  /* R equally long recordings processed in one call                        */
  int32_t n_channels;    /* C (61)                                                                  */
  int64_t n_samples;     /* samples per recording (150 000 = 300 s x 500 Hz)                        */
  int32_t in_dtype;      /* BCI_OUT_F64 (mne's raw.get_data(), 02:200) or BCI_OUT_F32               */
  int32_t order;         /* len(a) - 1 == len(b) - 1; butter(4, ..., 'band') -> 8   (02:128-130)    */
  const double* b_host;  /* HOST: numerator   (order + 1)                                           */
  const double* a_host;  /* HOST: denominator (order + 1)                                           */
  const double* zi_host; /* HOST: scipy.signal.lfilter_zi(b, a) (order)                             */
  int32_t padlen;        /* filtfilt edge padding; scipy default 3 * max(len(a), len(b)) = 27       */
  int32_t seq_len;       /* SEQUENCE_LENGTH 256 (02:50)                                             */
  int32_t step;          /* int(seq_len * (1 - overlap)) = 128 (02:51,170)                          */
  const double* mean_in; /* optional DEVICE (C): reuse normalisation parameters (02:207-210) ...    */
  const double* std_in;  /* ... together with std_in; NULL = per-recording statistics (02:212)      */
} bci_preproc_args;

int bci_preprocess_workspace_bytes(const bci_preproc_args* a, size_t* bytes);
/* replaces bandpass_filter + normalize_data + create_sequences (02:114-180) for a batch of recordings:
 *   raw      (R, C, n_samples) DEVICE, in_dtype
 *   windows  (R * n_seq, seq_len, C) fp32, n_seq = (n_samples - seq_len) / step + 1   -- the LSTM's input layout
 *   mean_out, std_out (R, C) fp64: the statistics used (02:145-151; std floored at 1e-10)
 *   filtered optional (R, C, n_samples) fp64: the band-passed signal before normalisation */
int bci_preprocess(const bci_preproc_args* a, const void* raw, float* windows, double* mean_out, double* std_out,
                   double* filtered, void* workspace, size_t workspace_bytes, void* stream);

/* Host-side staging copy of the drop-in callers (06_lstm_ode_integration.py:346 `torch.FloatTensor(X[i:i+bs]).to(device)`: a
 * synchronous pageable copy per batch).  Copies n fp32 values from pageable `src` into the pinned staging buffer `dst` on `threads`
 * host threads with streaming stores; to_bf16 = 1 narrows to bf16 on the way (round to nearest even, bit-identical to the
 * conversion the input projection does on load, so the bf16 engine's results do not change while the PCIe bytes halve).
 * Pure host code: no CUDA call, no stream. */
int bci_host_stage(void* dst, const float* src, int64_t n, int32_t to_bf16, int32_t threads);

/* Permutation-importance inputs (SURVEY.md §8 f rank 1; replaces the host-side `X_permuted = X_subset.copy();
 * X_permuted[:, :, ch] = X_subset[perm_idx, :, ch]` + per-batch upload of 07_explainability.py:336-339,312-322).  The test subset
 * x (n, seq_len, channels) fp32 stays on the DEVICE; rows [row0, row0 + rows) of the V x n stack of variants are written in the
 * forward's input layout:   out[r][t][c] = x[c == channel[v] ? perm[g] : i][t][c],   g = row0 + r, v = g / n, i = g % n
 *   perm     (V*n) int32 DEVICE: numpy's np.random.permutation(n) of each variant, concatenated (07:338)
 *   channel  (V)   int32 DEVICE: the permuted channel of each variant; < 0 = unpermuted copy (the baseline sweep, 07:325)
 *   out      (rows, seq_len, channels) fp32 (BCI_IN_F32) or bf16 (BCI_IN_BF16: rounded exactly as the bf16 engine rounds x on load) */
int bci_permute_channels(const float* x, int32_t n, int32_t seq_len, int32_t channels, const int32_t* perm,
                         const int32_t* channel, int64_t row0, int64_t rows, int32_t out_dtype, void* out, void* stream);

/* Micro-benchmark used by bench.py for the FP32 roofline denominator (SURVEY.md §8 d: the FP32
 * FMA peak is not in MEASURED_PEAKS.json): launches a dependent-FMA kernel, returns TFLOP/s. */
int bci_fp32_peak_probe(double* tflops, void* stream);
/* the same for the FP64 pipe (roofline denominator of the preprocessing recursion) */
int bci_fp64_peak_probe(double* tflops, void* stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics: the two tensor-core kernels of the bf16 path, callable in isolation so the GPU unit
 * tests can check each against a plain matmul / a step-by-step recurrence.
 *   proj_gemm: C[M][N] (bf16) = A[M][K] (bf16) . W[N][K]^T (bf16) + bias[N] (fp32);  N % 256 == 0, K % 64 == 0
 *   rec:       G bf16 in the blocked streaming layout [row/128][dir*64 + perm_G/8][row%128][perm_G%8], row = t*Bc + b,
 *              padded to whole 128-row blocks (bias included, i/f/o pre-activations pre-scaled by 1/2);
 *              whh_* [512][128] bf16 with rows in perm_T(unit,gate) order -> out [T][Bc][256] bf16, where for
 *              unit = half*64 + slab*8 + u:  perm_T = half*256 + slab*32 + gate*8 + u,
 *                                            perm_G = slab*64 + half*32 + gate*8 + u
 * ---------------------------------------------------------------------------------------- */
int bci_selftest_proj_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N,
                                int32_t K, void* stream);
int bci_selftest_proj_gemm_bf16_blocked(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N,
                                        int32_t K, void* stream);
int bci_selftest_rec_bf16(const void* G, const void* whh_f, const void* whh_r, void* out, int32_t Bc, int32_t T,
                          void* stream);
/* fused projection + recurrence of one layer on a 4-CTA cluster (csrc/lstm_bf16_fused.cu):
 *   in [T][Bc][Kin] bf16 (Kin = 128 or 256); wih [2][512][Kin], whh_* [512][128] bf16 with rows in perm_T order (i,f,o rows
 *   pre-scaled by 1/2); bias [2][512] fp32 in the same order and scaling -> out [T][Bc][256] bf16; stats optional
 *   [T][8][Bc] float2 partial (sum, sum of squares) of h over 32 units, slot = dir*4 + unit/32 */
/* H = 256 cluster recurrence (csrc/lstm_bf16_h256.cu): G bf16 in the blocked streaming layout [row/128][256 chunks][row%128][8] with
 * columns dir*1024 + perm_256(unit, gate) = (unit/128)*512 + ((unit/64)%2)*256 + ((unit/8)%8)*32 + gate*8 + unit%8 (bias included,
 * i/f/o pre-scaled by 1/2); whh [2][1024][256] bf16 rows in the same order -> out [T][Bc][512] bf16 */
int bci_selftest_rec256_bf16(const void* G, const void* whh, void* out, int32_t Bc, int32_t T, void* stream);
int bci_selftest_fused_rec_bf16(const void* in, const void* wih, const void* whh_f, const void* whh_r, const float* bias,
                                void* out, void* stats, int32_t Bc, int32_t T, int32_t Kin, void* stream);

/* split-precision (3 x TF32) tcgen05 GEMMs of the fp32 path (csrc/gemm_tf32x3.cu), fp32 in / fp32 out:
 *   mode 0 (NT): C[M][N] = A[M][K] . B[N][K]^T + bias[N] (bias optional)
 *   mode 1 (TN): C[M][N] = A[K][M]^T . B[K][N]            (split-K, partial tiles reduce-added)
 * explicit_hi != 0: the tf32-rounded high parts are separate arrays instead of the raw operands. */
int bci_selftest_gemm_tf32x3(int32_t mode, const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N,
                             int64_t K, int32_t explicit_hi, void* stream);

/* single-pass TF32 form of the NT product (mixed training step): C[M][N] (=|+=) A[M][K] . B[N][K]^T + bias; M >= 512 and N >= 256 run on
 * CTA pairs (cta_group::2, 256 x 256 tiles); _tn: C[M][N] = A[K][M]^T . B[K][N] (weight gradients; pairs when M % 256 == 0, N % 128 == 0) */
int bci_selftest_gemm_tf32_single(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                                  int32_t accumulate, void* stream);

int bci_selftest_gemm_tf32_single_tn(const float* A, const float* B, float* C, int32_t M, int32_t N, int64_t K, void* stream);

/* the same NT product with both operands split into FP16 (hi, lo) pairs (kind::f16, twice the MMA rate; forward projections of the
 * fp32 inference path): C[M][N] = A[M][K] . B[N][K]^T + bias, fp32 in / fp32 out */
int bci_selftest_gemm_f16x3(const float* A, const float* B, const float* bias, float* C, int32_t M, int32_t N, int32_t K, void* stream);

/* fp32-grade recurrence on the tensor cores (csrc/lstm_fp32_tc.cu: h . W_hh^T as three fp16 tcgen05 MMA chains on a CTA pair):
 *   G [T*Bc][ND*512] fp32, column dir*512 + unit*4 + gate (bias included); w_hh [ND][512][128] fp32 in the PyTorch layout
 *   (gate-major rows i,f,g,o); packed: [ND][2][512][128] fp16 scratch (filled here); out [T][Bc][ND*128] fp32 */
int bci_selftest_rec_f16x3(const float* G, const float* w_hh, void* packed, float* out, int32_t Bc, int32_t T, int32_t ND,
                           void* stream);

/* swapped (weights-as-A-operand) tensor-core recurrences of the mixed-precision training step (csrc/lstm_rec_swap.cu), in isolation:
 *   G / gates / dG [T*Bc][ND*512] fp32, column dir*512 + unit*4 + gate; out / csave / dout [T][Bc][ND*128] fp32; w_hh [ND][512][128]
 *   fp32 in the PyTorch layout; packed: 5 x ND x 512 x 128 16-bit values of scratch (filled here); split != 0: the fp32-parity form
 *   of the forward (three fp16 product chains, accurate gate activations) */
int bci_selftest_rec_swap_fwd(const float* G, const float* w_hh, void* packed, float* out, float* gates, float* csave, int32_t Bc,
                              int32_t T, int32_t ND, int32_t split, void* stream);
int bci_selftest_bptt_swap(const float* dout, const float* gates, const float* csave, const float* w_hh, void* packed, float* dG,
                           int32_t Bc, int32_t T, int32_t ND, int32_t split, void* stream);
/* the hidden-size-256 pair kernels (mixed precision): G / gates / dG [T*Bc][ND*1024] (column dir*1024 + unit*4 + gate), out / csave /
 * dout [T][Bc][ND*256], w_hh [ND][1024][256] fp32; packed: 5 x ND x 1024 x 256 sixteen-bit values of scratch */
int bci_selftest_rec_swap256_fwd(const float* G, const float* w_hh, void* packed, float* out, float* gates, float* csave, int32_t Bc,
                                 int32_t T, int32_t ND, void* stream);
int bci_selftest_bptt_swap256(const float* dout, const float* gates, const float* csave, const float* w_hh, void* packed, float* dG,
                              int32_t Bc, int32_t T, int32_t ND, void* stream);
/* the stateless dropout mask of the training step (a hash of seed / site / element index): out[i] = 0 or 1/(1-p) for i < n */
int bci_selftest_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint32_t site, void* stream);
/* selftest only: clock64 stamps (8 per step, steps 100-103; int64[32]) of CTA (0,0) of the following swapped forward launches */
int bci_selftest_swap_set_debug(long long* stamps);
/* layout probe: one M128 x N16 x K16 tcgen05.mma whose A operand is read from tensor memory; out [128][16] fp32 */
int bci_selftest_tmem_a_probe(float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BCI_B200_H */
