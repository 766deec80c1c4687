// extern "C" entry points of libbci_b200.so (see include/bci_b200.h).
#include "lstm_handle.cuh"
#include <cstring>
#include <cstdlib>
#include <new>

namespace bci {

static thread_local char g_err[512] = "";

static long long g_launches = 0;
void note_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int lstm_pack_f32(bci_lstm_s* h, cudaStream_t st);
size_t lstm_store_bytes_f32(const bci_lstm_config& c);
void lstm_carve_f32(bci_lstm_s* h, char* base);
int lstm_forward_train(bci_lstm_s* h, const float* x, int batch, int T, float dropout, uint64_t seed, float* logits,
                       float* probs, float* attn, void* ws, size_t ws_bytes, cudaStream_t st);
size_t lstm_workspace_train(const bci_lstm_config& c, int batch, int T);
int lstm_backward_impl(bci_lstm_s* h, const float* x, const float* dlogits, int batch, int T, float* dx,
                       const bci_lstm_grads* grads, void* ws, size_t ws_bytes, cudaStream_t st);

// Makes the handle's device current for the duration of an entry point (a caller driving several GPUs from one process may have
// another one current); restores the previous device on exit.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
    if (cur != dev) {
      ok = cudaSetDevice(dev) == cudaSuccess;
      if (ok) prev = cur;
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define BCI_ON_DEVICE_OF(h)                                                                             \
  DeviceGuard _guard((h)->device);                                                                      \
  BCI_REQUIRE(_guard.ok, BCI_ECUDA, "cannot make device %d (the handle's) current", (h)->device)

}  // namespace bci

using namespace bci;

extern "C" int bci_abi_version(void) { return BCI_ABI_VERSION; }
extern "C" const char* bci_last_error(void) { return g_err; }

extern "C" int64_t bci_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int bci_lstm_set_profiling(bci_lstm_t h, int32_t enable) {
  BCI_REQUIRE(h, BCI_EINVAL, "bci_lstm_set_profiling: NULL handle");
  h->prof.enabled = enable != 0;
  h->prof.n = 0;
  return BCI_OK;
}

extern "C" int bci_lstm_get_profile(bci_lstm_t h, float ms[BCI_PROF_PHASES], int32_t launches[BCI_PROF_PHASES]) {
  BCI_REQUIRE(h && ms && launches, BCI_EINVAL, "bci_lstm_get_profile: NULL argument");
  BCI_ON_DEVICE_OF(h);
  for (int i = 0; i < BCI_PROF_PHASES; ++i) { ms[i] = 0.f; launches[i] = 0; }
  Profiler& p = h->prof;
  if (p.n > 0) BCI_CUDA_OK(cudaEventSynchronize(p.ev[p.n - 1]));
  for (int i = 1; i < p.n; ++i) {
    if (p.phase[i] < 0) continue;
    float t = 0.f;
    BCI_CUDA_OK(cudaEventElapsedTime(&t, p.ev[i - 1], p.ev[i]));
    ms[p.phase[i]] += t;
    launches[p.phase[i]] += 1;
  }
  p.n = 0;
  return BCI_OK;
}

extern "C" int bci_device_check(int device, int* sm) {
  int n = 0;
  BCI_CUDA_OK(cudaGetDeviceCount(&n));
  BCI_REQUIRE(device >= 0 && device < n, BCI_EINVAL, "bci_device_check: device %d of %d", device, n);
  int major = 0, minor = 0, sms = 0;
  BCI_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  BCI_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  BCI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (sm) *sm = sms;
  BCI_REQUIRE(major == 10, BCI_EUNSUPPORTED, "bci_b200 is built for sm_100a only; device %d is sm_%d%d", device, major, minor);
  return BCI_OK;
}

extern "C" int bci_lstm_create(const bci_lstm_config* cfg, bci_lstm_t* out) {
  BCI_REQUIRE(cfg && out, BCI_EINVAL, "bci_lstm_create: NULL argument");
  BCI_REQUIRE(cfg->hidden_size == 128 || cfg->hidden_size == 256, BCI_EINVAL,
              "bci_lstm_create: hidden_size must be 128 or 256 (got %d)", cfg->hidden_size);
  BCI_REQUIRE(cfg->input_size >= 1 && cfg->input_size <= 64, BCI_EINVAL, "bci_lstm_create: input_size must be 1..64 (got %d)",
              cfg->input_size);
  BCI_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= BCI_MAX_LAYERS, BCI_EINVAL, "bci_lstm_create: num_layers must be 1..%d",
              BCI_MAX_LAYERS);
  BCI_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 8, BCI_EINVAL, "bci_lstm_create: num_classes must be 1..8");
  BCI_REQUIRE((cfg->bidirectional | 1) == 1 && (cfg->use_attention | 1) == 1 && (cfg->use_layer_norm | 1) == 1, BCI_EINVAL,
              "bci_lstm_create: bidirectional, use_attention and use_layer_norm must be 0 or 1");
  BCI_REQUIRE(cfg->precision == BCI_PRECISION_FP32 || cfg->precision == BCI_PRECISION_BF16, BCI_EINVAL,
              "bci_lstm_create: bad precision %d", cfg->precision);
  BCI_REQUIRE(cfg->precision == BCI_PRECISION_FP32 || (cfg->bidirectional && cfg->use_attention && cfg->use_layer_norm), BCI_EINVAL,
              "bci_lstm_create: the bf16 (tcgen05) mode is built for the full model only (bidirectional, attention pooling, "
              "LayerNorm); the ablation variants of 09_sensitivity_analysis.py run in fp32 precision");
  bci_lstm_s* h = new (std::nothrow) bci_lstm_s();
  BCI_REQUIRE(h, BCI_ENOMEM, "bci_lstm_create: host allocation failed");
  std::memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e != cudaSuccess) { delete h; set_error("cudaGetDevice failed: %s", cudaGetErrorString(e)); return BCI_ECUDA; }
  const size_t f32_bytes = lstm_store_bytes_f32(*cfg);
  const size_t bf_bytes = lstm_store_bytes_bf16(*cfg);
  h->store_bytes = f32_bytes + bf_bytes;
  e = cudaMalloc(&h->store, h->store_bytes);
  if (e != cudaSuccess) { delete h; set_error("cudaMalloc(%zu) for packed weights failed: %s", f32_bytes + bf_bytes, cudaGetErrorString(e)); return BCI_ENOMEM; }
  lstm_carve_f32(h, (char*)h->store);
  lstm_carve_bf16(h, (char*)h->store + f32_bytes);
  *out = h;
  return BCI_OK;
}

extern "C" int bci_lstm_destroy(bci_lstm_t h) {
  if (!h) return BCI_OK;
  DeviceGuard _guard(h->device);
  if (h->store) cudaFree(h->store);
  if (h->prof.created) for (int i = 0; i < Profiler::MAX_EV; ++i) cudaEventDestroy(h->prof.ev[i]);
  if (h->side_ready) {
    cudaStreamDestroy(h->side);
    cudaEventDestroy(h->ev_dg); cudaEventDestroy(h->ev_side[0]); cudaEventDestroy(h->ev_side[1]); cudaEventDestroy(h->ev_join);
  }
  delete h;
  return BCI_OK;
}

extern "C" int bci_lstm_load_weights(bci_lstm_t h, const bci_lstm_weights* w, void* stream) {
  BCI_REQUIRE(h && w, BCI_EINVAL, "bci_lstm_load_weights: NULL argument");
  BCI_ON_DEVICE_OF(h);
  const bci_lstm_config& c = h->cfg;
  const void* must[] = {w->input_proj_w, w->input_proj_b, w->cls_w0, w->cls_b0, w->cls_w3, w->cls_b3, w->cls_w6, w->cls_b6};
  for (const void* p : must) BCI_REQUIRE(p, BCI_EINVAL, "bci_lstm_load_weights: a required weight pointer is NULL");
  if (c.use_layer_norm)
    BCI_REQUIRE(w->input_ln_w && w->input_ln_b && w->ln_w && w->ln_b, BCI_EINVAL, "bci_lstm_load_weights: a LayerNorm pointer is NULL");
  if (c.use_attention)
    BCI_REQUIRE(w->attn_w1 && w->attn_b1 && w->attn_w2 && w->attn_b2, BCI_EINVAL, "bci_lstm_load_weights: an attention pointer is NULL");
  for (int l = 0; l < c.num_layers; ++l)
    for (int d = 0; d < num_dirs(c); ++d)
      BCI_REQUIRE(w->w_ih[l][d] && w->w_hh[l][d] && w->b_ih[l][d] && w->b_hh[l][d], BCI_EINVAL,
                  "bci_lstm_load_weights: LSTM weight pointer NULL (layer %d dir %d)", l, d);
  h->raw = *w;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = lstm_pack_f32(h, st);
  if (rc) return rc;
  if (c.precision == BCI_PRECISION_BF16) {
    rc = lstm_pack_bf16(h, st);
    if (rc) return rc;
  }
  h->loaded = true;
  return BCI_OK;
}

extern "C" int bci_lstm_set_train_mode(bci_lstm_t h, int32_t mode) {
  BCI_REQUIRE(h, BCI_EINVAL, "bci_lstm_set_train_mode: NULL handle");
  BCI_REQUIRE(mode == BCI_TRAIN_FP32 || mode == BCI_TRAIN_MIXED, BCI_EINVAL, "bci_lstm_set_train_mode: mode must be BCI_TRAIN_FP32 or BCI_TRAIN_MIXED");
  h->train_mode = mode;
  return BCI_OK;
}

extern "C" int bci_lstm_chunk_windows(bci_lstm_t h, int32_t* windows) {
  BCI_REQUIRE(h && windows, BCI_EINVAL, "bci_lstm_chunk_windows: NULL argument");
  *windows = max_chunk(h->cfg, 0);
  return BCI_OK;
}

extern "C" int bci_lstm_workspace_bytes(bci_lstm_t h, int32_t batch, int32_t seq_len, int32_t train, size_t* bytes) {
  BCI_REQUIRE(h && bytes, BCI_EINVAL, "bci_lstm_workspace_bytes: NULL argument");
  BCI_REQUIRE(batch >= 0 && seq_len >= 1, BCI_EINVAL, "bci_lstm_workspace_bytes: bad shape (%d,%d)", batch, seq_len);
  if (train) *bytes = lstm_workspace_train(h->cfg, batch, seq_len);
  else *bytes = h->cfg.precision == BCI_PRECISION_BF16 ? lstm_workspace_bf16(h->cfg, batch, seq_len)
                                                        : lstm_workspace_fp32(h->cfg, batch, seq_len);
  return BCI_OK;
}

static int forward_infer(bci_lstm_t h, const InputView& v, int batch, int T, float* logits, float* probs, float* attn, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  if (h->cfg.precision == BCI_PRECISION_BF16) return lstm_forward_bf16(h, v, batch, T, logits, probs, attn, ws, ws_bytes, st);
  return lstm_forward_fp32(h, v, batch, T, logits, probs, attn, ws, ws_bytes, st);
}

extern "C" int bci_lstm_forward(bci_lstm_t h, const float* x, int32_t batch, int32_t seq_len, int32_t train, float dropout,
                                uint64_t seed, float* logits, float* probs, float* attn, void* workspace,
                                size_t workspace_bytes, void* stream) {
  BCI_REQUIRE(h, BCI_EINVAL, "bci_lstm_forward: NULL handle");
  BCI_REQUIRE(h->loaded, BCI_ESTATE, "bci_lstm_forward: call bci_lstm_load_weights first");
  BCI_REQUIRE(batch >= 0 && seq_len >= 1 && seq_len <= 65536, BCI_EINVAL, "bci_lstm_forward: bad shape (%d,%d)", batch, seq_len);
  if (batch == 0) return BCI_OK;
  BCI_REQUIRE(x && logits, BCI_EINVAL, "bci_lstm_forward: x and logits are required");
  BCI_REQUIRE(workspace, BCI_ENOMEM, "bci_lstm_forward: workspace is NULL");
  BCI_REQUIRE(dropout >= 0.f && dropout < 1.f, BCI_EINVAL, "bci_lstm_forward: dropout must be in [0,1)");
  BCI_ON_DEVICE_OF(h);
  cudaStream_t st = (cudaStream_t)stream;
  if (train) return lstm_forward_train(h, x, batch, seq_len, dropout, seed, logits, probs, attn, workspace, workspace_bytes, st);
  return forward_infer(h, packed_view(x, seq_len, h->cfg.input_size), batch, seq_len, logits, probs, attn, workspace, workspace_bytes, st);
}

extern "C" int bci_lstm_forward_view(bci_lstm_t h, const bci_lstm_input* in, int32_t batch, int32_t seq_len, float* logits,
                                     float* probs, float* attn, void* workspace, size_t workspace_bytes, void* stream) {
  BCI_REQUIRE(h && in, BCI_EINVAL, "bci_lstm_forward_view: NULL argument");
  BCI_REQUIRE(h->loaded, BCI_ESTATE, "bci_lstm_forward_view: call bci_lstm_load_weights first");
  BCI_REQUIRE(batch >= 0 && seq_len >= 1 && seq_len <= 65536, BCI_EINVAL, "bci_lstm_forward_view: bad shape (%d,%d)", batch, seq_len);
  if (batch == 0) return BCI_OK;
  BCI_REQUIRE(in->data && logits, BCI_EINVAL, "bci_lstm_forward_view: data and logits are required");
  BCI_REQUIRE(workspace, BCI_ENOMEM, "bci_lstm_forward_view: workspace is NULL");
  BCI_REQUIRE(in->dtype == BCI_IN_F32 || in->dtype == BCI_IN_BF16, BCI_EINVAL, "bci_lstm_forward_view: dtype must be BCI_IN_F32 or BCI_IN_BF16");
  BCI_REQUIRE(in->window_stride >= 1 && in->windows_per_run >= 0 && (in->windows_per_run == 0 || in->run_stride >= 0), BCI_EINVAL,
              "bci_lstm_forward_view: window_stride must be >= 1, windows_per_run and run_stride >= 0");
  BCI_ON_DEVICE_OF(h);
  BCI_REQUIRE(in->first_window >= 0, BCI_EINVAL, "bci_lstm_forward_view: first_window must be >= 0");
  const InputView v{in->data, in->dtype, in->windows_per_run, in->window_stride, in->windows_per_run ? in->run_stride : 0,
                    in->first_window};
  return forward_infer(h, v, batch, seq_len, logits, probs, attn, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int bci_lstm_backward(bci_lstm_t h, const float* x, const float* dlogits, int32_t batch, int32_t seq_len, float* dx,
                                 const bci_lstm_grads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  bci::NvtxRange nvtx_range("bci_lstm_backward");
  BCI_REQUIRE(h && x && dlogits && grads && workspace, BCI_EINVAL, "bci_lstm_backward: NULL argument");
  BCI_REQUIRE(h->loaded, BCI_ESTATE, "bci_lstm_backward: call bci_lstm_load_weights first");
  BCI_ON_DEVICE_OF(h);
  return lstm_backward_impl(h, x, dlogits, batch, seq_len, dx, grads, workspace, workspace_bytes, (cudaStream_t)stream);
}
