// GPU preprocessing of raw EEG recordings into the LSTM's input windows (SURVEY.md §8 f row 4): the step
// immediately in front of the hot path.  Reference: 02_preprocessing.py:114-180 --
//   bandpass_filter  (114-131): scipy.signal.filtfilt(b, a, data, axis=1) with b, a = butter(4, [1,45] Hz, 'band')
//   normalize_data   (134-154): per-channel z-score, population std floored at 1e-10
//   create_sequences (157-180): windows of seq_len samples every `step` samples, transposed to (seq_len, C)
// scipy is a third-party dependency of the reference (`scipy>=1.11.0`, unpinned; 1.18.1 in the build container); its
// published algorithm, restated here: filtfilt pads both ends by `padlen` samples with the odd extension
// (2 x[0] - x[padlen..1], x, 2 x[-1] - x[-2..-padlen-1]), runs lfilter (direct form II transposed) forward with the
// initial state zi * ext[0] (zi = lfilter_zi(b, a): the steady state of a unit step), runs it again over the reversed
// output with zi * y[-1], reverses and trims the padding.
//
// Layout: one thread per (recording, channel) row walks its samples serially in fp64 (the recursion is inherently
// sequential per row; a batch of recordings supplies the parallelism: 60 subjects x 3 sessions x 2 tasks x 61 channels =
// 21 960 rows), loading 8 samples ahead so one L2 round trip is paid per 8 recursion steps.  The forward pass writes the
// extended signal to the workspace, the backward pass overwrites it in place and accumulates the row's sum / sum of
// squares; a second kernel normalises, transposes through shared memory and writes every sample into all windows that
// contain it (coalesced 4-byte stores, C contiguous floats per time step).
#include "common.cuh"

namespace bci {

constexpr int PP_MAX_ORDER = 16;

struct FiltCoef {
  double b[PP_MAX_ORDER + 1], a[PP_MAX_ORDER + 1], zi[PP_MAX_ORDER];
};

template <typename InT>
__device__ __forceinline__ double ext_sample(const InT* __restrict__ x, long long n, int p, long long i) {
  // odd extension of x[0..n) by p samples on each side, index i in [0, n + 2p)
  if (i < p) return __dsub_rn(__dmul_rn(2.0, (double)x[0]), (double)x[p - i]);
  if (i < p + n) return (double)x[i - p];
  return __dsub_rn(__dmul_rn(2.0, (double)x[n - 1]), (double)x[n - 2 - (i - p - n)]);
}

template <int ORD>
__device__ __forceinline__ double df2t_step(const FiltCoef& c, double (&z)[ORD], double x) {
  // scipy _linear_filter: y = z0 + b0 x;  z_j = z_{j+1} + b_{j+1} x - a_{j+1} y;  z_last = b_ORD x - a_ORD y
  // Unfused IEEE operations in scipy's evaluation order ((z_{j+1} + x b) - y a): the band-pass has poles next to the
  // unit circle (1 Hz corner at 500 Hz), so FMA contraction alone moves the output by 5e-8 of its scale; written this
  // way the recursion reproduces scipy's fp64 results.
  const double y = __dadd_rn(z[0], __dmul_rn(c.b[0], x));
#pragma unroll
  for (int j = 0; j < ORD - 1; ++j)
    z[j] = __dsub_rn(__dadd_rn(z[j + 1], __dmul_rn(x, c.b[j + 1])), __dmul_rn(y, c.a[j + 1]));
  z[ORD - 1] = __dsub_rn(__dmul_rn(x, c.b[ORD]), __dmul_rn(y, c.a[ORD]));
  return y;
}

template <typename InT, int ORD>
__global__ void __launch_bounds__(128)
filtfilt_rows_kernel(const InT* __restrict__ raw, long long n, int rows, int p, const FiltCoef c,
                     double* __restrict__ ybuf /* rows x (n + 2p) */, double* __restrict__ sums /* rows x 2 */) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const InT* x = raw + (long long)row * n;
  const long long m = n + 2 * (long long)p;
  double* y = ybuf + (long long)row * m;
  double z[ORD];
  // ---- forward over the extended signal ----
  const double x0 = ext_sample(x, n, p, 0);
#pragma unroll
  for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], x0);
  long long i = 0;
  for (; i < p; ++i) y[i] = df2t_step<ORD>(c, z, ext_sample(x, n, p, i));
  const long long mid_end = p + n;
  for (; i + 8 <= mid_end; i += 8) {
    double xv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xv[k] = (double)x[i - p + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[i + k] = df2t_step<ORD>(c, z, xv[k]);
  }
  for (; i < m; ++i) y[i] = df2t_step<ORD>(c, z, ext_sample(x, n, p, i));
  // ---- backward, in place ----
  const double yl = y[m - 1];
#pragma unroll
  for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], yl);
  double s = 0.0, ss = 0.0;
  i = m - 1;
  for (; i >= mid_end; --i) y[i] = df2t_step<ORD>(c, z, y[i]);
  for (; i - 7 >= p; i -= 8) {
    double yv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) yv[k] = y[i - k];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double o = df2t_step<ORD>(c, z, yv[k]);
      y[i - k] = o;
      s += o;
      ss = fma(o, o, ss);
    }
  }
  for (; i >= p; --i) {
    const double o = df2t_step<ORD>(c, z, y[i]);
    y[i] = o;
    s += o;
    ss = fma(o, o, ss);
  }
  // (the left padding is never read again)
  sums[2 * row] = s;
  sums[2 * row + 1] = ss;
}

// mean / std per row from the sums (np.mean, np.std ddof=0, std floored at 1e-10: 02:145-151) unless given
__global__ void rowstats_kernel(const double* __restrict__ sums, long long n, int rows, const double* __restrict__ mean_in,
                                const double* __restrict__ std_in, int C, double* __restrict__ mean_out, double* __restrict__ std_out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double mu, sd;
  if (mean_in) {
    mu = mean_in[row % C];
    sd = std_in[row % C];
  } else {
    mu = sums[2 * row] / (double)n;
    double var = sums[2 * row + 1] / (double)n - mu * mu;
    if (var < 0.0) var = 0.0;
    sd = sqrt(var);
    if (sd < 1e-10) sd = 1e-10;
  }
  mean_out[row] = mu;
  std_out[row] = sd;
}

constexpr int WIN_TILE = 128;

// grid = (sample tiles, recordings).  ybuf rows are (recording, channel) of length m = n + 2p, valid part at offset p.
__global__ void __launch_bounds__(256)
zscore_window_kernel(const double* __restrict__ ybuf, long long n, int p, int C, int seq_len, int step, long long n_seq,
                     const double* __restrict__ mean, const double* __restrict__ stdv, float* __restrict__ X /* (R*n_seq, seq_len, C) */,
                     double* __restrict__ filtered /* optional (R, C, n) */) {
  extern __shared__ float win_tile[];  // [WIN_TILE][C + 1]
  const int r = blockIdx.y;
  const long long k0 = (long long)blockIdx.x * WIN_TILE;
  const long long m = n + 2 * (long long)p;
  const int CS = C + 1;
  for (int e = threadIdx.x; e < C * WIN_TILE; e += blockDim.x) {
    const int c = e / WIN_TILE, k = e - c * WIN_TILE;
    const long long g = k0 + k;
    if (g < n) {
      const long long row = (long long)r * C + c;
      const double v = ybuf[row * m + p + g];
      if (filtered) filtered[row * n + g] = v;
      win_tile[k * CS + c] = (float)((v - mean[row]) / stdv[row]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * WIN_TILE; e += blockDim.x) {
    const int k = e / C, c = e - k * C;
    const long long g = k0 + k;
    if (g >= n) break;
    const float v = win_tile[k * CS + c];
    // windows w with w*step <= g < w*step + seq_len and w < n_seq
    long long w_hi = g / step;
    if (w_hi >= n_seq) w_hi = n_seq - 1;
    long long w_lo = (g - seq_len + step) / step;  // ceil((g - seq_len + 1) / step) for g - seq_len + 1 > 0
    if (g < seq_len) w_lo = 0;
    for (long long w = w_lo; w <= w_hi; ++w) {
      const long long t = g - w * step;
      if (t >= 0 && t < seq_len) X[(((long long)r * n_seq + w) * seq_len + t) * C + c] = v;
    }
  }
}

template <typename InT>
static int launch_filtfilt(const InT* raw, long long n, int rows, int p, const FiltCoef& c, int order, double* ybuf, double* sums,
                           cudaStream_t st) {
  const int blocks = ceil_div(rows, 128);
  switch (order) {
    case 2: filtfilt_rows_kernel<InT, 2><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    case 4: filtfilt_rows_kernel<InT, 4><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    case 6: filtfilt_rows_kernel<InT, 6><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    case 8: filtfilt_rows_kernel<InT, 8><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    case 12: filtfilt_rows_kernel<InT, 12><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    case 16: filtfilt_rows_kernel<InT, 16><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, ybuf, sums); break;
    default:
      set_error("bci_preprocess: filter order %d not built (2,4,6,8,12,16; butter(N,'band') has order 2N)", order);
      return BCI_EINVAL;
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci

using namespace bci;

extern "C" int bci_preprocess_workspace_bytes(const bci_preproc_args* a, size_t* bytes) {
  BCI_REQUIRE(a && bytes, BCI_EINVAL, "bci_preprocess_workspace_bytes: NULL argument");
  const size_t rows = (size_t)a->n_recordings * a->n_channels;
  *bytes = align_up(rows * (size_t)(a->n_samples + 2 * (int64_t)a->padlen) * 8, 256) + align_up(rows * 16, 256);
  return BCI_OK;
}

extern "C" int bci_preprocess(const bci_preproc_args* a, const void* raw, float* windows, double* mean_out, double* std_out,
                              double* filtered, void* workspace, size_t workspace_bytes, void* stream) {
  BCI_REQUIRE(a && raw && windows && mean_out && std_out && workspace, BCI_EINVAL, "bci_preprocess: NULL argument");
  BCI_REQUIRE(a->n_channels >= 1 && a->n_channels <= 256 && a->n_recordings >= 1, BCI_EINVAL, "bci_preprocess: bad channel/recording count");
  BCI_REQUIRE(a->order >= 1 && a->order <= PP_MAX_ORDER && a->b_host && a->a_host && a->zi_host, BCI_EINVAL, "bci_preprocess: bad filter");
  BCI_REQUIRE(a->padlen >= 0 && a->n_samples > a->padlen, BCI_EINVAL,
              "bci_preprocess: the input must be longer than padlen (%d) samples (scipy.filtfilt raises ValueError here)", a->padlen);
  BCI_REQUIRE(a->seq_len >= 1 && a->step >= 1 && a->n_samples >= a->seq_len, BCI_EINVAL, "bci_preprocess: bad window geometry");
  BCI_REQUIRE(a->in_dtype == BCI_OUT_F32 || a->in_dtype == BCI_OUT_F64, BCI_EINVAL, "bci_preprocess: in_dtype must be BCI_OUT_F32/F64");
  BCI_REQUIRE((a->mean_in == nullptr) == (a->std_in == nullptr), BCI_EINVAL, "bci_preprocess: mean_in and std_in go together");
  size_t need = 0;
  bci_preprocess_workspace_bytes(a, &need);
  BCI_REQUIRE(workspace_bytes >= need, BCI_ENOMEM, "bci_preprocess: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = a->n_recordings * a->n_channels;
  const long long n = a->n_samples, m = n + 2 * (long long)a->padlen;
  FiltCoef c;
  const double a0 = a->a_host[0];
  BCI_REQUIRE(a0 != 0.0, BCI_EINVAL, "bci_preprocess: a[0] == 0");
  for (int i = 0; i <= PP_MAX_ORDER; ++i) {
    c.b[i] = i <= a->order ? a->b_host[i] / a0 : 0.0;
    c.a[i] = i <= a->order ? a->a_host[i] / a0 : 0.0;
  }
  for (int i = 0; i < PP_MAX_ORDER; ++i) c.zi[i] = i < a->order ? a->zi_host[i] : 0.0;
  double* ybuf = reinterpret_cast<double*>(workspace);
  double* sums = reinterpret_cast<double*>((char*)workspace + align_up((size_t)rows * m * 8, 256));
  int rc = a->in_dtype == BCI_OUT_F64 ? launch_filtfilt<double>((const double*)raw, n, rows, a->padlen, c, a->order, ybuf, sums, st)
                                      : launch_filtfilt<float>((const float*)raw, n, rows, a->padlen, c, a->order, ybuf, sums, st);
  if (rc) return rc;
  rowstats_kernel<<<ceil_div(rows, 128), 128, 0, st>>>(sums, n, rows, a->mean_in, a->std_in, a->n_channels, mean_out, std_out);
  BCI_LAUNCH_OK();
  const long long n_seq = (n - a->seq_len) / a->step + 1;
  const size_t smem = (size_t)WIN_TILE * (a->n_channels + 1) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    BCI_CUDA_OK(cudaFuncSetAttribute(zscore_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)ceil_div64(n, WIN_TILE), (unsigned)a->n_recordings);
  zscore_window_kernel<<<grid, 256, smem, st>>>(ybuf, n, a->padlen, a->n_channels, a->seq_len, a->step, n_seq, mean_out, std_out,
                                                windows, filtered);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
