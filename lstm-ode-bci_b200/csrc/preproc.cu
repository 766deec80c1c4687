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
// Layout: one thread per (recording, channel, chunk of 8 192 samples) walks its samples serially in fp64 with 8 192 samples of
// warm-up (see filtfilt_tile_kernel, the warp-cooperative default, and filtfilt_chunk_kernel, the first form); the forward pass writes the extended signal to the workspace, the backward pass reads
// it and writes a second buffer plus per-chunk sum / sum of squares; a last kernel normalises, transposes through shared
// memory and writes every sample into all windows that contain it (coalesced 4-byte stores, C contiguous floats per step).
// Bound: the FP64 pipe (33 unfused operations per sample and pass); bci_fp64_peak_probe measures its peak.
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

// Time-parallel form of the recursion.  A row is cut into chunks of PP_CHUNK output samples; every chunk is one thread that
// first runs PP_WARM samples of warm-up from a zero state (outputs discarded), except where the chunk reaches the start of the
// pass, which begins exactly as scipy does (state zi * first sample).  The filter's slowest pole (Butterworth band-pass, 1 Hz
// corner at 500 Hz) has |p| = 0.9952 per sample, so after 8192 samples the influence of the unknown state has decayed by
// e^-39.  What a chunked evaluation cannot reproduce is scipy's ROUNDING: this direct-form recursion amplifies rounding differences to
// ~2e-7 of the output scale (scipy's own lfilter moves by 2.5e-7 under a 1e-15 perturbation of its initial state), so rows of one chunk
// are bit-identical to scipy and longer rows agree to 1.5-2.1e-7 of the scale (tests/test_gpu_preproc.py), while
// a batch of R recordings exposes R x 61 x n/PP_CHUNK threads instead of R x 61 (the fp64 pipe issues one warp instruction per
// 8 cycles per SM sub-partition on B200; a single warp per 32 rows left 130 of 148 SMs idle).
// (Round 2 tried 4096-sample chunks with 6144 of warm-up -- 4x the threads for 1.67x the arithmetic: 14 recordings per call ran 6 %
// faster, 36 recordings 34 % slower (15.3 instead of 11.4 ms): from ~80 k threads on the kernel is no longer latency-bound.  Kept as is.)
// (Session 5: with the warp-cooperative kernel below the passes are no longer bound by their scattered memory requests, and more, shorter
// chunks now pay: 8192-sample chunks -- 2x the warps for 1.33x the arithmetic -- take 7.30 instead of 7.62 ms for 36 recordings and 3.65
// instead of 5.30 ms for the 14-recording batches of the config-5 pipeline; 4096-sample chunks 8.31 / 3.68 ms.)
#ifndef BCI_PP_CHUNK
#define BCI_PP_CHUNK 8192
#endif
constexpr int PP_CHUNK = BCI_PP_CHUNK, PP_WARM = 8192;
// (Session 5, ncu source view: 60 % of the forward pass's samples sit on the use of the batch loaded one batch earlier, 64 % in the
// backward pass -- the 32 lanes of a warp walk 32 rows 0.6-1.2 MB apart and the passes stream 6.5 GB in all, so a load costs more than
// the ~600 cycles of recursion it is hidden behind.  An L2 prefetch 96 samples ahead made both passes SLOWER (4.0 -> 4.5, 4.5 -> 6.3 ms):
// the fetch itself is not what the warps wait for.  Left as measured.)

// pass = 0: forward over the odd-extended input -> yf (rows x m);  pass = 1: backward over yf -> yb (rows x m), plus per-chunk
// (sum, sumsq) of the kept samples [p, p+n)
template <typename InT, int ORD, int PASS>
__global__ void __launch_bounds__(128)
filtfilt_chunk_kernel(const InT* __restrict__ raw, long long n, int rows, int p, const FiltCoef c, const double* __restrict__ yf,
                      double* __restrict__ yout, int chunks, double* __restrict__ partial /* rows x chunks x 2 */) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)rows * chunks) return;
  // consecutive threads = consecutive rows of the same chunk index: the 32 lanes of a warp walk 32 different rows in step
  // (the other mapping -- a warp = the chunks of 3-4 rows, i.e. 4-8 pages per warp access instead of 32 -- measured 9 % slower)
  const int ck = (int)(gid / rows), row = (int)(gid - (long long)ck * rows);
  const long long m = n + 2 * (long long)p;
  const InT* x = raw + (long long)row * n;
  const double* src = yf + (long long)row * m;
  double* dst = yout + (long long)row * m;
  double z[ORD];
  if (PASS == 0) {
    const long long s0 = (long long)ck * PP_CHUNK, e0 = (s0 + PP_CHUNK < m) ? s0 + PP_CHUNK : m;
    long long i = s0 - PP_WARM;
    if (i <= 0) {
      i = 0;
      const double x0 = ext_sample(x, n, p, 0);
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], x0);
    } else {
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = 0.0;
    }
    // edges of the extended signal go through ext_sample one by one; the interior [p, p+n) is read in batches of 8, the next
    // batch loaded before the current one is consumed, so one memory round trip is paid per 8 recursion steps (every lane walks
    // its own row: the loads of a warp are 32 different sectors)
    auto run = [&](long long from, long long to, bool keep) {
      long long k = from;
      for (; k < to && k < p; ++k) { const double o = df2t_step<ORD>(c, z, ext_sample(x, n, p, k)); if (keep) dst[k] = o; }
      const long long in_end = (to < p + n) ? to : p + n;
      if (k + 8 <= in_end) {
        double cur[8], nxt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) cur[u] = (double)x[k - p + u];
        for (; k + 16 <= in_end; k += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) nxt[u] = (double)x[k - p + 8 + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const double o = df2t_step<ORD>(c, z, cur[u]); if (keep) dst[k + u] = o; }
#pragma unroll
          for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const double o = df2t_step<ORD>(c, z, cur[u]); if (keep) dst[k + u] = o; }
        k += 8;
      }
      for (; k < to; ++k) { const double o = df2t_step<ORD>(c, z, ext_sample(x, n, p, k)); if (keep) dst[k] = o; }
    };
    run(i, s0, false);
    run(s0, e0, true);
  } else {
    // backward: chunk ck covers [s0, e0) counted from the END of the extended signal
    const long long hi = m - (long long)ck * PP_CHUNK;              // exclusive upper index
    const long long lo = (hi - PP_CHUNK > 0) ? hi - PP_CHUNK : 0;   // inclusive lower index
    long long i = hi - 1 + PP_WARM;
    if (i >= m - 1) {
      i = m - 1;
      const double yl = src[m - 1];
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], yl);
    } else {
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = 0.0;
    }
    double sm = 0.0, ss = 0.0;
    // descending batches of 8 with the next batch in flight (see the forward pass)
    auto run = [&](long long from, long long to, bool keep) {  // processes from, from-1, ..., to (inclusive), from >= to
      long long k = from;
      if (k - 7 >= to) {
        double cur[8], nxt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) cur[u] = src[k - u];
        for (; k - 15 >= to; k -= 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) nxt[u] = src[k - 8 - u];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double o = df2t_step<ORD>(c, z, cur[u]);
            if (keep) { dst[k - u] = o; if (k - u >= p && k - u < p + n) { sm += o; ss = fma(o, o, ss); } }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double o = df2t_step<ORD>(c, z, cur[u]);
          if (keep) { dst[k - u] = o; if (k - u >= p && k - u < p + n) { sm += o; ss = fma(o, o, ss); } }
        }
        k -= 8;
      }
      for (; k >= to; --k) {
        const double o = df2t_step<ORD>(c, z, src[k]);
        if (keep) { dst[k] = o; if (k >= p && k < p + n) { sm += o; ss = fma(o, o, ss); } }
      }
    };
    if (i >= hi) run(i, hi, false);
    if (hi - 1 >= lo) run(hi - 1, lo, true);
    partial[((long long)row * chunks + ck) * 2] = sm;
    partial[((long long)row * chunks + ck) * 2 + 1] = ss;
  }
}

// Warp-cooperative form of the same passes (session 5; the default).  ncu on filtfilt_chunk_kernel (profiles/r5_preproc.md): every
// load / store instruction of a warp touches 32 sectors in 32 different rows; the warps wait on those requests (long scoreboard 3.8 of
// 6.0 cycles per issued instruction), and putting MORE of them in flight (L2 prefetch, a third register batch) made the passes slower --
// the memory pipeline is bound by the number of scattered requests, not by latency or bytes.  Here the 32 lanes of a warp still own 32
// rows for the serial recursion, but global memory is only touched a TILE at a time: 32 consecutive samples of each of the 32 rows,
// loaded with lane = sample (one contiguous 128- / 256-byte request per row), transposed through shared memory (row stride 33), consumed
// and overwritten in place by the owning lane, and written back the same way.  The next tile's loads are issued before the current tile
// is computed (~5 000 cycles of recursion).  Arithmetic per row is unchanged (same operations, same order): results are bit-identical.
constexpr int PP_TILE = 32;
constexpr int PP_TW = 4;   // warps per block

template <typename InT, int ORD, int PASS>
__global__ void __launch_bounds__(PP_TW * 32)
filtfilt_tile_kernel(const InT* __restrict__ raw, long long n, int rows, int p, const FiltCoef c, const double* __restrict__ yf,
                     double* __restrict__ yout, int chunks, double* __restrict__ partial /* rows x chunks x 2 */) {
  __shared__ double tile_s[PP_TW][PP_TILE][PP_TILE + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warps_per_ck = (rows + 31) / 32;
  const long long wg = (long long)blockIdx.x * PP_TW + wib;
  if (wg >= (long long)warps_per_ck * chunks) return;   // warp-uniform
  const int ck = (int)(wg / warps_per_ck), row0 = (int)(wg - (long long)ck * warps_per_ck) * 32;
  const int row = row0 + lane;
  const bool valid = row < rows;
  const int rowc = valid ? row : rows - 1;               // lanes past the last row shadow it and store nothing
  const long long m = n + 2 * (long long)p;
  const InT* x = raw + (long long)rowc * n;
  const double* src = yf + (long long)rowc * m;
  double* dst = yout + (long long)rowc * m;
  double (*tl)[PP_TILE + 1] = tile_s[wib];
  double z[ORD];
  auto row_of = [&](int rr) { const int r2 = row0 + rr; return r2 < rows ? r2 : rows - 1; };

  if (PASS == 0) {
    const long long s0 = (long long)ck * PP_CHUNK, e0 = (s0 + PP_CHUNK < m) ? s0 + PP_CHUNK : m;
    long long i = s0 - PP_WARM;
    if (i <= 0) {
      i = 0;
      const double x0 = ext_sample(x, n, p, 0);
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], x0);
    } else {
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = 0.0;
    }
    const long long in_end = (e0 < p + n) ? e0 : p + n;
    InT nxt[PP_TILE];
    auto load_tile = [&](long long kt) {     // lane = sample kt + lane of row rr
#pragma unroll
      for (int rr = 0; rr < PP_TILE; ++rr) nxt[rr] = raw[(long long)row_of(rr) * n + (kt - p) + lane];
    };
    auto stage_tile = [&]() {
#pragma unroll
      for (int rr = 0; rr < PP_TILE; ++rr) tl[rr][lane] = (double)nxt[rr];
    };
    bool have = false;
    for (long long k = i; k < e0; k += PP_TILE) {
      const bool keep = k >= s0;
      const bool interior = k >= p && k + PP_TILE <= in_end;
      if (!interior) {   // the odd-extended edges and the last partial tile: per-lane, sample by sample
        const long long kb = (k + PP_TILE < e0) ? k + PP_TILE : e0;
        for (long long kk = k; kk < kb; ++kk) {
          const double o = df2t_step<ORD>(c, z, ext_sample(x, n, p, kk));
          if (keep && valid) dst[kk] = o;
        }
        have = false;
        continue;
      }
      if (!have) { load_tile(k); stage_tile(); __syncwarp(); }
      const bool next_interior = k + PP_TILE >= p && k + 2 * PP_TILE <= in_end && k + PP_TILE < e0;
      if (next_interior) load_tile(k + PP_TILE);
#pragma unroll 8
      for (int j = 0; j < PP_TILE; ++j) tl[lane][j] = df2t_step<ORD>(c, z, tl[lane][j]);
      __syncwarp();
      if (keep) {
#pragma unroll 8
        for (int rr = 0; rr < PP_TILE; ++rr)
          if (row0 + rr < rows) yout[(long long)(row0 + rr) * m + k + lane] = tl[rr][lane];
      }
      __syncwarp();
      if (next_interior) { stage_tile(); __syncwarp(); }
      have = next_interior;
    }
  } else {
    // backward: chunk ck covers [lo, hi) counted from the END of the extended signal; descending tiles [k - 31, k]
    const long long hi = m - (long long)ck * PP_CHUNK;              // exclusive upper index
    const long long lo = (hi - PP_CHUNK > 0) ? hi - PP_CHUNK : 0;   // inclusive lower index
    long long i = hi - 1 + PP_WARM;
    if (i >= m - 1) {
      i = m - 1;
      const double yl = src[m - 1];
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = __dmul_rn(c.zi[j], yl);
    } else {
#pragma unroll
      for (int j = 0; j < ORD; ++j) z[j] = 0.0;
    }
    double sm = 0.0, ss = 0.0;
    double nxt[PP_TILE];
    auto load_tile = [&](long long kt) {     // tile [kt - 31, kt]: lane = sample kt - 31 + lane of row rr
#pragma unroll
      for (int rr = 0; rr < PP_TILE; ++rr) nxt[rr] = yf[(long long)row_of(rr) * m + (kt - (PP_TILE - 1)) + lane];
    };
    auto stage_tile = [&]() {
#pragma unroll
      for (int rr = 0; rr < PP_TILE; ++rr) tl[rr][lane] = nxt[rr];
    };
    bool have = false;
    for (long long k = i; k >= lo; k -= PP_TILE) {
      const bool keep = k <= hi - 1;
      const bool full = k - (PP_TILE - 1) >= lo;
      if (!full) {
        for (long long kk = k; kk >= lo; --kk) {
          const double o = df2t_step<ORD>(c, z, src[kk]);
          if (keep) { if (valid) dst[kk] = o; if (kk >= p && kk < p + n) { sm += o; ss = fma(o, o, ss); } }
        }
        break;
      }
      if (!have) { load_tile(k); stage_tile(); __syncwarp(); }
      const bool next_full = k - (2 * PP_TILE - 1) >= lo;
      if (next_full) load_tile(k - PP_TILE);
#pragma unroll 8
      for (int j = PP_TILE - 1; j >= 0; --j) {
        const double o = df2t_step<ORD>(c, z, tl[lane][j]);
        tl[lane][j] = o;
        const long long kk = k - (PP_TILE - 1) + j;
        if (keep && kk >= p && kk < p + n) { sm += o; ss = fma(o, o, ss); }
      }
      __syncwarp();
      if (keep) {
#pragma unroll 8
        for (int rr = 0; rr < PP_TILE; ++rr)
          if (row0 + rr < rows) yout[(long long)(row0 + rr) * m + (k - (PP_TILE - 1)) + lane] = tl[rr][lane];
      }
      __syncwarp();
      if (next_full) { stage_tile(); __syncwarp(); }
      have = next_full;
    }
    if (valid) {
      partial[((long long)row * chunks + ck) * 2] = sm;
      partial[((long long)row * chunks + ck) * 2 + 1] = ss;
    }
  }
}

// mean / std per row from the sums (np.mean, np.std ddof=0, std floored at 1e-10: 02:145-151) unless given
__global__ void rowstats_kernel(const double* __restrict__ partial, int chunks, long long n, int rows, const double* __restrict__ mean_in,
                                const double* __restrict__ std_in, int C, double* __restrict__ mean_out, double* __restrict__ std_out) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double mu, sd;
  if (mean_in) {
    mu = mean_in[row % C];
    sd = std_in[row % C];
  } else {
    double s1 = 0.0, s2 = 0.0;  // fixed order: deterministic
    for (int k = 0; k < chunks; ++k) { s1 += partial[((long long)row * chunks + k) * 2]; s2 += partial[((long long)row * chunks + k) * 2 + 1]; }
    mu = s1 / (double)n;
    double var = s2 / (double)n - mu * mu;
    if (var < 0.0) var = 0.0;
    sd = sqrt(var);
    if (sd < 1e-10) sd = 1e-10;
  }
  mean_out[row] = mu;
  std_out[row] = sd;
}

constexpr int WIN_TILE = 128;

// grid = (sample tiles, recordings).  ybuf rows are (recording, channel) of length m = n + 2p, valid part at offset p.
// ncu on the first version (profiles/r5_preproc.md): issue slots 79 % busy at 23 % of DRAM peak -- ~240 instructions per sample, most
// of them 64-bit and runtime-divisor integer divisions of the index arithmetic.  Here the load phase indexes with shifts (a warp =
// 32 consecutive samples of one channel: coalesced fp64 reads, mean / std broadcast), the store phase maps lanes to channels
// (c = tid % 64, four samples per pass: no division by C), and the window range of a sample costs two 32-bit divisions.
__global__ void __launch_bounds__(256)
zscore_window_kernel(const double* __restrict__ ybuf, int n, int p, int C, int seq_len, int step, int n_seq,
                     const double* __restrict__ mean, const double* __restrict__ stdv, float* __restrict__ X /* (R*n_seq, seq_len, C) */,
                     double* __restrict__ filtered /* optional (R, C, n) */) {
  extern __shared__ float win_tile[];  // [WIN_TILE][CS], CS odd (<= C + 1): the transposing writes of a warp hit 32 different banks
  __shared__ int w_first[WIN_TILE], w_last[WIN_TILE];   // windows containing sample k0 + k (two 32-bit divisions per SAMPLE, not per element)
  const int r = blockIdx.y;
  const int k0 = blockIdx.x * WIN_TILE;
  const long long m = (long long)n + 2 * (long long)p;
  const int CS = C | 1;
  if (threadIdx.x < WIN_TILE) {
    const int g = k0 + threadIdx.x;
    // windows w with w*step <= g < w*step + seq_len and w < n_seq
    int w_hi = g / step;
    if (w_hi >= n_seq) w_hi = n_seq - 1;
    w_first[threadIdx.x] = g < seq_len ? 0 : (g - seq_len + step) / step;   // ceil((g - seq_len + 1) / step)
    w_last[threadIdx.x] = w_hi;
  }
  for (int e = threadIdx.x; e < C * WIN_TILE; e += blockDim.x) {
    const int c = e >> 7, k = e & (WIN_TILE - 1);   // a warp = 32 consecutive samples of one channel: coalesced reads, mean / std broadcast
    const int g = k0 + k;
    if (g < n) {
      const long long row = (long long)r * C + c;
      const double v = ybuf[row * m + p + g];
      if (filtered) filtered[row * (long long)n + g] = v;
      win_tile[k * CS + c] = (float)((v - mean[row]) / stdv[row]);
    }
  }
  __syncthreads();
  const int c0 = threadIdx.x & 63;
  for (int k = threadIdx.x >> 6; k < WIN_TILE; k += 4) {
    const int g = k0 + k;
    if (g >= n) break;
    const int w_hi = w_last[k];
    for (int w = w_first[k]; w <= w_hi; ++w) {
      const int t = g - w * step;
      if (t < 0 || t >= seq_len) continue;
      float* dst = X + (((long long)r * n_seq + w) * seq_len + t) * C;
      for (int c = c0; c < C; c += 64) dst[c] = win_tile[k * CS + c];
    }
  }
}

template <typename InT, int ORD>
static int launch_filtfilt_ord(const InT* raw, long long n, int rows, int p, const FiltCoef& c, double* yf, double* yb, int chunks,
                               double* partial, cudaStream_t st) {
  static const bool lanes_form = [] { const char* e = getenv("BCI_PP_FILTER"); return e && e[0] == 'l'; }();   // BCI_PP_FILTER=lanes: the first form
  if (lanes_form) {
    const long long threads = (long long)rows * chunks;
    const unsigned blocks = (unsigned)ceil_div64(threads, 128);
    filtfilt_chunk_kernel<InT, ORD, 0><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, nullptr, yf, chunks, partial);
    BCI_LAUNCH_OK();
    filtfilt_chunk_kernel<InT, ORD, 1><<<blocks, 128, 0, st>>>(raw, n, rows, p, c, yf, yb, chunks, partial);
    BCI_LAUNCH_OK();
    return BCI_OK;
  }
  const long long warps = (long long)((rows + 31) / 32) * chunks;
  const unsigned blocks = (unsigned)ceil_div64(warps, PP_TW);
  filtfilt_tile_kernel<InT, ORD, 0><<<blocks, PP_TW * 32, 0, st>>>(raw, n, rows, p, c, nullptr, yf, chunks, partial);
  BCI_LAUNCH_OK();
  filtfilt_tile_kernel<InT, ORD, 1><<<blocks, PP_TW * 32, 0, st>>>(raw, n, rows, p, c, yf, yb, chunks, partial);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

template <typename InT>
static int launch_filtfilt(const InT* raw, long long n, int rows, int p, const FiltCoef& c, int order, double* yf, double* yb, int chunks,
                           double* partial, cudaStream_t st) {
  switch (order) {
    case 2: return launch_filtfilt_ord<InT, 2>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    case 4: return launch_filtfilt_ord<InT, 4>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    case 6: return launch_filtfilt_ord<InT, 6>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    case 8: return launch_filtfilt_ord<InT, 8>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    case 12: return launch_filtfilt_ord<InT, 12>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    case 16: return launch_filtfilt_ord<InT, 16>(raw, n, rows, p, c, yf, yb, chunks, partial, st);
    default:
      set_error("bci_preprocess: filter order %d not built (2,4,6,8,12,16; butter(N,'band') has order 2N)", order);
      return BCI_EINVAL;
  }
}

static inline int pp_chunks(long long m) { return (int)((m + PP_CHUNK - 1) / PP_CHUNK); }

// dependent-free DFMA stream: the FP64 pipe's peak, the roofline denominator of the filter kernels
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
  const double mm = 0.999, cc = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, mm, cc); a1 = fma(a1, mm, cc); a2 = fma(a2, mm, cc); a3 = fma(a3, mm, cc);
      a4 = fma(a4, mm, cc); a5 = fma(a5, mm, cc); a6 = fma(a6, mm, cc); a7 = fma(a7, mm, cc);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace bci

using namespace bci;

extern "C" int bci_preprocess_workspace_bytes(const bci_preproc_args* a, size_t* bytes) {
  BCI_REQUIRE(a && bytes, BCI_EINVAL, "bci_preprocess_workspace_bytes: NULL argument");
  const size_t rows = (size_t)a->n_recordings * a->n_channels;
  const size_t m = (size_t)(a->n_samples + 2 * (int64_t)a->padlen);
  // forward output, backward output (the passes are chunk-parallel, so the backward pass cannot work in place), chunk partials
  *bytes = 2 * align_up(rows * m * 8, 256) + align_up(rows * (size_t)pp_chunks((long long)m) * 16, 256);
  return BCI_OK;
}

extern "C" int bci_preprocess(const bci_preproc_args* a, const void* raw, float* windows, double* mean_out, double* std_out,
                              double* filtered, void* workspace, size_t workspace_bytes, void* stream) {
  bci::NvtxRange nvtx_range("bci_preprocess");
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
  const int chunks = pp_chunks(m);
  double* yf = reinterpret_cast<double*>(workspace);
  double* ybuf = reinterpret_cast<double*>((char*)workspace + align_up((size_t)rows * m * 8, 256));
  double* sums = reinterpret_cast<double*>((char*)workspace + 2 * align_up((size_t)rows * m * 8, 256));
  int rc = a->in_dtype == BCI_OUT_F64 ? launch_filtfilt<double>((const double*)raw, n, rows, a->padlen, c, a->order, yf, ybuf, chunks, sums, st)
                                      : launch_filtfilt<float>((const float*)raw, n, rows, a->padlen, c, a->order, yf, ybuf, chunks, sums, st);
  if (rc) return rc;
  rowstats_kernel<<<ceil_div(rows, 128), 128, 0, st>>>(sums, chunks, n, rows, a->mean_in, a->std_in, a->n_channels, mean_out, std_out);
  BCI_LAUNCH_OK();
  BCI_REQUIRE(m < (1ll << 31) - WIN_TILE, BCI_EINVAL, "bci_preprocess: recordings of more than 2^31 samples are not supported");
  const long long n_seq = (n - a->seq_len) / a->step + 1;
  const size_t smem = (size_t)WIN_TILE * (a->n_channels + 1) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    BCI_CUDA_OK(cudaFuncSetAttribute(zscore_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)ceil_div64(n, WIN_TILE), (unsigned)a->n_recordings);
  zscore_window_kernel<<<grid, 256, smem, st>>>(ybuf, (int)n, a->padlen, a->n_channels, a->seq_len, a->step, (int)n_seq, mean_out, std_out,
                                                windows, filtered);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

extern "C" int bci_fp64_peak_probe(double* tflops, void* stream) {
  BCI_REQUIRE(tflops, BCI_EINVAL, "bci_fp64_peak_probe: NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = sm_count() * 8, threads = 256, iters = 2048;
  double* buf = nullptr;
  BCI_CUDA_OK(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, st);
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(buf, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  BCI_CUDA_OK(cudaGetLastError());
  *tflops = best;
  return BCI_OK;
}
