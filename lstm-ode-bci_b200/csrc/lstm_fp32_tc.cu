// fp32-parity recurrence on the tensor cores: h_{t-1} . W_hh^T as a SPLIT-PRECISION fp16 product (3 x fp16 MMAs, fp32-grade).
//
// The fp32 mode must stay within 1e-5 of the reference's fp32 CPU path on the logits (north_star), which rules out one bf16 /
// tf32 MMA per step over 3 x 256 dependent steps.  The first version therefore ran the recurrent product on the CUDA cores
// (lstm_rec_f32: FFMA pipe 40 %, tensor pipe 0 %, 96 k windows/s).  Here every fp32 operand is split into two fp16 numbers,
//
//     h = h_hi + h_lo,   h_hi = fp16(h),      h_lo = fp16(h - h_hi)          |h| < 1: 22 significant bits, abs. error <= 3e-8
//     w = (w_hi + w_lo) / 16,  w_hi = fp16(16 w), w_lo = fp16(16 w - w_hi)   (the factor keeps w_lo out of fp16's subnormals)
//
// and  h . w^T  =  (h_lo . w_hi + h_hi . w_lo + h_hi . w_hi) / 16  + O(2^-22)  is three chains of tcgen05.mma.kind::f16 (fp16 x fp16
// products are exact in the fp32 accumulator).  The two small terms are accumulated FIRST: the tensor core adds into TMEM with
// truncation, an error that grows with the accumulator's magnitude (gemm_tf32x3.cu), and this way only the last 8 of the 24
// accumulation steps run at full magnitude.  fp16 rather than tf32 because the operands then take half the shared memory:
//
//   W_hh of one direction, hi + lo     512 x 128 x 2 B x 2 = 256 KB  -> resident across a CTA PAIR (128 KB each)
//   h_{t-1} of a 128-window tile       128 x 128 x 2 B x 2 =  64 KB  per CTA (the MMA's A operand, written by the epilogue)
//
//   cluster = 2 CTAs, one tcgen05.mma.cta_group::2 (M 256 x N 256 x K 16) per K slice: each CTA supplies the A rows of its OWN 128
//   windows and half of the B rows, and receives the accumulators of its own windows (128 lanes x 512 columns: all of TMEM).
//   So the pair shares the weights but no CTA ever needs the other's h: unlike the bf16 cluster kernel there is NO h exchange,
//   only "my h is written / my accumulator is drained" from the peer to the leader before the next step's MMAs are issued.
//
//   per step:  leader thread: 48 MMAs (2 column blocks x 3 terms x 8 K slices) -> commit (multicast to both CTAs)
//              8 epilogue warps per CTA: thread = (window, 64 hidden units): tcgen05.ld 32 columns = i,f,g,o of 8 units (the B rows
//              are ordered unit*4 + gate, the column order of G), + G_t (x . W_ih^T + b from the 3xTF32 GEMM), gates (ex2-based
//              sigma/tanh as lstm_rec_f32), cell state in fp32 registers, h_t -> out (fp32, 32-byte stores) and -> fp16 hi/lo
//              into the swizzled A buffers
//   G_t        row-major fp32, 128 B per (thread, slab).  Read per thread (lane = window row, 4 KB apart) every warp-level load
//              touched 32 different lines for 16 B each: ncu showed the epilogue stalled on those loads for 12 of the 22 us per
//              step (long-scoreboard 7.3 per issue; tensor pipe 15 %, XU 26 %).  Now a warp copies its 32 rows x 128 B slab with
//              fully coalesced 16-byte cp.async (8 lanes per row; no registers: a register-staged version stalled on the
//              write-after-read hazard between the next slab's loads and this slab's shared-memory stores) into a private 4 KB
//              swizzled buffer and reads it back row-wise, conflict-free; the copy of slab s+1 runs while the gates of slab s are
//              evaluated, and the next step's rows are pulled into L2 one step ahead by one bulk prefetch per thread.
//
// Reference semantics: nn.LSTM inside EnhancedLSTMModel (04_lstm_model.py:181-188,211), fp32 mode.
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int TC_M = 128;                  // windows per CTA
constexpr int TC_THREADS = 256;            // 8 epilogue warps; thread 0 is also the MMA issuer (leader) / relay (peer)
constexpr uint32_t TC_ATOM = 128 * 128;    // [128 rows][64 fp16] SW128 atom
constexpr uint32_t TC_OFF_H = 8 * TC_ATOM;     // W: [column block 2][part 2][K atom 2]
constexpr uint32_t TC_OFF_CTL = 12 * TC_ATOM;  // h: [part 2][K atom 2]
constexpr uint32_t TC_OFF_STAGE = TC_OFF_CTL + 256;   // per-warp 4 KB transpose buffers for the G slabs
constexpr size_t TC_SMEM = 1024 + TC_OFF_STAGE + 8 * 4096;
constexpr float TC_WSCALE = F16X3_WSCALE;

// instruction descriptor, kind::f16 with FP16 inputs (a_format = b_format = 0), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 16 bytes global -> shared without passing through registers (LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// one 32-byte store per thread (STG.256: sm_100 has 256-bit global accesses)
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// w_hh (4H, H) PyTorch layout (gate-major rows) -> dst [part 2][n' = unit*4 + gate][k] fp16, scaled by TC_WSCALE
__global__ void pack_whh_f16x3_kernel(const float* __restrict__ w, __half* __restrict__ dst, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H * H) return;
  const int n = i / H, k = i - n * H, unit = n >> 2, gate = n & 3;
  const float v = w[((size_t)gate * H + unit) * H + k] * TC_WSCALE;
  const __half hi = __float2half_rn(v);
  dst[i] = hi;
  dst[(size_t)4 * H * H + i] = __float2half_rn(v - __half2float(hi));
}

int pack_whh_f16x3(const float* w_hh, __half* dst, int H, cudaStream_t st) {
  pack_whh_f16x3_kernel<<<ceil_div(4 * H * H, 256), 256, 0, st>>>(w_hh, dst, H);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// work item w (0 .. ND * n_pairs): direction = w / n_pairs, window-tile pair = w % n_pairs; CTA `rank` owns windows
// [(2 pair + rank) * 128, +128)
template <bool SAVE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
lstm_rec_f16x3(const float* __restrict__ G,        // [T*Bc][ldg] fp32: column dir*512 + unit*4 + gate, bias included
               int ldg,
               const __half* __restrict__ whh16,   // [ND][2 parts][512][128]
               float* __restrict__ out,            // [T][Bc][D]: h_t at column dir*128 + unit (nullptr: not needed)
               __half* __restrict__ out_hi16,      // optional [T][Bc][D] fp16 pair (hi, lo) of h_t: the next layer's projection GEMM
               __half* __restrict__ out_lo16,      //   reads its A operand in this form (gemm_f16x3_nt) -- no separate split pass
               float* __restrict__ gates,          // SAVE: [T*Bc][ldg] gate ACTIVATIONS (i,f,g,o), same layout as G
               float* __restrict__ csave,          // SAVE: [T*Bc][D] cell states
               int D, int Bc, int T, int n_pairs, int ND, int pf_mode, int jitter) {  // jitter: 0 or a power of two (max sleep, ns)
  extern __shared__ uint8_t tc_smem_raw[];
  // BCI_FUSED_JITTER (tests only): every thread sleeps a pseudo-random time at its synchronisation points, to shake out ordering
  // assumptions of the pair protocol that only hold at the natural timing (compute-sanitizer is not available on the GPU pool)
  uint32_t jit_state = jitter ? (uint32_t)(blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) : 0u;
  auto jit = [&]() {
    if (jitter) {
      jit_state = jit_state * 1664525u + 1013904223u;
      __nanosleep((jit_state >> 20) & (uint32_t)(jitter - 1));
    }
  };
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = tc_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + TC_OFF_H;
  uint8_t* genH = gen + TC_OFF_H;
  uint8_t* ctl = gen + TC_OFF_CTL;
  const uint32_t bar0 = smem_u32(ctl);
  const uint32_t acc_full = bar0;          // every CTA: multicast commit -- this step's accumulators are complete
  const uint32_t h_local = bar0 + 8;       // every CTA: its 8 epilogue warps wrote h_t and drained TMEM
  const uint32_t peer_local = bar0 + 16;   // leader: relay of the peer's h_local
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (tid == 0) {
    mbar_init(acc_full, 1);
    mbar_init(h_local, TC_THREADS / 32);
    mbar_init(peer_local, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals them

  // all MMAs of one step: column block nb accumulates lo.hi + hi.lo (small terms first) and then hi.hi
  auto issue_step = [&]() {
    constexpr uint32_t idesc = umma_idesc_f16(256, 256);
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t ap = term == 0 ? 1u : 0u;   // part of h: lo, hi, hi
        const uint32_t bp = term == 1 ? 1u : 0u;   // part of W: hi, lo, hi
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t atom = k >> 2, kk = k & 3;
          const uint64_t da = umma_desc_sw128(sH + (ap * 2 + atom) * TC_ATOM + kk * 32);
          const uint64_t db = umma_desc_sw128(sW + ((nb * 2 + bp) * 2 + atom) * TC_ATOM + kk * 32);
          umma_bf16_2sm(tmem_base + nb * 256, da, db, idesc, (term | k) != 0 ? 1u : 0u);
        }
      }
    }
    umma_commit_2sm_mc(acc_full, (uint16_t)3);
  };

  const int n_work = ND * n_pairs;
  const int n_clusters = (int)cluster_nclusters_x();
  const int quarter = warp & 3, hf = warp >> 2;
  const int r = quarter * 32 + lane;   // window row of the tile == TMEM lane
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)hf * 256;
  const uint32_t peer_local_on_leader = mapa_u32(peer_local, 0);

  int g0 = 0;  // running step counter of this cluster (mbarrier parities are functions of it)
  for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
    const int dir = w / n_pairs, tp = w - dir * n_pairs;
    const int b0 = (2 * tp + (int)rank) * TC_M;
    if (g0 > 0) cluster_sync_all();  // every thread of both CTAs finished the previous item
    {
      // this CTA's weight rows: n' in [nb*256 + rank*128, +128) for both column blocks and both parts; h_{-1} = 0
      const uint4* src = reinterpret_cast<const uint4*>(whh16 + (size_t)dir * 2 * 512 * 128);
      for (int i = tid; i < 4 * 128 * 16; i += TC_THREADS) {
        const int blk = i >> 11, rem = i & 2047, row = rem >> 4, cc = rem & 15;   // blk = nb*2 + part
        const int nb = blk >> 1, part = blk & 1;
        const uint4 v = __ldg(src + ((size_t)part * 512 + nb * 256 + rank * 128 + row) * 16 + cc);
        *reinterpret_cast<uint4*>(gen + (blk * 2 + (cc >> 3)) * TC_ATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) = v;
      }
      for (int i = tid; i < (int)(4 * TC_ATOM / 16); i += TC_THREADS) reinterpret_cast<uint4*>(genH)[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_all();
    __syncthreads();
    cluster_sync_all();
    if (leader && tid == 0) issue_step();   // step 0: h_{-1} = 0 (one all-zero product of 256 keeps every step identical)

    const bool live = b0 + r < Bc;
    const int brow = live ? b0 + r : Bc - 1;   // dead rows read a valid row and store nothing
    // coalesced G loads: load j of a slab covers rows 4 j + lane / 8 of this warp's 32 rows, 16-byte chunk lane % 8
    uint32_t goff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = b0 + quarter * 32 + 4 * j + (lane >> 3);
      goff[j] = (uint32_t)(rr < Bc ? rr : Bc - 1) * (uint32_t)ldg + (uint32_t)(lane & 7) * 4u;
    }
    uint8_t* stg = gen + TC_OFF_STAGE + warp * 4096;
    const uint32_t stg_s = base + TC_OFF_STAGE + warp * 4096;
    float c[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) c[i] = 0.f;

    for (int st = 0; st < T; ++st) {
      const int g = g0 + st;
      const int t = dir ? (T - 1 - st) : st;
      const long long row = (long long)t * Bc + brow;
      const float* gstep = G + (long long)t * Bc * ldg + dir * 512 + hf * 256;
      const int tn = dir ? (T - 2 - st) : st + 1;
      const float* gnext = G + ((long long)tn * Bc + brow) * ldg + dir * 512 + hf * 256;   // this thread's row of the next step
      if (pf_mode == 1 && st + 1 < T) bulk_prefetch_l2(gnext, 1024u);  // 1 KB -> L2, one bulk prefetch (TMA engine, not the LSU)
      // slab 0 of this step's G rows -> the warp's staging buffer (row rr, chunk ch at rr * 128 + ((ch ^ (rr & 7)) << 4))
      auto copy_slab = [&](int sl) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t rr = 4 * j + (lane >> 3);
          cp_async16(stg_s + rr * 128 + ((((uint32_t)lane & 7u) ^ (rr & 7u)) << 4), gstep + goff[j] + sl * 32);
        }
        cp_async_commit();
      };
      copy_slab(0);
      if (lane == 0) jit();
      __syncwarp();
      mbar_wait(acc_full, (uint32_t)(g & 1));
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld32(taddr, acc);
      float* orow = out + row * D + dir * 128 + hf * 64;
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        cp_async_wait_all();
        __syncwarp();   // every lane's part of the slab has landed
        tmem_ld_wait();
        float pre[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 gv = *reinterpret_cast<const uint4*>(stg + lane * 128 + (((uint32_t)i ^ ((uint32_t)lane & 7u)) << 4));
          pre[4 * i + 0] = fmaf(__uint_as_float(acc[4 * i + 0]), 1.0f / TC_WSCALE, __uint_as_float(gv.x));
          pre[4 * i + 1] = fmaf(__uint_as_float(acc[4 * i + 1]), 1.0f / TC_WSCALE, __uint_as_float(gv.y));
          pre[4 * i + 2] = fmaf(__uint_as_float(acc[4 * i + 2]), 1.0f / TC_WSCALE, __uint_as_float(gv.z));
          pre[4 * i + 3] = fmaf(__uint_as_float(acc[4 * i + 3]), 1.0f / TC_WSCALE, __uint_as_float(gv.w));
        }
        __syncwarp();  // every lane has read the buffer before the next slab overwrites it
        if (pf_mode == 2 && st + 1 < T) prefetch_l2(gnext + sl * 32);   // rolling: the same slab of the next step, one line per thread
        if (sl + 1 < 8) {
          copy_slab(sl + 1);   // in flight while this slab's gates are evaluated
          tmem_ld32(taddr + (sl + 1) * 32, acc);
        }
        float hv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float& cc = c[sl * 8 + u];
          if (SAVE) {
            const float ig = rec_sigmoid(pre[4 * u + 0]);
            const float fg = rec_sigmoid(pre[4 * u + 1]);
            const float gg = rec_tanh(pre[4 * u + 2]);
            const float og = rec_sigmoid(pre[4 * u + 3]);
            cc = fmaf(fg, cc, ig * gg);
            hv[u] = og * rec_tanh(cc);
            pre[4 * u + 0] = ig; pre[4 * u + 1] = fg; pre[4 * u + 2] = gg; pre[4 * u + 3] = og;
          } else {
            // the epilogue is MUFU-bound (ncu: XU 26 % of a step that is half stalls): sigma(i) tanh(g) and sigma(o) tanh(c) share
            // one reciprocal each -- (b - 1) / ((1 + a)(b + 1)), a = e^-i, b = e^2g -- 8 MUFU operations per unit instead of 10.
            // Arguments are clamped where the functions are saturated to fp32 precision (|tanh| = 1 beyond 15, sigma(-30) =
            // 9e-14), so no product of exponentials overflows.
            const float a_i = fast_expf(-fmaxf(pre[4 * u + 0], -30.f));
            const float b_g = fast_expf(2.0f * fminf(fmaxf(pre[4 * u + 2], -15.f), 15.f));
            const float ig_gg = fast_divf(b_g - 1.0f, (1.0f + a_i) * (b_g + 1.0f));
            const float fg = rec_sigmoid(pre[4 * u + 1]);
            cc = fmaf(fg, cc, ig_gg);
            const float a_o = fast_expf(-fmaxf(pre[4 * u + 3], -30.f));
            const float b_c = fast_expf(2.0f * fminf(fmaxf(cc, -15.f), 15.f));
            hv[u] = fast_divf(b_c - 1.0f, (1.0f + a_o) * (b_c + 1.0f));
          }
        }
        if (live) {
          if (out) stg256(orow + sl * 8, hv);
          if (SAVE) {
            float* gs = gates + row * ldg + dir * 512 + hf * 256 + sl * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(gs + 4 * i) = make_float4(pre[4 * i], pre[4 * i + 1], pre[4 * i + 2], pre[4 * i + 3]);
            float* cs = csave + row * D + dir * 128 + hf * 64 + sl * 8;
            *reinterpret_cast<float4*>(cs) = make_float4(c[sl * 8], c[sl * 8 + 1], c[sl * 8 + 2], c[sl * 8 + 3]);
            *reinterpret_cast<float4*>(cs + 4) = make_float4(c[sl * 8 + 4], c[sl * 8 + 5], c[sl * 8 + 6], c[sl * 8 + 7]);
          }
        }
        // h_t -> fp16 hi / lo: elements [8 sl, 8 sl + 8) of K atom `hf` of this row (the MMAs of this step have all retired:
        // acc_full is committed after the last of them, so nobody reads h_{t-1} any more)
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int u2 = 0; u2 < 4; ++u2) {
          const __half2 h2 = __floats2half2_rn(hv[2 * u2], hv[2 * u2 + 1]);
          const float2 back = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(hv[2 * u2] - back.x, hv[2 * u2 + 1] - back.y);
          hi[u2] = *reinterpret_cast<const uint32_t*>(&h2);
          lo[u2] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        if (out_hi16 && live) {
          const long long eo = row * D + dir * 128 + hf * 64 + sl * 8;
          *reinterpret_cast<uint4*>(out_hi16 + eo) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(out_lo16 + eo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        const uint32_t off = sw128_chunk_off((uint32_t)r, (uint32_t)sl);
        *reinterpret_cast<uint4*>(genH + (0 * 2 + hf) * TC_ATOM + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(genH + (1 * 2 + hf) * TC_ATOM + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();  // generic-proxy stores of h -> visible to the next step's tcgen05.mma
      tc_fence_before();         // order this thread's TMEM reads before the arrive
      __syncwarp();
      if (lane == 0) { jit(); mbar_arrive(h_local); }
      if (tid == 0) {
        jit();
        // issuer / relay duty: both CTAs hold h_t and have drained their accumulators -> next step's MMAs.  The peer's arrive is
        // RELAXED: its data sits in ITS shared memory, written by its epilogue warps who fenced generic -> async before arriving
        // on its h_local (same protocol as the relay of lstm_bf16_fused.cu, where a release arrive measured 640 ns per step)
        mbar_wait(h_local, (uint32_t)(g & 1));
        if (leader) {
          mbar_wait_cluster(peer_local, (uint32_t)(g & 1));
          tc_fence_after();
          if (st + 1 < T) issue_step();
        } else {
          mbar_arrive_cluster_relaxed(peer_local_on_leader);
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while the pair's MMAs may still read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Pipelined form (inference): the step above is  MMA (48 instructions, ~3 us) -> epilogue (MUFU-bound, ~4 us) -> MMA ...  with the
// tensor pipe idle during the epilogue and the epilogue warps idle during the product (ncu: tensor 22 %, XU 29 %).  Here the two
// column blocks of the accumulator (units 0-63 / 64-127) are STAGGERED: all eight epilogue warps work on block 0, then on block 1
// (thread = window x 32 units of the current block), and a ninth warp issues the MMAs in four groups (column block, K atom):
//
//     b0a0(t+1)  as soon as phase 0 of step t is done: it needs K atom 0 of h_t (units 0-63: written by phase 0) and accumulator
//                block 0 drained (read by phase 0) -- it runs on the tensor pipe while the warps evaluate block 1 of step t
//     after phase 1 of step t:  b0a1(t+1), commit acc_full[0];  b1a0(t+1), commit a0_free;  b1a1(t+1), commit acc_full[1]
//                the warps start on block 0 of step t+1 after ONE quarter of the product, and the rest runs under that phase
//
// Who may overwrite h: phase nb of step t+1 replaces K atom nb of h_t by h_{t+1}.  Atom 1 is read by b0a1 and b1a1, both in front
// of the commit phase 1 waits for.  Atom 0 is read by b0a0 and by b1a0 -- the latter is issued AFTER acc_full[0], so phase 0 waits
// for a third commit (a0_free) right before its first store into the atom (one slab of gate math later: b1a0 is done by then).
// Every group runs [lo.hi, hi.lo, hi.hi] over its own K atom (small terms first within the group).  Barriers per CTA: acc_full[2],
// a0_free (multicast commits), h_local[2] (one arrive per epilogue warp and phase), and on the leader peer_local[2] (the peer's
// ninth warp relays its CTA's h_local with a relaxed remote arrive, as above).
constexpr int TCP_THREADS = 384;   // 8 epilogue warps + a control warpgroup: warp 8 = MMA issuer (leader) / relay (peer), warps 9-11 idle

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCP_THREADS, 1)
lstm_rec_f16x3_pipe(const float* __restrict__ G, int ldg, const __half* __restrict__ whh16, float* __restrict__ out,
                    __half* __restrict__ out_hi16, __half* __restrict__ out_lo16, int D, int Bc, int T, int n_pairs, int ND, int jitter) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint32_t jit_state = jitter ? (uint32_t)(blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) : 0u;
  auto jit = [&]() {
    if (jitter) {
      jit_state = jit_state * 1664525u + 1013904223u;
      __nanosleep((jit_state >> 20) & (uint32_t)(jitter - 1));
    }
  };
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = tc_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + TC_OFF_H;
  uint8_t* genH = gen + TC_OFF_H;
  uint8_t* ctl = gen + TC_OFF_CTL;
  const uint32_t bar0 = smem_u32(ctl);
  auto acc_full = [&](int nb) { return bar0 + 8u * nb; };          // every CTA: multicast commit -- column block nb is complete
  auto h_local = [&](int nb) { return bar0 + 16u + 8u * nb; };     // every CTA: its 8 warps wrote K atom nb of h_t and drained block nb
  auto peer_local = [&](int nb) { return bar0 + 32u + 8u * nb; };  // leader: relay of the peer's h_local[nb]
  const uint32_t a0_free = bar0 + 48u;                             // every CTA: multicast commit -- no MMA reads K atom 0 of h_{t-1} any more
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 56);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (tid == 0) {
    for (int nb = 0; nb < 2; ++nb) { mbar_init(acc_full(nb), 1); mbar_init(h_local(nb), 8); mbar_init(peer_local(nb), 1); }
    mbar_init(a0_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();

  constexpr uint32_t idesc = umma_idesc_f16(256, 256);
  // one (column block, K atom) group in the order lo.hi, hi.lo, hi.hi; `first` clears the accumulator
  auto issue_group = [&](int nb, int atom, bool first) {
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t ap = term == 0 ? 1u : 0u, bp = term == 1 ? 1u : 0u;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = umma_desc_sw128(sH + (ap * 2 + atom) * TC_ATOM + kk * 32);
        const uint64_t db = umma_desc_sw128(sW + ((nb * 2 + bp) * 2 + atom) * TC_ATOM + kk * 32);
        umma_bf16_2sm(tmem_base + nb * 256, da, db, idesc, (first && term == 0 && kk == 0) ? 0u : 1u);
      }
    }
  };
  // everything of a step that follows its first group
  auto issue_rest = [&]() {
    issue_group(0, 1, false);
    umma_commit_2sm_mc(acc_full(0), (uint16_t)3);
    issue_group(1, 0, true);
    umma_commit_2sm_mc(a0_free, (uint16_t)3);
    issue_group(1, 1, false);
    umma_commit_2sm_mc(acc_full(1), (uint16_t)3);
  };

  const int n_work = ND * n_pairs;
  const int n_clusters = (int)cluster_nclusters_x();
  const int quarter = warp & 3, wq = (warp >> 2) & 1;
  const int r = quarter * 32 + lane;   // window row of the tile == TMEM lane
  const uint32_t peer0_on_leader = mapa_u32(peer_local(0), 0);

  // Register budget: a CTA's registers are a pool fixed at launch (threads x the compiled count, in whole groups of four warps:
  // 384 x 168 = 64 512), and `setmaxnreg` only moves registers between warpgroups through that pool -- a ninth warp alone cannot
  // give the eight epilogue warps the ~210 registers they need (a first version asked for them without a matching release and
  // spun in USETMAXREG.TRY_ALLOC forever).  So the CTA has a whole control warpgroup (warp 8 issues / relays, warps 9-11 only keep
  // the barriers company) that shrinks to 64 registers per thread, and the two epilogue warpgroups grow to 216:
  // 128 x 64 + 256 x 216 = 63 488.  The roles split ONCE, in front of the work loop, so each side is compiled against its budget.
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
    int g0 = 0;
    for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
      const int dir = w / n_pairs, tp = w - dir * n_pairs;
      const int b0 = (2 * tp + (int)rank) * TC_M;
      if (g0 > 0) cluster_sync_all();
      {
        const uint4* src = reinterpret_cast<const uint4*>(whh16 + (size_t)dir * 2 * 512 * 128);
        for (int i = tid; i < 4 * 128 * 16; i += TCP_THREADS) {
          const int blk = i >> 11, rem = i & 2047, row = rem >> 4, cc = rem & 15;   // blk = nb*2 + part
          const int nb = blk >> 1, part = blk & 1;
          const uint4 v = __ldg(src + ((size_t)part * 512 + nb * 256 + rank * 128 + row) * 16 + cc);
          *reinterpret_cast<uint4*>(gen + (blk * 2 + (cc >> 3)) * TC_ATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) = v;
        }
        for (int i = tid; i < (int)(4 * TC_ATOM / 16); i += TCP_THREADS) reinterpret_cast<uint4*>(genH)[i] = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_all();
      __syncthreads();
      cluster_sync_all();

      // ---------------- MMA issuer (leader) / relay (peer) ----------------
      if (warp != 8) {
        // idle members of the control warpgroup
      } else if (leader) {
        if (tp_elect_one()) {   // step 0: h_{-1} = 0
          issue_group(0, 0, true);
          issue_rest();
        }
        __syncwarp();
        for (int st = 0; st < T; ++st) {
          const uint32_t par = (uint32_t)((g0 + st) & 1);
          const bool more = st + 1 < T;
          mbar_wait(h_local(0), par);
          mbar_wait_cluster(peer_local(0), par);
          tc_fence_after();
          if (more && tp_elect_one()) issue_group(0, 0, true);
          __syncwarp();
          mbar_wait(h_local(1), par);
          mbar_wait_cluster(peer_local(1), par);
          tc_fence_after();
          if (more && tp_elect_one()) issue_rest();
          __syncwarp();
        }
      } else {
        for (int st = 0; st < T; ++st) {
          const uint32_t par = (uint32_t)((g0 + st) & 1);
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            mbar_wait(h_local(nb), par);
            if (tp_elect_one()) mbar_arrive_cluster_relaxed(peer0_on_leader + 8u * (uint32_t)nb);
            __syncwarp();
          }
        }
      }

      __syncthreads();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
    int g0 = 0;
    for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
      const int dir = w / n_pairs, tp = w - dir * n_pairs;
      const int b0 = (2 * tp + (int)rank) * TC_M;
      if (g0 > 0) cluster_sync_all();
      {
        const uint4* src = reinterpret_cast<const uint4*>(whh16 + (size_t)dir * 2 * 512 * 128);
        for (int i = tid; i < 4 * 128 * 16; i += TCP_THREADS) {
          const int blk = i >> 11, rem = i & 2047, row = rem >> 4, cc = rem & 15;   // blk = nb*2 + part
          const int nb = blk >> 1, part = blk & 1;
          const uint4 v = __ldg(src + ((size_t)part * 512 + nb * 256 + rank * 128 + row) * 16 + cc);
          *reinterpret_cast<uint4*>(gen + (blk * 2 + (cc >> 3)) * TC_ATOM + sw128_chunk_off((uint32_t)row, (uint32_t)(cc & 7))) = v;
        }
        for (int i = tid; i < (int)(4 * TC_ATOM / 16); i += TCP_THREADS) reinterpret_cast<uint4*>(genH)[i] = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_all();
      __syncthreads();
      cluster_sync_all();

      // ---------------- epilogue: thread = (window r, 32 units of the current column block) ----------------
      const bool live = b0 + r < Bc;
      const int brow = live ? b0 + r : Bc - 1;   // dead rows read a valid row and store nothing
      uint32_t goff[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int rr = b0 + quarter * 32 + 4 * j + (lane >> 3);
        goff[j] = (uint32_t)(rr < Bc ? rr : Bc - 1) * (uint32_t)ldg + (uint32_t)(lane & 7) * 4u;
      }
      uint8_t* stg = gen + TC_OFF_STAGE + warp * 4096;
      const uint32_t stg_s = base + TC_OFF_STAGE + warp * 4096;
      float c[64];   // [block][32 units]
#pragma unroll
      for (int i = 0; i < 64; ++i) c[i] = 0.f;

      for (int st = 0; st < T; ++st) {
        const int g = g0 + st;
        const int t = dir ? (T - 1 - st) : st;
        const long long row = (long long)t * Bc + brow;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const int col0 = nb * 256 + wq * 128;                     // this thread's 128 accumulator columns = 32 units x 4 gates
          const float* gstep = G + (long long)t * Bc * ldg + dir * 512 + col0;
          auto copy_slab = [&](int sl) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t rr = 4 * j + (lane >> 3);
              cp_async16(stg_s + rr * 128 + ((((uint32_t)lane & 7u) ^ (rr & 7u)) << 4), gstep + goff[j] + sl * 32);
            }
            cp_async_commit();
          };
          copy_slab(0);
          if (lane == 0) jit();
          __syncwarp();
          mbar_wait(acc_full(nb), (uint32_t)(g & 1));
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
          uint32_t acc[32];
          tmem_ld32(taddr, acc);
          const int unit0 = nb * 64 + wq * 32;                       // first hidden unit of this thread in this phase
          float* orow = out + row * D + dir * 128 + unit0;
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            cp_async_wait_all();
            __syncwarp();
            tmem_ld_wait();
            float pre[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 gv = *reinterpret_cast<const uint4*>(stg + lane * 128 + (((uint32_t)i ^ ((uint32_t)lane & 7u)) << 4));
              pre[4 * i + 0] = fmaf(__uint_as_float(acc[4 * i + 0]), 1.0f / TC_WSCALE, __uint_as_float(gv.x));
              pre[4 * i + 1] = fmaf(__uint_as_float(acc[4 * i + 1]), 1.0f / TC_WSCALE, __uint_as_float(gv.y));
              pre[4 * i + 2] = fmaf(__uint_as_float(acc[4 * i + 2]), 1.0f / TC_WSCALE, __uint_as_float(gv.z));
              pre[4 * i + 3] = fmaf(__uint_as_float(acc[4 * i + 3]), 1.0f / TC_WSCALE, __uint_as_float(gv.w));
            }
            __syncwarp();
            if (sl + 1 < 4) {
              copy_slab(sl + 1);
              tmem_ld32(taddr + (sl + 1) * 32, acc);
            }
            float hv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float& cc = c[nb * 32 + sl * 8 + u];
              const float a_i = fast_expf(-fmaxf(pre[4 * u + 0], -30.f));
              const float b_g = fast_expf(2.0f * fminf(fmaxf(pre[4 * u + 2], -15.f), 15.f));
              const float ig_gg = fast_divf(b_g - 1.0f, (1.0f + a_i) * (b_g + 1.0f));
              const float fg = rec_sigmoid(pre[4 * u + 1]);
              cc = fmaf(fg, cc, ig_gg);
              const float a_o = fast_expf(-fmaxf(pre[4 * u + 3], -30.f));
              const float b_c = fast_expf(2.0f * fminf(fmaxf(cc, -15.f), 15.f));
              hv[u] = fast_divf(b_c - 1.0f, (1.0f + a_o) * (b_c + 1.0f));
            }
            if (live && out) stg256(orow + sl * 8, hv);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; ++u2) {
              const __half2 h2 = __floats2half2_rn(hv[2 * u2], hv[2 * u2 + 1]);
              const float2 back = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(hv[2 * u2] - back.x, hv[2 * u2 + 1] - back.y);
              hi[u2] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[u2] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            if (out_hi16 && live) {
              const long long eo = row * D + dir * 128 + unit0 + sl * 8;
              *reinterpret_cast<uint4*>(out_hi16 + eo) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(out_lo16 + eo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            // K atom nb of h_t, 16-byte chunk wq*4 + sl of this row (see "who may overwrite h" above)
            if (nb == 0 && sl == 0) mbar_wait(a0_free, (uint32_t)(g & 1));
            const uint32_t off = sw128_chunk_off((uint32_t)r, (uint32_t)(wq * 4 + sl));
            *reinterpret_cast<uint4*>(genH + (0 * 2 + nb) * TC_ATOM + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(genH + (1 * 2 + nb) * TC_ATOM + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { jit(); mbar_arrive(h_local(nb)); }
        }
      }

      __syncthreads();
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int tc_setup(int* max_clusters_out) {
  static PerDeviceInt state_pd, max_pd;  // state: 0 = not tried, 1 = ok, -1 = unavailable
  int& state = state_pd.cur();
  int& max_clusters = max_pd.cur();
  if (state == 0) {
    state = -1;
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_f16x3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_f16x3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count(), 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TC_SMEM;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    cfg.attrs = la; cfg.numAttrs = 1;
    BCI_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, lstm_rec_f16x3<false>, &cfg));
    BCI_REQUIRE(max_clusters > 0, BCI_ECUDA, "fp32 tensor-core recurrence: no CTA pair fits on this device");
    state = 1;
  }
  if (max_clusters_out) *max_clusters_out = max_clusters;
  return state == 1 ? BCI_OK : BCI_ECUDA;
}

// BCI_FP32_REC=simt keeps the CUDA-core recurrence (the first version of the fp32 path) for comparison
bool tc_rec_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_FP32_REC");
    v = (e && e[0] == 's') ? 0 : 1;
  }
  return v != 0;
}

int tc_max_clusters() {
  int n = 0;
  return tc_setup(&n) == BCI_OK ? n : 0;
}

// the pair kernel pays off once its work items (256 windows x direction) cover a good part of the machine; below that the
// latency-oriented CUDA-core variants (small tiles on many SMs) are faster
bool tc_rec_ok(int H, int ND, int Bc, const void* G, int ldg, const void* out, int D) {
  if (!tc_rec_enabled() || H != 128) return false;
  if (((uintptr_t)G & 15) || ((uintptr_t)out & 31) || (ldg & 3) || (D & 7)) return false;   // 16-byte cp.async, 32-byte stores
  return ND * ceil_div(Bc, 2 * TC_M) >= 16 && tc_max_clusters() > 0;
}

int launch_rec_f16x3(int ND, const float* G, int ldg, const __half* whh16, float* out, __half* out_hi16, __half* out_lo16, float* gates,
                     float* csave, int D, int Bc, int T, cudaStream_t st) {
  int max_clusters = 0;
  int rc = tc_setup(&max_clusters);
  if (rc) return rc;
  const int n_pairs = ceil_div(Bc, 2 * TC_M);
  const int items = ND * n_pairs;
  const int clusters = items < max_clusters ? items : max_clusters;
  // L2 prefetch of the next step's rows (BCI_TC_PF = 1 bulk per thread, 2 rolling per slab) measured no gain and doubled the DRAM
  // reads (ncu: 17.8 GB against 9.9 GB of G, L2 hit rate 14 %): off by default
  static const int pf_mode = [] { const char* e = getenv("BCI_TC_PF"); return e ? atoi(e) : 0; }();
  static const int jitter = [] { const char* e = getenv("BCI_FUSED_JITTER"); int v = e ? atoi(e) : 0; return (v > 0 && (v & (v - 1)) == 0) ? v : 0; }();
  // BCI_TC_PIPE=0 keeps the plain (unstaggered) kernel for inference as well
  static const bool pipe = [] { const char* e = getenv("BCI_TC_PIPE"); return !(e && e[0] == '0'); }();
  if (!gates && pipe) {
    static PerDeviceFlag attr_pd;
    bool& attr = attr_pd.cur();
    if (!attr) {
      BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec_f16x3_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
      attr = true;
    }
    lstm_rec_f16x3_pipe<<<2 * clusters, TCP_THREADS, TC_SMEM, st>>>(G, ldg, whh16, out, out_hi16, out_lo16, D, Bc, T, n_pairs, ND, jitter);
    BCI_LAUNCH_OK();
    return BCI_OK;
  }
  if (gates) lstm_rec_f16x3<true><<<2 * clusters, TC_THREADS, TC_SMEM, st>>>(G, ldg, whh16, out, out_hi16, out_lo16, gates, csave, D, Bc, T, n_pairs, ND, pf_mode, jitter);
  else lstm_rec_f16x3<false><<<2 * clusters, TC_THREADS, TC_SMEM, st>>>(G, ldg, whh16, out, out_hi16, out_lo16, nullptr, nullptr, D, Bc, T, n_pairs, ND, pf_mode, jitter);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci

// diagnostics (tests/test_gpu_tensorcore.py): the pair recurrence in isolation.  G [T*Bc][ND*512] fp32 (column dir*512 + unit*4 + gate),
// w_hh [ND][512][128] fp32 in the PyTorch layout (gate-major rows), packed [ND][2][512][128] fp16 scratch, out [T][Bc][ND*128] fp32
extern "C" int bci_selftest_rec_f16x3(const float* G, const float* w_hh, void* packed, float* out, int32_t Bc, int32_t T, int32_t ND,
                                      void* stream) {
  using namespace bci;
  BCI_REQUIRE(G && w_hh && packed && out && Bc >= 1 && T >= 1 && (ND == 1 || ND == 2), BCI_EINVAL, "bci_selftest_rec_f16x3: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int d = 0; d < ND; ++d) {
    int rc = pack_whh_f16x3(w_hh + (size_t)d * 512 * 128, reinterpret_cast<__half*>(packed) + (size_t)d * 2 * 512 * 128, 128, st);
    if (rc) return rc;
  }
  return launch_rec_f16x3(ND, G, ND * 512, reinterpret_cast<const __half*>(packed), out, nullptr, nullptr, nullptr, nullptr, ND * 128, Bc, T, st);
}
