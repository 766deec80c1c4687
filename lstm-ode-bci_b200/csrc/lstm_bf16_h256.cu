// bf16 tensor-core mode for hidden_size = 256 -- the size of the reference's trained checkpoint on 61 channels
// (04_lstm_model.py:876-877: hidden_size = 256 if n_channels > 30).
//
// W_hh of one direction is 1024 x 256 bf16 = 512 KB and W_ih of layers 1-2 another 1 MB: the fully fused kernel of
// lstm_bf16_fused.cu (both matrices resident in a 4-CTA cluster) would need 16 SMs per cluster.  This first H = 256 version is
// the hybrid: the time-parallel projection G = in . W_ih^T + b is the tcgen05 GEMM of lstm_bf16.cu (128-column W blocks, K up
// to 512, blocked streaming layout), and the recurrence runs on a 4-CTA cluster that keeps W_hh resident:
//
//   cluster rank r = 2 p + s    p = half of the hidden units ([128 p, 128 p + 128): 512 gate columns), s = window tile
//   pair p = CTAs (p,0),(p,1)   tcgen05.mma.cta_group::2, M = 256 (both tiles) x N = 256 x K = 16, two N blocks: each CTA keeps
//                               2 x 128 of the pair's 512 W_hh rows (128 KB) and its tile's h (4 K-atoms, 64 KB, single-buffered);
//                               the 128 x 512 fp32 accumulator fills the CTA's TMEM
//   per step                    MMA_hh -> commit (multicast to the pair) -> 8 epilogue warps (thread = window row x 64 units):
//                               tcgen05.ld + G_t (coalesced 16-byte streaming loads, two slabs ahead) -> sigma/tanh -> fp32 cell
//                               state in registers -> h_t (bf16) into the operand buffer; the CTA's two atoms (32 KB) go to the
//                               CTA (1-p, s) with one DSMEM bulk copy; handshakes as in lstm_bf16_fused.cu (recv_ready, h_in
//                               re-armed by its waiter, relaxed relays).  One tile per CTA: TMEM is full, so MMA latency and
//                               the exchange are exposed here -- the price of H = 256 until the 16-SM design exists.
//
// Gate column order ("perm_256"): R = p*512 + nb*256 + slab*32 + gate*8 + u  for unit = 128 p + 64 nb + 8 slab + u, used for the
// rows of W_hh and W_ih, the bias and G's columns (a 16-byte chunk of G = 8 units of one gate; the four gates of a slab are
// four consecutive chunks).
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <cstdlib>

namespace bci {
using namespace sm100;

__host__ __device__ constexpr int perm_256(int unit, int gate) {
  return (unit >> 7) * 512 + ((unit >> 6) & 1) * 256 + ((unit >> 3) & 7) * 32 + gate * 8 + (unit & 7);
}

// src (4H, K) gate-major rows -> dst bf16 [row0 + perm_256(unit, gate)][K], i/f/o rows pre-scaled by 1/2 (sigmoid via tanh)
__global__ void pack_rows_perm256_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int K, int row0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)1024 * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int gate = row >> 8, unit = row & 255;
  const float sc = (gate == 2) ? 1.0f : 0.5f;
  dst[(long long)(row0 + perm_256(unit, gate)) * K + k] = __float2bfloat16_rn(sc * src[i]);
}
__global__ void pack_bias_perm256(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ dst, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1024) return;
  const int gate = i >> 8, unit = i & 255;
  dst[col0 + perm_256(unit, gate)] = ((gate == 2) ? 1.0f : 0.5f) * (bih[i] + bhh[i]);
}

constexpr int HR_M = 128, HR_EPI_WARPS = 8;
constexpr int HR_THREADS = (HR_EPI_WARPS + 3) * 32;  // + MMA issuer / relay, h store warp, G prefetcher
constexpr uint32_t HR_ATOM = 128 * 128;              // [128 rows][64 bf16] SW128 atom
constexpr uint32_t HR_OFF_H = 8 * HR_ATOM;           // W: 2 N blocks x 4 K-atoms
constexpr uint32_t HR_OFF_CTL = HR_OFF_H + 4 * HR_ATOM;
constexpr size_t HR_SMEM = 1024 + HR_OFF_CTL + 256;

__device__ __forceinline__ float hr_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// work item w (0 .. 2*tile_pairs): direction = w / tile_pairs, tile pair = w % tile_pairs; CTA (p, s) owns tile 2 pair + s
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(HR_THREADS, 1)
lstm_rec256_bf16(const __nv_bfloat16* __restrict__ G,        // blocked: [row/128][256 chunks][row%128][8], columns dir*1024 + perm_256
                 const __grid_constant__ CUtensorMap tmOut,  // out [T][Bc][512] bf16, box 64 x 128 x 1
                 const __nv_bfloat16* __restrict__ whh,      // [2][1024][256] rows in perm_256 order (i,f,o pre-scaled by 1/2)
                 int Bc, int T, int tile_pairs, int jitter) {
  extern __shared__ uint8_t hr_smem_raw[];
  uint32_t jit_state = jitter ? (uint32_t)(blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) : 0u;
  auto jit = [&]() {
    if (jitter) {
      jit_state = jit_state * 1664525u + 1013904223u;
      __nanosleep((jit_state >> 20) & (uint32_t)(jitter - 1));
    }
  };
  const uint32_t raw = smem_u32(hr_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = hr_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + HR_OFF_H;
  uint8_t* genH = gen + HR_OFF_H;
  uint8_t* ctl = gen + HR_OFF_CTL;
  const uint32_t bar0 = smem_u32(ctl);
  // every barrier completes once per step g: parity g & 1
  // acc_full[nb] / h_local[ch]: the two N blocks (= the two 64-unit column halves) are committed and consumed separately, so the
  // epilogue warps of block 0 start while the MMAs of block 1 run, and the first K-atom of h_t is on its way to the partner
  // while the second is still being computed
  const uint32_t acc_full0 = bar0, h_local0 = bar0 + 8, h_in = bar0 + 16, st_free = bar0 + 24, peer_local = bar0 + 32,
                 peer_in = bar0 + 40, copy_done = bar0 + 48, recv_ready = bar0 + 56, acc_full1 = bar0 + 64, h_local1 = bar0 + 72,
                 loc_free = bar0 + 80;  // the MMAs that read this CTA's OWN two atoms of h_{t-1} have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 88);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int p = (int)(rank >> 1), s = (int)(rank & 1);
  const bool leader = (s == 0);
  const uint32_t partner = (uint32_t)(2 * (1 - p) + s);

  if (tid == 0) {
    mbar_init(acc_full0, 1);
    mbar_init(acc_full1, 1);
    mbar_init(loc_free, 1);
    mbar_init(h_local0, HR_EPI_WARPS / 2);
    mbar_init(h_local1, HR_EPI_WARPS / 2);
    mbar_init(h_in, 1);
    mbar_init(st_free, 1);
    mbar_init(peer_local, 1);
    mbar_init(peer_in, 1);
    mbar_init(copy_done, 1);
    mbar_init(recv_ready, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(h_in, 2 * HR_ATOM);  // armed for the first step; re-armed by its single waiter afterwards
    tma_prefetch_desc(&tmOut);
  }
  if (warp == HR_EPI_WARPS) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();

  const int n_work = 2 * tile_pairs;
  const int n_clusters = (int)cluster_nclusters_x();
  int g0 = 0;
  for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
    const int dir = w / tile_pairs, tp = w - dir * tile_pairs;
    const int b0 = (2 * tp + s) * HR_M;

    if (g0 > 0) cluster_sync_all();
    {
      // this CTA's W_hh rows: for N block nb, rows [p*512 + nb*256 + 128 s, +128); h_{-1} = 0
      for (int i = tid; i < 2 * 128 * 32; i += HR_THREADS) {
        const uint32_t nb = i >> 12, rem = i & 4095, row = rem >> 5, cc = rem & 31, atom = cc >> 3, c = cc & 7;
        const uint4* src = reinterpret_cast<const uint4*>(whh + ((size_t)dir * 1024 + p * 512 + nb * 256 + 128 * s + row) * 256);
        *reinterpret_cast<uint4*>(gen + (nb * 4 + atom) * HR_ATOM + sw128_chunk_off(row, c)) = __ldg(src + cc);
      }
      for (int i = tid; i < (int)(4 * HR_ATOM / 16); i += HR_THREADS) reinterpret_cast<uint4*>(genH)[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_all();
    __syncthreads();
    cluster_sync_all();

    // chunk 0 of (row, dir) in the blocked G layout: 256 chunks of 2 KB per 128-row block
    auto g_block = [&](long long row) { return reinterpret_cast<const uint8_t*>(G) + (row >> 7) * (256ll * 2048); };

    if (warp == HR_EPI_WARPS + 2) {
      // ---------------- L2 prefetcher: this CTA's part of the NEXT step's G block (64 chunks of 2 KB = 128 KB) ----------------
      const long long n_blocks = ((long long)T * Bc + 127) >> 7;  // G is allocated in whole 128-row blocks
      for (int st = 0; st + 1 < T && b0 < Bc; ++st) {
        const int sn = st + 1;
        const long long row0 = (long long)(dir ? (T - 1 - sn) : sn) * Bc + b0;
        const uint8_t* blk = g_block(row0) + (long long)(dir * 128 + p * 64) * 2048;
        if (lane < 8) bulk_prefetch_l2(blk + lane * 16384, 16384u);
        if ((row0 & 127) != 0 && (row0 >> 7) + 1 < n_blocks && lane >= 8 && lane < 16)
          bulk_prefetch_l2(blk + 256ll * 2048 + (lane - 8) * 16384, 16384u);
        // pace: one step of prefetch per step of compute.  Nothing depends on this warp, so it may fall behind and see the
        // barrier two phases later (same parity): bounded polling instead of a wait that could then never return
        for (int polls = 0; polls < 50000 && !mbar_try_wait(h_local1, (uint32_t)((g0 + st) & 1)); ++polls) { }
      }
    } else if (warp == HR_EPI_WARPS) {
      if (leader && lane == 0) {
        // ---------------- MMA issuer (pair leader) ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
        const uint16_t pair_mask = (uint16_t)(3u << (2 * p));
        const uint32_t ack = mapa_u32(copy_done, partner);
        auto wait_h = [&](int gp) {  // h of step gp complete in both CTAs of the pair, accumulator drained
          jit();
          mbar_wait(h_local0, (uint32_t)(gp & 1));
          mbar_wait(h_local1, (uint32_t)(gp & 1));
          mbar_wait_cluster(peer_local, (uint32_t)(gp & 1));
          mbar_wait(h_in, (uint32_t)(gp & 1));
          mbar_arrive_expect_tx(h_in, 2 * HR_ATOM);
          mbar_arrive_cluster_relaxed(ack);
          mbar_wait_cluster(peer_in, (uint32_t)(gp & 1));
          tc_fence_after();
        };
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
          if (st > 0) wait_h(g - 1);
          // h is single-buffered and the epilogue warps of N block 0 start (and overwrite this pair's own atoms 2p, 2p+1 with
          // h_t) while N block 1 is still being multiplied: block 1 therefore consumes the pair's own atoms FIRST and commits
          // `loc_free` once those K slices have retired; the other pair's atoms are only replaced after recv_ready (all MMAs done)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const uint32_t atom = (uint32_t)(((k >> 2) + 2 * p) & 3), kk = k & 3;
              const uint64_t da = umma_desc_sw128(sH + atom * HR_ATOM + kk * 32);
              const uint64_t db = umma_desc_sw128(sW + (nb * 4 + atom) * HR_ATOM + kk * 32);
              umma_bf16_2sm(tmem_base + nb * 256, da, db, idesc, k != 0 ? 1u : 0u);
              if (nb == 1 && k == 7) umma_commit_2sm_mc(loc_free, pair_mask);
            }
            umma_commit_2sm_mc(nb == 0 ? acc_full0 : acc_full1, pair_mask);
          }
        }
        wait_h(g0 + T - 1);
      } else if (!leader && lane == 0) {
        // ---------------- relay (peer CTA of the pair): relaxed arrives, see lstm_bf16_fused.cu ----------------
        const uint32_t pl = mapa_u32(peer_local, rank & ~1u), pi = mapa_u32(peer_in, rank & ~1u);
        const uint32_t ack = mapa_u32(copy_done, partner);
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
          jit();
          mbar_wait(h_local0, (uint32_t)(g & 1));
          mbar_wait(h_local1, (uint32_t)(g & 1));
          mbar_arrive_cluster_relaxed(pl);
          mbar_wait(h_in, (uint32_t)(g & 1));
          mbar_arrive_expect_tx(h_in, 2 * HR_ATOM);
          mbar_arrive_cluster_relaxed(pi);
          mbar_arrive_cluster_relaxed(ack);
        }
      }
    } else if (warp == HR_EPI_WARPS + 1) {
      // ---------------- h store warp: this CTA's two K-atoms of h_t -> partner CTA (DSMEM) and -> out[t] ----------------
      if (lane == 0) {
        const uint32_t atoms = sH + 2 * p * HR_ATOM;
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
          const int t = dir ? (T - 1 - st) : st;
          jit();
          mbar_wait(h_local0, (uint32_t)(g & 1));
          mbar_wait_cluster(recv_ready, (uint32_t)(g & 1));  // the partner pair's MMAs of this step (both N blocks) have retired
          bulk_copy_s2s_cluster(mapa_u32(atoms, partner), atoms, HR_ATOM, mapa_u32(h_in, partner));
          tma_store_3d(&tmOut, atoms, dir * 256 + 128 * p, b0, t);
          mbar_wait(h_local1, (uint32_t)(g & 1));
          bulk_copy_s2s_cluster(mapa_u32(atoms + HR_ATOM, partner), atoms + HR_ATOM, HR_ATOM, mapa_u32(h_in, partner));
          tma_store_3d(&tmOut, atoms + HR_ATOM, dir * 256 + 128 * p + 64, b0, t);
          tma_store_commit();
          tma_store_wait_read();
          mbar_arrive(st_free);
        }
        tma_store_wait_all();
      }
    } else {
      // ---------------- epilogue: thread = (window row, 64 of the CTA's 128 hidden units) ----------------
      const int quarter = warp & 3, ch = warp >> 2;
      const int r = quarter * 32 + lane;
      const bool live = b0 + r < Bc;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ch * 256;
      uint8_t* hrow = genH + (2 * p + ch) * HR_ATOM;  // this thread's 64 units = K-atom 2p + ch
      const uint32_t rr = mapa_u32(recv_ready, partner);
      float c[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) c[i] = 0.f;

      for (int st = 0; st < T; ++st) {
        const int g = g0 + st;
        const int t = dir ? (T - 1 - st) : st;
        // rows of windows beyond the batch (partial or absent tiles) read a valid row instead; their results are never stored
        const long long row = (long long)t * Bc + (live ? b0 + r : (b0 < Bc ? b0 : 0));
        // chunks (sl*4 + gate) of this thread's 64 units: units 128 p + 64 ch + 8 sl .. +7
        const uint8_t* gp = g_block(row) + (long long)(dir * 128 + p * 64 + ch * 32) * 2048 + (row & 127) * 16ll;
        uint4 gbuf[3][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gbuf[0][q] = ldg_stream_v4(gp + q * 2048);
#pragma unroll
        for (int q = 0; q < 4; ++q) gbuf[1][q] = ldg_stream_v4(gp + (4 + q) * 2048);
        if (lane == 0) jit();
        __syncwarp();
        mbar_wait(ch == 0 ? acc_full0 : acc_full1, (uint32_t)(g & 1));
        tc_fence_after();
        // all MMAs of this step (the second N block is committed last) retired: the partner may send its atoms of h_g
        if (warp == 4 && lane == 0) mbar_arrive_cluster_relaxed(rr);
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
          uint32_t acc[32];
          tmem_ld32(taddr + sl * 32, acc);
          if (sl < 6) {
#pragma unroll
            for (int q = 0; q < 4; ++q) gbuf[(sl + 2) % 3][q] = ldg_stream_v4(gp + ((sl + 2) * 4 + q) * 2048);
          }
          tmem_ld_wait();
          const uint32_t* gw = reinterpret_cast<const uint32_t*>(gbuf[sl % 3]);
          auto gval = [&](int gate, int u) {
            const uint32_t wv = gw[gate * 4 + (u >> 1)];
            return __uint_as_float((u & 1) ? (wv & 0xFFFF0000u) : (wv << 16));
          };
          uint32_t hp[4];
#pragma unroll
          for (int u2 = 0; u2 < 4; ++u2) {
            float hv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int u = u2 * 2 + e;
              const float ig = fmaf(0.5f, hr_tanh(__uint_as_float(acc[0 * 8 + u]) + gval(0, u)), 0.5f);
              const float fg = fmaf(0.5f, hr_tanh(__uint_as_float(acc[1 * 8 + u]) + gval(1, u)), 0.5f);
              const float gg = hr_tanh(__uint_as_float(acc[2 * 8 + u]) + gval(2, u));
              const float og = fmaf(0.5f, hr_tanh(__uint_as_float(acc[3 * 8 + u]) + gval(3, u)), 0.5f);
              float& cc = c[sl * 8 + u];
              cc = fmaf(fg, cc, ig * gg);
              hv[e] = og * hr_tanh(cc);
            }
            __nv_bfloat162 pk = __floats2bfloat162_rn(hv[0], hv[1]);
            hp[u2] = *reinterpret_cast<uint32_t*>(&pk);
          }
          // the local atoms still hold h_{g-1}: their TMA store and their copy to the partner must have finished reading them
          if (sl == 0) {
            if (g > 0) {
              mbar_wait(st_free, (uint32_t)((g - 1) & 1));
              mbar_wait_cluster(copy_done, (uint32_t)((g - 1) & 1));
            }
            mbar_wait(loc_free, (uint32_t)(g & 1));  // (N block 0's warps run ahead of N block 1's MMAs)
          }
          *reinterpret_cast<uint4*>(hrow + sw128_chunk_off((uint32_t)r, (uint32_t)sl)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ch == 0 ? h_local0 : h_local1);
      }
    }
    __syncthreads();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == HR_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Staggered form of the recurrence above.  There a step is  MMA block 0 -> MMA block 1 -> both 4-warp epilogue halves side by side ->
// exchange -> next MMA: ~8 us, with the tensor pipe idle during the gates and one warp per scheduler in each half.  Here (as in
// lstm_rec_f16x3_pipe, lstm_fp32_tc.cu) all eight epilogue warps evaluate accumulator block 0 (units 128 p + 0..63), then block 1,
// and the next step's MMAs are issued in (block, K atom) groups as soon as their operands exist:
//
//     own atom nb  (units 128 p + 64 nb ..)  is written by phase nb of both CTAs of the pair        -> h_local[nb] (+ relay)
//     partner atom nb                         arrives by DSMEM bulk copy from CTA (1 - p, s)          -> h_in[nb]   (+ relay)
//     step t+1:  b0.own0 | b0.par0 | b0.own1 | b0.par1, commit acc_full[0] | b1.own0, commit a0_free | b1.par0, b1.own1, b1.par1,
//                commit acc_full[1]      (each group waits only for its own atom; block 1 runs under phase 0 of step t+1)
//
// Who may overwrite what: phase nb replaces own atom nb (h_t by h_{t+1}) after (i) every MMA that reads it has retired -- atom 1:
// all in front of acc_full[1]; atom 0: b1.own0 follows acc_full[0], hence the extra commit a0_free -- and (ii) its TMA store and
// its copy to the partner have finished reading it (st_free[nb], copy_done[nb] = the receiver's ack).  The partner's copy of
// step t+1 may land in this CTA's partner atoms once this pair's MMAs of step t+1 have all retired: the h-store warp forwards
// acc_full[1] to the partner as recv_ready.  h_in[nb] are transaction barriers re-armed by their single waiter.
// Register budget (see lstm_rec_f16x3_pipe): control warpgroup (warp 8 MMA / relay, 9 h store, 10 L2 prefetch, 11 idle) at 96
// registers, the two epilogue warpgroups at 200: 128 x 96 + 256 x 200 = 63 488 of the CTA's 384 x 168.
constexpr int HP_THREADS = 384;

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(HP_THREADS, 1)
lstm_rec256_bf16_pipe(const __nv_bfloat16* __restrict__ G, const __grid_constant__ CUtensorMap tmOut,
                      const __nv_bfloat16* __restrict__ whh, int Bc, int T, int tile_pairs, int jitter) {
  extern __shared__ uint8_t hr_smem_raw[];
  uint32_t jit_state = jitter ? (uint32_t)(blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) : 0u;
  auto jit = [&]() {
    if (jitter) {
      jit_state = jit_state * 1664525u + 1013904223u;
      __nanosleep((jit_state >> 20) & (uint32_t)(jitter - 1));
    }
  };
  const uint32_t raw = smem_u32(hr_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = hr_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + HR_OFF_H;
  uint8_t* genH = gen + HR_OFF_H;
  uint8_t* ctl = gen + HR_OFF_CTL;
  const uint32_t bar0 = smem_u32(ctl);
  // every barrier completes once per step g: parity g & 1
  auto acc_full = [&](int nb) { return bar0 + 8u * nb; };
  const uint32_t a0_free = bar0 + 16u;
  auto h_local = [&](int nb) { return bar0 + 24u + 8u * nb; };
  auto h_in = [&](int nb) { return bar0 + 40u + 8u * nb; };
  auto st_free = [&](int nb) { return bar0 + 56u + 8u * nb; };
  auto copy_done = [&](int nb) { return bar0 + 72u + 8u * nb; };
  const uint32_t recv_ready = bar0 + 88u;
  auto peer_local = [&](int nb) { return bar0 + 96u + 8u * nb; };
  auto peer_in = [&](int nb) { return bar0 + 112u + 8u * nb; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 128);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  const int p = (int)(rank >> 1), s = (int)(rank & 1);
  const bool leader = (s == 0);
  const uint32_t partner = (uint32_t)(2 * (1 - p) + s);

  if (tid == 0) {
    for (int nb = 0; nb < 2; ++nb) {
      mbar_init(acc_full(nb), 1);
      mbar_init(h_local(nb), HR_EPI_WARPS);
      mbar_init(h_in(nb), 1);
      mbar_init(st_free(nb), 1);
      mbar_init(copy_done(nb), 1);
      mbar_init(peer_local(nb), 1);
      mbar_init(peer_in(nb), 1);
    }
    mbar_init(a0_free, 1);
    mbar_init(recv_ready, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(h_in(0), HR_ATOM);   // armed for the first step; re-armed by their single waiter afterwards
    mbar_arrive_expect_tx(h_in(1), HR_ATOM);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == HR_EPI_WARPS) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();

  const int n_work = 2 * tile_pairs;
  const int n_clusters = (int)cluster_nclusters_x();
  // start of a work item, executed by every thread: this CTA's W_hh rows (for N block nb, rows [p*512 + nb*256 + 128 s, +128)), h_{-1} = 0
  auto item_begin = [&](int dir, bool first) {
    if (!first) cluster_sync_all();
    for (int i = tid; i < 2 * 128 * 32; i += HP_THREADS) {
      const uint32_t nb = i >> 12, rem = i & 4095, row = rem >> 5, cc = rem & 31, atom = cc >> 3, c = cc & 7;
      const uint4* src = reinterpret_cast<const uint4*>(whh + ((size_t)dir * 1024 + p * 512 + nb * 256 + 128 * s + row) * 256);
      *reinterpret_cast<uint4*>(gen + (nb * 4 + atom) * HR_ATOM + sw128_chunk_off(row, c)) = __ldg(src + cc);
    }
    for (int i = tid; i < (int)(4 * HR_ATOM / 16); i += HP_THREADS) reinterpret_cast<uint4*>(genH)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_all();
    __syncthreads();
    cluster_sync_all();
  };
  // chunk 0 of a row's 128-row block in the blocked G layout: 256 chunks of 2 KB per block
  auto g_block = [&](long long row) { return reinterpret_cast<const uint8_t*>(G) + (row >> 7) * (256ll * 2048); };

  if (warp >= HR_EPI_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;" ::: "memory");
    int g0 = 0;
    for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
      const int dir = w / tile_pairs, tp = w - dir * tile_pairs;
      const int b0 = (2 * tp + s) * HR_M;
      item_begin(dir, g0 == 0);
      if (warp == HR_EPI_WARPS + 2) {
        // ---------------- L2 prefetcher: this CTA's part of the NEXT step's G block (64 chunks of 2 KB = 128 KB) ----------------
        const long long n_blocks = ((long long)T * Bc + 127) >> 7;
        for (int st = 0; st + 1 < T && b0 < Bc; ++st) {
          const int sn = st + 1;
          const long long row0 = (long long)(dir ? (T - 1 - sn) : sn) * Bc + b0;
          const uint8_t* blk = g_block(row0) + (long long)(dir * 128 + p * 64) * 2048;
          if (lane < 8) bulk_prefetch_l2(blk + lane * 16384, 16384u);
          if ((row0 & 127) != 0 && (row0 >> 7) + 1 < n_blocks && lane >= 8 && lane < 16)
            bulk_prefetch_l2(blk + 256ll * 2048 + (lane - 8) * 16384, 16384u);
          // pace: nothing depends on this warp, so it may fall behind and see the barrier two phases later (same parity):
          // bounded polling instead of a wait that could then never return
          for (int polls = 0; polls < 50000 && !mbar_try_wait(h_local(1), (uint32_t)((g0 + st) & 1)); ++polls) { }
        }
      } else if (warp == HR_EPI_WARPS + 1) {
        // ---------------- h store warp: this CTA's two K-atoms of h_t -> partner CTA (DSMEM) and -> out[t] ----------------
        if (lane == 0) {
          const uint32_t atoms = sH + 2 * p * HR_ATOM;
          const uint32_t rr = mapa_u32(recv_ready, partner);
          for (int st = 0; st < T; ++st) {
            const uint32_t par = (uint32_t)((g0 + st) & 1);
            const int t = dir ? (T - 1 - st) : st;
            jit();
            mbar_wait(acc_full(1), par);              // this pair's MMAs of the step have all retired:
            mbar_arrive_cluster_relaxed(rr);          // the partner may replace its atoms in this CTA
            mbar_wait(h_local(0), par);
            mbar_wait_cluster(recv_ready, par);       // ... and this CTA its atoms in the partner
            bulk_copy_s2s_cluster(mapa_u32(atoms, partner), atoms, HR_ATOM, mapa_u32(h_in(0), partner));
            tma_store_3d(&tmOut, atoms, dir * 256 + 128 * p, b0, t);
            tma_store_commit();
            mbar_wait(h_local(1), par);
            bulk_copy_s2s_cluster(mapa_u32(atoms + HR_ATOM, partner), atoms + HR_ATOM, HR_ATOM, mapa_u32(h_in(1), partner));
            tma_store_3d(&tmOut, atoms + HR_ATOM, dir * 256 + 128 * p + 64, b0, t);
            tma_store_commit();
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(st_free(0));
            tma_store_wait_read();
            mbar_arrive(st_free(1));
          }
          tma_store_wait_all();
        }
      } else if (warp == HR_EPI_WARPS) {
        // ---------------- MMA issuer (pair leader) / relay (its peer) ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
        const uint16_t pair_mask = (uint16_t)(3u << (2 * p));
        const uint32_t own = (uint32_t)(2 * p), par_atom = (uint32_t)(2 * (1 - p));
        auto issue4 = [&](int nb, uint32_t atom, bool first) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc_sw128(sH + atom * HR_ATOM + kk * 32);
            const uint64_t db = umma_desc_sw128(sW + (nb * 4 + atom) * HR_ATOM + kk * 32);
            umma_bf16_2sm(tmem_base + nb * 256, da, db, idesc, (first && kk == 0) ? 0u : 1u);
          }
        };
        auto issue_tail = [&]() {   // everything of a step behind b0.own0, b0.par0, b0.own1
          issue4(0, par_atom + 1, false);
          umma_commit_2sm_mc(acc_full(0), pair_mask);
          issue4(1, own, true);
          umma_commit_2sm_mc(a0_free, pair_mask);
          issue4(1, par_atom, false);
          issue4(1, own + 1, false);
          issue4(1, par_atom + 1, false);
          umma_commit_2sm_mc(acc_full(1), pair_mask);
        };
        const uint32_t lead = rank & ~1u;
        const uint32_t ack0 = mapa_u32(copy_done(0), partner);
        const uint32_t pl0 = mapa_u32(peer_local(0), lead), pi0 = mapa_u32(peer_in(0), lead);
        if (leader) {
          if (tp_elect_one()) {   // step 0: h_{-1} = 0
            issue4(0, own, true);
            issue4(0, par_atom, false);
            issue4(0, own + 1, false);
            issue_tail();
          }
          __syncwarp();
        }
        for (int st = 0; st < T; ++st) {
          const uint32_t par = (uint32_t)((g0 + st) & 1);
          const bool more = st + 1 < T;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            if (lane == 0) jit();
            __syncwarp();
            // own atom nb of h_t (and accumulator block nb drained)
            mbar_wait(h_local(nb), par);
            if (leader) {
              mbar_wait_cluster(peer_local(nb), par);
              tc_fence_after();
              if (more && tp_elect_one()) issue4(0, own + nb, nb == 0);
            } else {
              if (tp_elect_one()) mbar_arrive_cluster_relaxed(pl0 + 8u * (uint32_t)nb);
            }
            __syncwarp();
            // the partner's atom nb of h_t
            mbar_wait(h_in(nb), par);
            if (tp_elect_one()) {
              mbar_arrive_expect_tx(h_in(nb), HR_ATOM);
              mbar_arrive_cluster_relaxed(ack0 + 8u * (uint32_t)nb);
              if (!leader) mbar_arrive_cluster_relaxed(pi0 + 8u * (uint32_t)nb);
            }
            __syncwarp();
            if (leader) {
              mbar_wait_cluster(peer_in(nb), par);
              tc_fence_after();
              if (more && tp_elect_one()) {
                if (nb == 0) issue4(0, par_atom, false);
                else issue_tail();
              }
              __syncwarp();
            }
          }
        }
      }
      __syncthreads();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
    // ---------------- epilogue: thread = (window row, 32 of the current block's 64 hidden units) ----------------
    const int quarter = warp & 3, wq = warp >> 2;
    const int r = quarter * 32 + lane;
    int g0 = 0;
    for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
      const int dir = w / tile_pairs, tp = w - dir * tile_pairs;
      const int b0 = (2 * tp + s) * HR_M;
      item_begin(dir, g0 == 0);
      const bool live = b0 + r < Bc;
      float2 c[32];   // [block][16 (even, odd) unit pairs]: fp32 cell state
#pragma unroll
      for (int i = 0; i < 32; ++i) c[i] = make_float2(0.f, 0.f);
      for (int st = 0; st < T; ++st) {
        const int g = g0 + st;
        const int t = dir ? (T - 1 - st) : st;
        // rows of windows beyond the batch (partial or absent tiles) read a valid row instead; their results are never stored
        const long long row = (long long)t * Bc + (live ? b0 + r : (b0 < Bc ? b0 : 0));
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          // chunks (slab*4 + gate) of units 128 p + 64 nb + 8 slab .. +7, this thread's slabs wq*4 .. wq*4 + 3
          const uint8_t* gp = g_block(row) + (long long)(dir * 128 + p * 64 + nb * 32 + wq * 16) * 2048 + (row & 127) * 16ll;
          uint4 gbuf[3][4];
#pragma unroll
          for (int q = 0; q < 4; ++q) gbuf[0][q] = ldg_stream_v4(gp + q * 2048);
#pragma unroll
          for (int q = 0; q < 4; ++q) gbuf[1][q] = ldg_stream_v4(gp + (4 + q) * 2048);
          if (lane == 0) jit();
          __syncwarp();
          mbar_wait(acc_full(nb), (uint32_t)(g & 1));
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(nb * 256 + wq * 128);
          uint8_t* hrow = genH + (2 * p + nb) * HR_ATOM;   // own atom nb
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            uint32_t acc[32];
            tmem_ld32(taddr + sl * 32, acc);
            if (sl < 2) {
#pragma unroll
              for (int q = 0; q < 4; ++q) gbuf[(sl + 2) % 3][q] = ldg_stream_v4(gp + ((sl + 2) * 4 + q) * 2048);
            }
            tmem_ld_wait();
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(gbuf[sl % 3]);
            // packed fp32 (FADD2 / FFMA2 / FMUL2, bit-identical to the scalar form): units in (even, odd) pairs -- one 32-bit word of
            // G holds the bf16 pre-activations of such a pair, the accumulator registers of the pair are adjacent
            auto pre2 = [&](int gate, int u2) {
              const uint32_t wv = gw[gate * 4 + u2];
              return __fadd2_rn(make_float2(__uint_as_float(acc[gate * 8 + 2 * u2]), __uint_as_float(acc[gate * 8 + 2 * u2 + 1])),
                                make_float2(__uint_as_float(wv << 16), __uint_as_float(wv & 0xFFFF0000u)));
            };
            const float2 half2v = make_float2(0.5f, 0.5f);
            uint32_t hp[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; ++u2) {
              const float2 pi = pre2(0, u2), pf = pre2(1, u2), pg = pre2(2, u2), po = pre2(3, u2);
              const float2 ig = __ffma2_rn(half2v, make_float2(hr_tanh(pi.x), hr_tanh(pi.y)), half2v);
              const float2 fg = __ffma2_rn(half2v, make_float2(hr_tanh(pf.x), hr_tanh(pf.y)), half2v);
              const float2 gg = make_float2(hr_tanh(pg.x), hr_tanh(pg.y));
              const float2 og = __ffma2_rn(half2v, make_float2(hr_tanh(po.x), hr_tanh(po.y)), half2v);
              float2& cc = c[nb * 16 + sl * 4 + u2];
              cc = __ffma2_rn(fg, cc, __fmul2_rn(ig, gg));
              const float2 hv = __fmul2_rn(og, make_float2(hr_tanh(cc.x), hr_tanh(cc.y)));
              __nv_bfloat162 pk = __floats2bfloat162_rn(hv.x, hv.y);
              hp[u2] = *reinterpret_cast<uint32_t*>(&pk);
            }
            if (sl == 0) {   // own atom nb still holds h_{g-1}: see "who may overwrite what"
              if (g > 0) {
                mbar_wait(st_free(nb), (uint32_t)((g - 1) & 1));
                mbar_wait_cluster(copy_done(nb), (uint32_t)((g - 1) & 1));
              }
              if (nb == 0) mbar_wait(a0_free, (uint32_t)(g & 1));
            }
            *reinterpret_cast<uint4*>(hrow + sw128_chunk_off((uint32_t)r, (uint32_t)(wq * 4 + sl))) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_local(nb));
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == HR_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int h256_setup(int* max_clusters_out) {
  static PerDeviceInt state_pd, max_pd;  // state: 0 = not tried, 1 = ok, -1 = unavailable
  int& state = state_pd.cur();
  int& max_clusters = max_pd.cur();
  if (state == 0) {
    state = -1;
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec256_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HR_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * sm_count(), 1, 1);
    cfg.blockDim = dim3(HR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = HR_SMEM;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = 4; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    cfg.attrs = la; cfg.numAttrs = 1;
    BCI_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, lstm_rec256_bf16, &cfg));
    BCI_REQUIRE(max_clusters > 0, BCI_ECUDA, "H=256 recurrence: no 4-CTA cluster fits on this device");
    state = 1;
  }
  if (max_clusters_out) *max_clusters_out = max_clusters;
  return state == 1 ? BCI_OK : BCI_ECUDA;
}

int h256_max_clusters() {
  int n = 0;
  return h256_setup(&n) == BCI_OK ? n : 0;
}

int launch_rec256_bf16(const __nv_bfloat16* G, const __nv_bfloat16* whh, __nv_bfloat16* out, int Bc, int T, cudaStream_t st) {
  int max_clusters = 0;
  int rc = h256_setup(&max_clusters);
  if (rc) return rc;
  CUtensorMap tmOut;
  rc = make_tmap_bf16_3d(&tmOut, out, (uint64_t)T, (uint64_t)Bc, 512, 64, HR_M);
  if (rc) return rc;
  const int tiles = ceil_div(Bc, HR_M), tile_pairs = (tiles + 1) / 2;
  const int clusters = 2 * tile_pairs < max_clusters ? 2 * tile_pairs : max_clusters;
  static const int jitter = [] { const char* e = getenv("BCI_FUSED_JITTER"); int v = e ? atoi(e) : 0; return (v > 0 && (v & (v - 1)) == 0) ? v : 0; }();
  // BCI_H256_PIPE=0 keeps the unstaggered kernel
  static const bool pipe = [] { const char* e = getenv("BCI_H256_PIPE"); return !(e && e[0] == '0'); }();
  if (pipe) {
    static PerDeviceFlag attr_pd;
    bool& attr = attr_pd.cur();
    if (!attr) {
      BCI_CUDA_OK(cudaFuncSetAttribute(lstm_rec256_bf16_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HR_SMEM));
      attr = true;
    }
    lstm_rec256_bf16_pipe<<<4 * clusters, HP_THREADS, HR_SMEM, st>>>(G, tmOut, whh, Bc, T, tile_pairs, jitter);
    BCI_LAUNCH_OK();
    return BCI_OK;
  }
  lstm_rec256_bf16<<<4 * clusters, HR_THREADS, HR_SMEM, st>>>(G, tmOut, whh, Bc, T, tile_pairs, jitter);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

// ---- K1 for H = 256: x (B,T,C) fp32 -> bf16 rows [T][Bc][64] -> tcgen05 GEMM (+ b0) -> LayerNorm + GELU row kernel ----------------
__global__ void x_to_bf16_rows(const InputView x, int Bc, int T, int C, __nv_bfloat16* __restrict__ xr) {
  // one thread per (row, pair of channels); rows are time-major r = t*Bc + b, K padded from C to 64 with zeros
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)Bc * T;
  if (i >= rows * 32) return;
  const long long r = i >> 5;
  const int k = (int)(i & 31) * 2;
  const int t = (int)(r / Bc), b = (int)(r - (long long)t * Bc);
  const long long src = x.elem_off(b) + (long long)t * C;
  const float v0 = k < C ? view_load(x, src + k) : 0.f, v1 = (k + 1) < C ? view_load(x, src + k + 1) : 0.f;
  reinterpret_cast<__nv_bfloat162*>(xr)[i] = __floats2bfloat162_rn(v0, v1);
}

// z = GELU(LayerNorm(pre)) in place over bf16 rows of 256 (one warp per row, 8 columns per lane); tanh-form GELU as in the
// H = 128 kernel (lstm_bf16_inproj.cu)
__global__ void __launch_bounds__(256)
ln_gelu_rows256(__nv_bfloat16* __restrict__ z, long long rows, const float* __restrict__ lnw, const float* __restrict__ lnb) {
  const int lane = threadIdx.x & 31;
  float gw[8], gb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { gw[i] = lnw[lane * 8 + i]; gb[i] = lnb[lane * 8 + i]; }
  const long long wstride = (long long)gridDim.x * 8;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += wstride) {
    uint4* p = reinterpret_cast<uint4*>(z + r * 256) + lane;
    const uint4 v = *p;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      sm += f[2 * i] + f[2 * i + 1];
      sq = fmaf(f[2 * i], f[2 * i], fmaf(f[2 * i + 1], f[2 * i + 1], sq));
    }
    sm = warp_sum(sm);
    sq = warp_sum(sq);
    const float mean = sm * (1.0f / 256.0f);
    const float rstd = 1.0f / sqrtf(fmaxf(sq * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-5f);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float y[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float yy = fmaf((f[2 * i + e] - mean) * rstd, gw[2 * i + e], gb[2 * i + e]);
        const float u = yy * fmaf(0.0356774081f, yy * yy, 0.7978845608f);
        const float hy = 0.5f * yy;
        y[e] = fmaf(hy, hr_tanh(u), hy);
      }
      __nv_bfloat162 pk = __floats2bfloat162_rn(y[0], y[1]);
      o[i] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *p = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void pack_w0_256(const float* __restrict__ w0, __nv_bfloat16* __restrict__ dst, int C) {  // (256, C) -> [256][64] bf16, K zero-padded
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 64) return;
  const int j = i >> 6, k = i & 63;
  dst[i] = __float2bfloat16_rn(k < C ? w0[j * C + k] : 0.f);
}

static int launch_input_proj256(bci_lstm_s* h, const InputView& x, int Bc, int T, __nv_bfloat16* xr, __nv_bfloat16* z, cudaStream_t st) {
  const long long rows = (long long)Bc * T;
  x_to_bf16_rows<<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(x, Bc, T, h->cfg.input_size, xr);
  BCI_LAUNCH_OK();
  int rc = launch_proj_gemm_bf16(xr, h->bf16.w0_bf, h->f32.b0, z, (int)rows, 256, 64, false, st);
  if (rc) return rc;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148ll * 16) blocks = 148ll * 16;
  ln_gelu_rows256<<<(unsigned)blocks, 256, 0, st>>>(z, rows, h->f32.ln0w, h->f32.ln0b);
  BCI_LAUNCH_OK();
  return BCI_OK;
}

int pack_h256_bf16(bci_lstm_s* h, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const bci_lstm_weights& w = h->raw;
  for (int l = 0; l < c.num_layers; ++l) {
    const int K = layer_in_width(c, l);
    for (int d = 0; d < 2; ++d) {
      pack_rows_perm256_bf16<<<(unsigned)ceil_div64((long long)1024 * K, 256), 256, 0, st>>>(w.w_ih[l][d], h->bf16.wih256[l], K, d * 1024);
      pack_rows_perm256_bf16<<<(unsigned)ceil_div64((long long)1024 * 256, 256), 256, 0, st>>>(w.w_hh[l][d], h->bf16.whh256[l], 256, d * 1024);
      pack_bias_perm256<<<4, 256, 0, st>>>(w.b_ih[l][d], w.b_hh[l][d], h->bf16.bias256[l], d * 1024);
    }
  }
  pack_w0_256<<<64, 256, 0, st>>>(w.input_proj_w, h->bf16.w0_bf, c.input_size);
  BCI_LAUNCH_OK();
  return pack_pool256_bf16(h, st);
}

static size_t chunk_bytes_h256(const bci_lstm_config& c, int Bc, int T) {
  const size_t rows = (size_t)Bc * T, rows_pad = (rows + 127) / 128 * 128;
  // z, G (also reused as the [rows][256] bf16 score pre-activations after the last layer), two outputs, scores, row statistics
  return align_up(rows * 256 * 2, 1024) + align_up(rows_pad * 2048 * 2, 1024) + 2 * align_up(rows * 512 * 2, 1024) + align_up(rows * 4, 1024) +
         align_up(rows * 8, 1024);
}

size_t lstm_workspace_h256(const bci_lstm_config& c, int batch, int T) {
  const int Bc = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  return chunk_bytes_h256(c, Bc > 0 ? Bc : 1, T) + 1024;
}

// every contraction runs on tensor cores: K1 = bf16 row conversion + proj_gemm_bf16 (K = 64) + LayerNorm/GELU row kernel
int lstm_forward_h256(bci_lstm_s* h, const InputView& x, int batch, int T, float* logits, float* probs, float* attn, void* ws,
                      size_t ws_bytes, cudaStream_t st) {
  const bci_lstm_config& c = h->cfg;
  const int chunk = batch < max_chunk(c, 0) ? batch : max_chunk(c, 0);
  char* ws_al = reinterpret_cast<char*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  BCI_REQUIRE(ws_bytes >= chunk_bytes_h256(c, chunk, T) + (size_t)(ws_al - (char*)ws), BCI_ENOMEM,
              "bci_lstm_forward: workspace %zu < %zu bytes", ws_bytes, chunk_bytes_h256(c, chunk, T) + 1024);
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int Bc = (batch - b0) < chunk ? (batch - b0) : chunk;
    const size_t rows = (size_t)Bc * T;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = ws_al + off; off += align_up(bytes, 1024); return p; };
    __nv_bfloat16* z = reinterpret_cast<__nv_bfloat16*>(take(rows * 256 * 2));
    __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(take((rows + 127) / 128 * 128 * 2048 * 2));
    __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(take(rows * 512 * 2));
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(take(rows * 512 * 2));
    float* scores = reinterpret_cast<float*>(take(rows * 4));
    float2* rowstat = reinterpret_cast<float2*>(take(rows * 8));
    h->prof.mark(-1, st);
    // (x as bf16 rows is staged in the second output buffer, which is not written before layer 1)
    int rc = launch_input_proj256(h, chunk_view(x, b0), Bc, T, o1, z, st);
    if (rc) return rc;
    h->prof.mark(0, st);
    const __nv_bfloat16* in = z;
    __nv_bfloat16* outs[2] = {o0, o1};
    for (int l = 0; l < c.num_layers; ++l) {
      rc = launch_proj_gemm_bf16(in, h->bf16.wih256[l], h->bf16.bias256[l], g, (int)rows, 2048, layer_in_width(c, l), true, st);
      if (rc) return rc;
      h->prof.mark(1, st);
      rc = launch_rec256_bf16(g, h->bf16.whh256[l], outs[l & 1], Bc, T, st);
      if (rc) return rc;
      h->prof.mark(2, st);
      in = outs[l & 1];
    }
    // pooling: score GEMM on tensor cores (G's buffer is free now and holds the bf16 pre-activations), LayerNorm folded in
    rc = launch_pool256_bf16(h, in, g, rowstat, scores, Bc, T, logits + (size_t)b0 * c.num_classes,
                             probs ? probs + (size_t)b0 * c.num_classes : nullptr, attn ? attn + (size_t)b0 * T : nullptr, st);
    if (rc) return rc;
    h->prof.mark(3, st);
  }
  return BCI_OK;
}

}  // namespace bci

// diagnostics (tests/test_gpu_tensorcore.py): the H = 256 cluster recurrence in isolation
extern "C" int bci_selftest_rec256_bf16(const void* G, const void* whh, void* out, int32_t Bc, int32_t T, void* stream) {
  return bci::launch_rec256_bf16((const __nv_bfloat16*)G, (const __nv_bfloat16*)whh, (__nv_bfloat16*)out, Bc, T, (cudaStream_t)stream);
}
