// Fused projection + recurrence of one bidirectional LSTM layer (bf16 mode, H = 128) on a 4-CTA cluster.
//
// The two-kernel path (lstm_bf16.cu: K2 proj_gemm_bf16 -> G in HBM -> K3 lstm_rec_bf16) is bound by the 512 KB per window per
// layer of G it writes and reads back (ncu: K2 DRAM 69 %, K3 DRAM 58 %).  Here G never exists: per step the gate
// pre-activations  [in_t | h_{t-1}] . [W_ih | W_hh]^T  are accumulated in TMEM by one chain of tcgen05 MMAs, with BOTH weight
// matrices of the direction (384 KB bf16 for a 256-wide input) resident in the shared memory of four SMs:
//
//   cluster rank r = 2 p + s      p = gate-column half (hidden units [64 p, 64 p + 64), all four gates: 256 columns)
//                                 s = window-tile column of the cluster's tile quad = position in the CTA pair
//   pair p = CTAs (p,0),(p,1)     one tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) per K slice; each CTA keeps 128 of the
//                                 pair's 256 weight rows (96 KB), supplies its own tiles' A operand ([in_t | h_{t-1}]) and
//                                 receives its own 128 x 256 accumulators in its TMEM
//   two tiles per CTA (q = 0,1)   PING-PONG: while the 8 epilogue warps work on tile q (tcgen05.ld, + bias, sigma/tanh on the
//                                 MUFU, cell update in fp32 registers, h_t -> bf16 operand buffer), the tensor pipe runs
//                                 MMA_ih + MMA_hh of tile 1-q and the h exchange of tile 1-q is in flight.  Both pipes stay busy:
//                                 per tile-step 3 072 tensor cycles against ~3 300 epilogue cycles.  (One tile per CTA left
//                                 MMA latency + exchange, ~1 350 ns of 3 200 ns per step, exposed: profiles/r1_fused_timeline.md.)
//   h exchange (DSMEM)            each CTA computes h_t for its 64 units = one K-atom (16 KB contiguous) of the operand buffer of
//                                 step t+1.  When the atom is complete one thread sends it to the CTA (1-p, s) that computes the
//                                 other gates of the same windows: a single cp.async.bulk shared::cta -> shared::cluster whose
//                                 bytes complete an mbarrier there (per-thread st.shared::cluster + proxy/cluster fences cost
//                                 2 300 cycles per step in the first version).  The pair leader issues MMA_hh once both CTAs of
//                                 its pair hold the complete h: two local barriers plus RELAXED relay arrives from its peer.
//   buffers                       W 6 atoms (96 KB) + h 2 tiles x 2 atoms (64 KB, single-buffered: the ping-pong schedule separates
//                                 readers and writers) + 4-slot TMA ring for in_t (64 KB) + bias; TMEM 2 tiles x 256 columns.
//   h_t -> HBM                    each CTA TMA-stores its own 64-unit atom of h_t.
//   epilogue arithmetic           (session 5, profiles/r5_fused_epilogue_stalls.md) the bias of slab s+1 is requested from shared memory while
//                                 slab s is evaluated, and the fp32 operations run in packed form (FADD2 / FFMA2 / FMUL2 on (even, odd)
//                                 unit pairs): per thread-step 320 MUFU.TANH + 320 packed FP instructions instead of 640 scalar ones.
//
// Reference semantics: nn.LSTM inside EnhancedLSTMModel (04_lstm_model.py:181-188,211).
#include "lstm_shared_kernels.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int FR_M = 128;                       // windows per tile
constexpr int FR_EPI_WARPS = 8;
constexpr int FR_THREADS = (FR_EPI_WARPS + 3) * 32;  // + MMA issuer / relay, TMA producer, h store warp
constexpr uint32_t FR_ATOM = 128 * 128;         // [128 rows][64 bf16] SW128 atom
constexpr int FR_STAGES = 4;
constexpr int FR_W_ATOMS = 6;                   // up to 4 K-atoms of W_ih (input width 256) + 2 of W_hh
constexpr uint32_t FR_OFF_H = FR_W_ATOMS * FR_ATOM;              // h of tile 0 and tile 1, two atoms each
constexpr uint32_t FR_OFF_RING = FR_OFF_H + 4 * FR_ATOM;
constexpr uint32_t FR_OFF_BIAS = FR_OFF_RING + FR_STAGES * FR_ATOM;  // 256 floats
constexpr uint32_t FR_OFF_CTL = FR_OFF_BIAS + 1024;
constexpr size_t FR_SMEM = 1024 + FR_OFF_CTL + 256;

__device__ __forceinline__ float fr_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// work item w (0 .. 2*tile_quads): direction = w / tile_quads, tile quad = w % tile_quads; CTA (p, s) owns tiles 4 quad + 2 q + s
template <bool STATS>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(FR_THREADS, 1)
lstm_fused_bf16(const __grid_constant__ CUtensorMap tmIn,   // in  [T][Bc][Kin] bf16, box 64 x 128 x 1
                const __grid_constant__ CUtensorMap tmOut,  // out [T][Bc][256] bf16, box 64 x 128 x 1
                const __nv_bfloat16* __restrict__ wih,      // [2][512][Kin] rows in perm_T order (i,f,o rows pre-scaled by 1/2)
                const __nv_bfloat16* __restrict__ whh_f,    // [512][128] perm_T rows, forward
                const __nv_bfloat16* __restrict__ whh_r,    // reverse
                const float* __restrict__ bias,             // [2][512] perm_T order, pre-scaled like the rows
                float2* __restrict__ stats,                 // STATS: [T][8][Bc] (sum, sumsq) of h over 32 units, slot = dir*4 + p*2 + ch
                int Bc, int T, int Kin, int tile_quads, int jitter,  // jitter: 0 or a power of two (max sleep in ns)
                long long* __restrict__ tl) {               // optional timeline (BCI_FUSED_TIMELINE): cluster 0, 8 stamps x step x rank
  extern __shared__ uint8_t fr_smem_raw[];
  // BCI_FUSED_JITTER (tests only): every role sleeps a pseudo-random time at its synchronisation points, to shake out ordering
  // assumptions that only hold at the natural timing
  uint32_t jit_state = jitter ? (uint32_t)(blockIdx.x * 7919u + threadIdx.x * 104729u + 12345u) : 0u;
  auto jit = [&]() {
    if (jitter) {
      jit_state = jit_state * 1664525u + 1013904223u;
      __nanosleep((jit_state >> 20) & (uint32_t)(jitter - 1));
    }
  };
  const bool tl_on = tl != nullptr && blockIdx.x < 4;
  auto stamp = [&](int st, int slot) {
    if (tl_on && st < 64) { long long c; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(c)); tl[(blockIdx.x * 64 + st) * 8 + slot] = c; }
  };
  const uint32_t raw = smem_u32(fr_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = fr_smem_raw + (base - raw);
  const uint32_t sW = base, sH = base + FR_OFF_H, sRing = base + FR_OFF_RING;
  uint8_t* genH = gen + FR_OFF_H;
  float* bias_s = reinterpret_cast<float*>(gen + FR_OFF_BIAS);
  uint8_t* ctl = gen + FR_OFF_CTL;
  const uint32_t bar0 = smem_u32(ctl);
  // per-tile barriers complete once per step g: parity g & 1
  auto in_full = [&](int i) { return bar0 + 8u * i; };            // leader: 1 arrival (expect_tx), bytes of both CTAs
  auto in_empty = [&](int i) { return bar0 + 8u * (4 + i); };     // every CTA: multicast commit
  auto acc_full = [&](int q) { return bar0 + 8u * (8 + q); };     // every CTA: multicast commit (tile q's accumulator is ready)
  auto h_local = [&](int q) { return bar0 + 8u * (10 + q); };     // every CTA: its 8 epilogue warps wrote h (and drained TMEM)
  auto h_in = [&](int q) { return bar0 + 8u * (12 + q); };        // every CTA: the partner's atom landed (expect_tx by the store warp)
  auto st_free = [&](int q) { return bar0 + 8u * (14 + q); };     // every CTA: TMA store of the local atom finished reading it
  auto peer_local = [&](int q) { return bar0 + 8u * (16 + q); };  // leader: relay of the peer's h_local
  auto peer_in = [&](int q) { return bar0 + 8u * (18 + q); };     // leader: relay of the peer's h_in
  auto copy_done = [&](int q) { return bar0 + 8u * (20 + q); };   // every CTA: ack -- my outgoing atom landed at the partner
  auto recv_ready = [&](int q) { return bar0 + 8u * (22 + q); };  // every CTA: the partner's MMA_hh(g) retired: its copy of h_{g-1}
                                                                   // may be overwritten by my atom of h_g (h is single-buffered)
  auto acc_free = [&](int q) { return bar0 + 8u * (24 + q); };    // every CTA: its 8 epilogue warps have read the whole accumulator
  auto peer_free = [&](int q) { return bar0 + 8u * (26 + q); };   // leader: relay of the peer's acc_free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * 28);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int p = (int)(rank >> 1), s = (int)(rank & 1);
  const bool leader = (s == 0);
  const uint32_t partner = (uint32_t)(2 * (1 - p) + s);  // same windows, other half of the gates
  const int nk = Kin / 64;  // K-atoms of the input part

  if (tid == 0) {
    for (int i = 0; i < FR_STAGES; ++i) { mbar_init(in_full(i), 1); mbar_init(in_empty(i), 1); }
    for (int q = 0; q < 2; ++q) {
      mbar_init(acc_full(q), 1);
      mbar_init(h_local(q), FR_EPI_WARPS);
      mbar_init(h_in(q), 1);
      mbar_init(st_free(q), 1);
      mbar_init(peer_local(q), 1);
      mbar_init(peer_in(q), 1);
      mbar_init(copy_done(q), 1);
      mbar_init(recv_ready(q), 1);
      mbar_init(acc_free(q), FR_EPI_WARPS);
      mbar_init(peer_free(q), 1);
    }
    fence_mbar_init();
    // h_in is armed for the first step here and re-armed by its (single) waiter after every completed phase: arming it from
    // the store warp let a second expect_tx arrive land on a phase whose bytes were still in flight (arrival-count
    // underflow -> launch failure as soon as the partner CTA ran a little late; found with BCI_FUSED_JITTER and under ncu)
    for (int q = 0; q < 2; ++q) mbar_arrive_expect_tx(h_in(q), FR_ATOM);
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == FR_EPI_WARPS) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();  // every CTA's barriers are initialised before anyone signals them

  const int n_work = 2 * tile_quads;
  const int n_clusters = (int)cluster_nclusters_x();
  // running step counter over all work items of this cluster (mbarrier parities are functions of it)
  int g0 = 0;
  for (int w = (int)cluster_id_x(); w < n_work; w += n_clusters, g0 += T) {
    const int dir = w / tile_quads, tq = w - dir * tile_quads;
    const int b0q[2] = {(4 * tq + s) * FR_M, (4 * tq + 2 + s) * FR_M};  // first window of this CTA's tiles q = 0, 1

    // ---- (re)load this CTA's weight rows: perm_T rows [256 p + 128 s, +128) of direction `dir` ----
    if (g0 > 0) cluster_sync_all();  // every role finished the previous item (its last step was waited on below)
    {
      const int row0 = 256 * p + 128 * s;
      const uint4* src_ih = reinterpret_cast<const uint4*>(wih + ((size_t)dir * 512 + row0) * Kin);
      const int cpr = Kin / 8;  // 16-byte chunks per row
      for (int i = tid; i < 128 * cpr; i += FR_THREADS) {
        const uint32_t row = i / cpr, cc = i - row * cpr, atom = cc >> 3, c = cc & 7;
        *reinterpret_cast<uint4*>(gen + atom * FR_ATOM + sw128_chunk_off(row, c)) = __ldg(src_ih + i);
      }
      const uint4* src_hh = reinterpret_cast<const uint4*>((dir ? whh_r : whh_f) + (size_t)row0 * 128);
      for (int i = tid; i < 128 * 16; i += FR_THREADS) {
        const uint32_t row = i >> 4, cc = i & 15, atom = cc >> 3, c = cc & 7;
        *reinterpret_cast<uint4*>(gen + (nk + atom) * FR_ATOM + sw128_chunk_off(row, c)) = __ldg(src_hh + i);
      }
      // bias of the pair's 256 gate columns (every CTA of the pair needs all of them for its own rows)
      for (int i = tid; i < 256; i += FR_THREADS) bias_s[i] = __ldg(bias + dir * 512 + 256 * p + i);
    }
    fence_proxy_async_all();
    __syncthreads();
    cluster_sync_all();

    if (warp == FR_EPI_WARPS + 1) {
      // ---------------- TMA producer: in_t of tile 0, then tile 1, K-atom by K-atom, through the ring ----------------
      if (lane == 0) {
        int k_total = g0 * 2 * nk;  // ring slots consumed before this item (same count in every CTA)
        const uint32_t lead_full0 = mapa_u32(in_full(0), rank & ~1u);
        for (int st = 0; st < T; ++st) {
          const int t = dir ? (T - 1 - st) : st;
          // the ring holds one tile-step: pull the tiles of step st+2 into L2 now so that their smem loads, which can only be
          // issued once earlier MMAs have freed the slots, are L2 hits (one of the two CTAs sharing the tiles does it)
          if (p == 0 && st + 2 < T) {
            const int t2 = dir ? (T - 3 - st) : st + 2;
            for (int q = 0; q < 2; ++q)
              for (int k = 0; k < nk; ++k) tma_prefetch_l2_3d(&tmIn, k * 64, b0q[q], t2);
          }
          for (int q = 0; q < 2; ++q) {
            for (int k = 0; k < nk; ++k, ++k_total) {
              const int stage = k_total % FR_STAGES;
              const uint32_t ph = (uint32_t)((k_total / FR_STAGES) & 1);
              jit();
              mbar_wait(in_empty(stage), ph ^ 1u);
              if (leader) mbar_arrive_expect_tx(in_full(stage), 2 * FR_ATOM);
              tma_load_3d_2sm(sRing + stage * FR_ATOM, &tmIn, k * 64, b0q[q], t, lead_full0 + 8u * stage);
            }
          }
        }
      }
    } else if (warp == FR_EPI_WARPS) {
      if (leader && lane == 0) {
        // ---------------- MMA issuer (pair leader) ----------------
        constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
        const uint16_t pair_mask = (uint16_t)(3u << (2 * p));
        const uint32_t ack0 = mapa_u32(copy_done(0), partner);
        int k_total = g0 * 2 * nk;
        // accumulator of tile q drained by both CTAs of the pair: signalled after the LAST tcgen05.ld of their epilogues of
        // step gp, i.e. ~1/4 of an epilogue before h is complete, so MMA_ih of the next step starts that much earlier (with the
        // single-buffered accumulator, waiting for the whole epilogue left the epilogue warps idle ~0.3 us per tile-step)
        auto wait_drained = [&](int q, int gp) {
          jit();
          mbar_wait(acc_free(q), (uint32_t)(gp & 1));
          mbar_wait_cluster(peer_free(q), (uint32_t)(gp & 1));
          tc_fence_after();
        };
        // h of step gp complete in both CTAs of the pair (both K-atoms: written locally and landed from the partner)
        auto wait_exchanged = [&](int q, int gp) {
          jit();
          mbar_wait(h_local(q), (uint32_t)(gp & 1));
          mbar_wait_cluster(peer_local(q), (uint32_t)(gp & 1));
          mbar_wait(h_in(q), (uint32_t)(gp & 1));
          mbar_arrive_expect_tx(h_in(q), FR_ATOM);     // re-arm for the next step's incoming atom
          mbar_arrive_cluster_relaxed(ack0 + 8u * q);  // tell the sender its outgoing copy has landed
          mbar_wait_cluster(peer_in(q), (uint32_t)(gp & 1));
          tc_fence_after();
        };
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (st > 0) wait_drained(q, g - 1);
            if (q == 0) stamp(st, 0);
            // ---- MMA_ih: acc_q = in_t(tile q) . W_ih^T ----
            for (int k = 0; k < nk; ++k, ++k_total) {
              const int stage = k_total % FR_STAGES;
              mbar_wait(in_full(stage), (uint32_t)((k_total / FR_STAGES) & 1));
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t da = umma_desc_sw128(sRing + stage * FR_ATOM + kk * 32);
                const uint64_t db = umma_desc_sw128(sW + k * FR_ATOM + kk * 32);
                umma_bf16_2sm(tmem_base + q * 256, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
              }
              umma_commit_2sm_mc(in_empty(stage), pair_mask);  // frees the slot in both CTAs when these MMAs retire
            }
            if (q == 0) stamp(st, 1);
            // ---- MMA_hh: acc_q += h_{t-1}(tile q) . W_hh^T ----
            if (st > 0) {
              wait_exchanged(q, g - 1);
              if (q == 0) stamp(st, 2);
              const uint32_t hq = sH + (uint32_t)q * 2 * FR_ATOM;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t atom = k >> 2, kk = k & 3;
                const uint64_t da = umma_desc_sw128(hq + atom * FR_ATOM + kk * 32);
                const uint64_t db = umma_desc_sw128(sW + (nk + atom) * FR_ATOM + kk * 32);
                umma_bf16_2sm(tmem_base + q * 256, da, db, idesc, 1u);
              }
            }
            umma_commit_2sm_mc(acc_full(q), pair_mask);
          }
        }
        // the item's last step must be complete everywhere before the weights are replaced / the kernel ends
        for (int q = 0; q < 2; ++q) {
          wait_drained(q, g0 + T - 1);
          wait_exchanged(q, g0 + T - 1);
        }
      } else if (!leader && lane == 0) {
        // ---------------- relay (peer CTA of the pair) ----------------
        // RELAXED arrives: this thread has written nothing the leader needs -- the data sits in this CTA's shared memory (written
        // by the bulk copy, whose completion the h_in wait observes, and by the local epilogue warps, who fenced generic->async
        // before arriving on h_local); a release arrive costs a MEMBAR.GPU = 640 ns per step (measured) on the critical path.
        const uint32_t pl0 = mapa_u32(peer_local(0), rank & ~1u), pi0 = mapa_u32(peer_in(0), rank & ~1u);
        const uint32_t pf0 = mapa_u32(peer_free(0), rank & ~1u);
        const uint32_t ack0 = mapa_u32(copy_done(0), partner);
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
          for (int q = 0; q < 2; ++q) {
            jit();
            mbar_wait(acc_free(q), (uint32_t)(g & 1));
            mbar_arrive_cluster_relaxed(pf0 + 8u * q);
            mbar_wait(h_local(q), (uint32_t)(g & 1));
            mbar_arrive_cluster_relaxed(pl0 + 8u * q);
            jit();
            mbar_wait(h_in(q), (uint32_t)(g & 1));
            mbar_arrive_expect_tx(h_in(q), FR_ATOM);  // re-arm for the next step's incoming atom
            mbar_arrive_cluster_relaxed(pi0 + 8u * q);
            mbar_arrive_cluster_relaxed(ack0 + 8u * q);
          }
        }
      }
    } else if (warp == FR_EPI_WARPS + 2) {
      // ---------------- h store warp: this CTA's 64-unit atom of h_t -> partner CTA (DSMEM) and -> out[t] ----------------
      if (lane == 0) {
        for (int st = 0; st < T; ++st) {
          const int g = g0 + st;
          const int t = dir ? (T - 1 - st) : st;
          for (int q = 0; q < 2; ++q) {
            const uint32_t atom = sH + (uint32_t)q * 2 * FR_ATOM + p * FR_ATOM;
            jit();
            mbar_wait(h_local(q), (uint32_t)(g & 1));
            if (q == 0) stamp(st, 5);
            mbar_wait_cluster(recv_ready(q), (uint32_t)(g & 1));  // the partner pair has finished reading h_{g-1} of this tile
            bulk_copy_s2s_cluster(mapa_u32(atom, partner), atom, FR_ATOM, mapa_u32(h_in(q), partner));
            tma_store_3d(&tmOut, atom, dir * 128 + 64 * p, b0q[q], t);
            tma_store_commit();
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the PREVIOUS store has finished reading its atom
            if (st > 0 || q > 0) mbar_arrive(st_free(q ^ 1));
          }
        }
        tma_store_wait_all();
        mbar_arrive(st_free(1));
      }
    } else {
      // ---------------- epilogue: thread = (window row, 32 of the CTA's 64 hidden units), tiles q = 0, 1 alternately ----------------
      const int quarter = warp & 3, ch = warp >> 2;
      const int r = quarter * 32 + lane;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ch * 128;
      const float4* bias4 = reinterpret_cast<const float4*>(bias_s) + ch * 32;
      float2 c[2][16];  // fp32 cell state of the thread's 32 units per tile, as (even, odd) unit pairs
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < 16; ++i) c[q][i] = make_float2(0.f, 0.f);
      float4 bq[8];  // bias of the next slab's 32 columns: warp-uniform 16-byte shared-memory loads (the same four slabs every tile-step)
#pragma unroll
      for (int i = 0; i < 8; ++i) bq[i] = bias4[i];

      for (int st = 0; st < T; ++st) {
        const int g = g0 + st;
        const int t = dir ? (T - 1 - st) : st;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint8_t* hloc = genH + (uint32_t)q * 2 * FR_ATOM + p * FR_ATOM;
          if (lane == 0) jit();
          __syncwarp();
          mbar_wait(acc_full(q), (uint32_t)(g & 1));
          tc_fence_after();
          // MMA_hh(g) of this pair has retired: the partner may now send its atom of h_g into this CTA's (single) h buffer
          if (tid == 0) mbar_arrive_cluster_relaxed(mapa_u32(recv_ready(q), partner));
          if (tid == 0 && q == 0) stamp(st, 3);  // accumulator ready
          float ssum = 0.f, ssq = 0.f;
          uint32_t acc[2][32];
          tmem_ld32(taddr0 + q * 256, acc[0]);
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            tmem_ld_wait();
            if (sl + 1 < 4) tmem_ld32(taddr0 + q * 256 + (sl + 1) * 32, acc[(sl + 1) & 1]);
            if (sl == 3) {  // the accumulator is in registers: MMA_ih of the next step may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_free(q));
            }
            const uint32_t* a = acc[sl & 1];
            // pre-activations = accumulator + bias; bq holds the slab's 32 bias values (gate*8 + u), loaded one slab AHEAD: ncu
            // (profiles/r5_fused_epilogue_stalls.md) showed the first FADDs of every slab waiting on their just-issued LDS (short
            // scoreboard, 12 % of the epilogue warps' samples), so the next slab's values are requested as soon as these adds have
            // consumed the current ones and arrive under the slab's MUFU work
            // Packed fp32 (FADD2 / FFMA2 / FMUL2: two independent IEEE operations per issued instruction, so every lane computes
            // exactly what the scalar form did): units are handled in (even, odd) pairs -- the accumulator registers of a
            // 32-column tcgen05.ld and the .xy / .zw halves of the bias float4s are already aligned register pairs.
            float2 pre[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              pre[i * 2 + 0] = __fadd2_rn(make_float2(__uint_as_float(a[i * 4 + 0]), __uint_as_float(a[i * 4 + 1])), make_float2(bq[i].x, bq[i].y));
              pre[i * 2 + 1] = __fadd2_rn(make_float2(__uint_as_float(a[i * 4 + 2]), __uint_as_float(a[i * 4 + 3])), make_float2(bq[i].z, bq[i].w));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) bq[i] = bias4[((sl + 1) & 3) * 8 + i];
            const float2 half2v = make_float2(0.5f, 0.5f);
            uint32_t hp[4];
#pragma unroll
            for (int u2 = 0; u2 < 4; ++u2) {   // units 2*u2, 2*u2 + 1 of the slab: pre[gate * 4 + u2]
              const float2 pi = pre[0 * 4 + u2], pf = pre[1 * 4 + u2], pg = pre[2 * 4 + u2], po = pre[3 * 4 + u2];
              const float2 ig = __ffma2_rn(half2v, make_float2(fr_tanh(pi.x), fr_tanh(pi.y)), half2v);
              const float2 fg = __ffma2_rn(half2v, make_float2(fr_tanh(pf.x), fr_tanh(pf.y)), half2v);
              const float2 gg = make_float2(fr_tanh(pg.x), fr_tanh(pg.y));
              const float2 og = __ffma2_rn(half2v, make_float2(fr_tanh(po.x), fr_tanh(po.y)), half2v);
              float2& cc = c[q][sl * 4 + u2];
              cc = __ffma2_rn(fg, cc, __fmul2_rn(ig, gg));
              const float2 hv = __fmul2_rn(og, make_float2(fr_tanh(cc.x), fr_tanh(cc.y)));
              if (STATS) { ssum += hv.x; ssq = fmaf(hv.x, hv.x, ssq); ssum += hv.y; ssq = fmaf(hv.y, hv.y, ssq); }
              __nv_bfloat162 pk = __floats2bfloat162_rn(hv.x, hv.y);
              hp[u2] = *reinterpret_cast<uint32_t*>(&pk);
            }
            const uint4 hvec = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            // the local atom still held h_{g-1}: its TMA store and its copy to the partner must have finished reading it
            if (sl == 0 && g > 0) {
              mbar_wait(st_free(q), (uint32_t)((g - 1) & 1));
              mbar_wait_cluster(copy_done(q), (uint32_t)((g - 1) & 1));
            }
            *reinterpret_cast<uint4*>(hloc + sw128_chunk_off((uint32_t)r, (uint32_t)(ch * 4 + sl))) = hvec;
          }
          if (tid == 0 && q == 0) stamp(st, 4);  // math + stores issued
          fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma / the bulk copies
          tc_fence_before();         // order this thread's TMEM reads before the arrive
          __syncwarp();
          if (lane == 0) mbar_arrive(h_local(q));
          if (STATS) {
            // [T][8][Bc]: the warp's 32 windows are 256 contiguous bytes (the row-major [row][8] layout of the first version made
            // every 8-byte store a partial sector: ncu showed 2.3 GB of DRAM read-modify-write reads and a 35 % slower kernel)
            if (b0q[q] + r < Bc) stats[((long long)t * 8 + dir * 4 + p * 2 + ch) * Bc + b0q[q] + r] = make_float2(ssum, ssq);
          }
        }
      }
    }
    __syncthreads();
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while a peer may still write into its shared memory / wait on its MMAs
  if (warp == FR_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int fused_setup(int* max_clusters_out) {
  static PerDeviceInt state_pd, max_pd;  // state: 0 = not tried, 1 = ok, -1 = unavailable
  int& state = state_pd.cur();
  int& max_clusters = max_pd.cur();
  if (state == 0) {
    state = -1;
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_fused_bf16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM));
    BCI_CUDA_OK(cudaFuncSetAttribute(lstm_fused_bf16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * sm_count(), 1, 1);
    cfg.blockDim = dim3(FR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = FR_SMEM;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = 4; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    cfg.attrs = la; cfg.numAttrs = 1;
    BCI_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, lstm_fused_bf16<false>, &cfg));
    BCI_REQUIRE(max_clusters > 0, BCI_ECUDA, "fused recurrence: no 4-CTA cluster fits on this device");
    // BCI_FUSED_CLUSTERS=n caps the co-resident clusters used (experiment: on a 148-SM B200 32 clusters x 512 windows were 2 % faster per
    // window than 33 x 512 in one 5-step sweep and 4 % slower in a sustained bench run on another box: within run-to-run variation, so
    // the occupancy limit stays the default)
    const int occ = max_clusters;
    const char* e = getenv("BCI_FUSED_CLUSTERS");
    if (e && atoi(e) > 0) max_clusters = atoi(e) < occ ? atoi(e) : occ;
    state = 1;
  }
  if (max_clusters_out) *max_clusters_out = max_clusters;
  return state == 1 ? BCI_OK : BCI_ECUDA;
}

int fused_max_clusters() {
  int n = 0;
  return fused_setup(&n) == BCI_OK ? n : 0;
}

int launch_fused_rec_bf16(const __nv_bfloat16* in, const __nv_bfloat16* wih, const __nv_bfloat16* whh_f, const __nv_bfloat16* whh_r,
                          const float* bias, __nv_bfloat16* out, float2* stats, int Bc, int T, int Kin, cudaStream_t st) {
  BCI_REQUIRE(Kin == 128 || Kin == 256, BCI_EINVAL, "fused recurrence: input width must be 128 or 256 (got %d)", Kin);
  int max_clusters = 0;
  int rc0 = fused_setup(&max_clusters);
  if (rc0) return rc0;
  CUtensorMap tmIn, tmOut;
  int rc = make_tmap_bf16_3d(&tmIn, in, (uint64_t)T, (uint64_t)Bc, (uint64_t)Kin, 64, FR_M);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tmOut, out, (uint64_t)T, (uint64_t)Bc, 256, 64, FR_M);
  if (rc) return rc;
  const int tiles = ceil_div(Bc, FR_M), tile_quads = (tiles + 3) / 4;
  int clusters = 2 * tile_quads < max_clusters ? 2 * tile_quads : max_clusters;
  // BCI_FUSED_TIMELINE=<path>: clock64 stamps of the first 64 steps of cluster 0 / rank 0 are written to <path> (debug only)
#ifdef BCI_DEBUG_SWITCHES
  static const char* tl_path = getenv("BCI_FUSED_TIMELINE");
#else
  static const char* tl_path = nullptr;
#endif
  static long long* tl_dev = nullptr;
  if (tl_path && !tl_dev) { BCI_CUDA_OK(cudaMalloc(&tl_dev, 4 * 64 * 8 * sizeof(long long))); }
  if (tl_dev) BCI_CUDA_OK(cudaMemsetAsync(tl_dev, 0, 4 * 64 * 8 * sizeof(long long), st));
  static const int jitter = [] { const char* e = getenv("BCI_FUSED_JITTER"); int v = e ? atoi(e) : 0; return (v > 0 && (v & (v - 1)) == 0) ? v : 0; }();
  if (stats) lstm_fused_bf16<true><<<4 * clusters, FR_THREADS, FR_SMEM, st>>>(tmIn, tmOut, wih, whh_f, whh_r, bias, stats, Bc, T, Kin, tile_quads, jitter, tl_dev);
  else lstm_fused_bf16<false><<<4 * clusters, FR_THREADS, FR_SMEM, st>>>(tmIn, tmOut, wih, whh_f, whh_r, bias, nullptr, Bc, T, Kin, tile_quads, jitter, tl_dev);
  if (tl_dev) {
    long long host[4 * 64 * 8];
    BCI_CUDA_OK(cudaMemcpyAsync(host, tl_dev, sizeof(host), cudaMemcpyDeviceToHost, st));
    BCI_CUDA_OK(cudaStreamSynchronize(st));
    FILE* f = fopen(tl_path, "w");
    if (f) {
      for (int i = 0; i < 256; ++i) {
        for (int j = 0; j < 8; ++j) fprintf(f, "%lld ", host[i * 8 + j] ? host[i * 8 + j] - host[0] : 0ll);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci

// diagnostics: one fused layer in isolation (tests/test_gpu_tensorcore.py)
extern "C" int bci_selftest_fused_rec_bf16(const void* in, const void* wih, const void* whh_f, const void* whh_r, const float* bias,
                                           void* out, void* stats, int32_t Bc, int32_t T, int32_t Kin, void* stream) {
  return bci::launch_fused_rec_bf16((const __nv_bfloat16*)in, (const __nv_bfloat16*)wih, (const __nv_bfloat16*)whh_f,
                                    (const __nv_bfloat16*)whh_r, bias, (__nv_bfloat16*)out, (float2*)stats, Bc, T, Kin,
                                    (cudaStream_t)stream);
}
