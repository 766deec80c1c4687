// bf16 projection GEMM on CTA pairs:  C[M][N] = A[M][K] . W[N][K]^T + bias   (bf16 in, fp32 accumulate, bf16 out)
//
// The one-CTA kernel (proj_gemm_bf16, lstm_bf16.cu) keeps a W block resident and streams 128-row A tiles past it.  At the
// hidden size of the reference's checkpoint (H = 256: N = 2048 gate columns, K = 512) only a 128-column W block fits next to
// the A ring, so every A tile crosses L2 -> SM sixteen times and the kernel sits on that path (840 TFLOP/s).  Here two CTAs
// form one tile 256 rows x 256 columns with `tcgen05.mma.cta_group::2`: each CTA still holds 128 W rows (K <= 512: 128 KB) and
// streams its OWN 128 A rows, but the pair's MMA multiplies them with BOTH CTAs' W rows -- an A tile now meets 256 columns per
// trip through the SM, i.e. half the L2 -> SM bytes per flop.
//
//   cluster (2,1,1), 320 threads per CTA:
//     warp 0   TMA producer: the CTA's half of the W block once (resident), then its A tiles through a 4/8-slot ring; all
//              transaction bytes are counted on the LEADER's mbarriers (cp.async.bulk.tensor ... cta_group::2)
//     warp 1   leader: MMA issuer (M 256 x N 256 x K 16 per instruction, commits multicast to both CTAs);
//              peer:   relays "my epilogue has drained accumulator a" to the leader (relaxed remote arrive)
//     warps 2-9  epilogue of the CTA's own 128 rows (TMEM lanes) x 256 columns: tcgen05.ld, + bias, bf16, staged in shared
//              memory, one bulk / TMA store per 64-column pass; two accumulators (2 x 256 TMEM columns) so the MMAs of the
//              next tile run under it
//   C layouts as in proj_gemm_bf16: row-major through a tensor map, or the recurrence's blocked streaming layout
//   [row / 128][n / 8][row % 128][8 bf16] (a CTA's 128 rows are exactly one row block).
#include "lstm_handle.cuh"
#include "sm100_prims.cuh"
#include "tmap.cuh"
#include <cstdlib>

namespace bci {
using namespace sm100;

constexpr int GP_BK = 64, GP_MAX_STAGES = 8, GP_MAX_K = 512;
constexpr int GP_THREADS = 320;
constexpr uint32_t GP_TILE = 128 * GP_BK * 2;      // 16 KB: 128 rows x 64 bf16, one SW128 atom column
constexpr uint32_t GP_C_BYTES = 2 * GP_TILE;       // C staging: one [128][64] bf16 block per epilogue group
static inline int gp_stages(int K) { return K <= 256 ? 8 : 4; }
static inline size_t gp_smem(int K) {
  return 1024 + (size_t)(K / GP_BK) * GP_TILE + (size_t)gp_stages(K) * GP_TILE + GP_C_BYTES + 256 * sizeof(float) + 256;
}

template <bool BLOCKED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
proj_gemm_bf16_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, __nv_bfloat16* __restrict__ Cblk, const float* __restrict__ bias, int M,
                    int N, int K, int nstages) {
  extern __shared__ uint8_t gp_smem_raw[];
  const uint32_t raw = smem_u32(gp_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = gp_smem_raw + (base - raw);
  const int k_blocks = K / GP_BK;
  const uint32_t b_bytes = (uint32_t)k_blocks * GP_TILE;
  const uint32_t sB = base, sA = base + b_bytes, sC = sA + (uint32_t)nstages * GP_TILE;
  uint8_t* genC = gen + b_bytes + (uint32_t)nstages * GP_TILE;
  float* bias_s = reinterpret_cast<float*>(genC + GP_C_BYTES);   // [256]
  uint8_t* ctl = genC + GP_C_BYTES + 256 * sizeof(float);
  const uint32_t bar0 = smem_u32(ctl);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                                  // leader: both CTAs' A tiles of the slot have landed
  auto empty_bar = [&](int s) { return bar0 + 8u * (GP_MAX_STAGES + s); };               // every CTA: the pair's MMAs have read the slot
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * GP_MAX_STAGES + a); };           // every CTA: accumulator a is complete
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * GP_MAX_STAGES + 2 + a); };      // every CTA: its epilogue has drained accumulator a
  auto peer_tempty_bar = [&](int a) { return bar0 + 8u * (2 * GP_MAX_STAGES + 4 + a); }; // leader: relay of the peer's tempty
  const uint32_t bfull_bar = bar0 + 8u * (2 * GP_MAX_STAGES + 6);                        // leader: both halves of the W block are resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctl + 8 * (2 * GP_MAX_STAGES + 7));

  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_blocks = N / 256, m_blocks = (M + 255) / 256, m_blocks128 = (M + 127) / 128;
  const int n_clusters = (int)cluster_nclusters_x(), cid = (int)cluster_id_x();
  const int n_blk = cid % n_blocks, n0 = n_blk * 256;
  const int m_first = cid / n_blocks, m_step = n_clusters / n_blocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < GP_MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 256); mbar_init(peer_tempty_bar(a), 1); }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  if (warp >= 2) {
    const int et = (warp - 2) * 32 + lane;   // 0..255: the tile's columns (both CTAs add the whole tile's bias to their own rows)
    bias_s[et] = __ldg(bias + n0 + et);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();   // both CTAs' barriers exist before anything is signalled across the pair

  if (warp == 0) {
    // ---- TMA producer
    if (tp_elect_one()) {
      if (leader) mbar_arrive_expect_tx(bfull_bar, 2 * b_bytes);
      for (int kb = 0; kb < k_blocks; ++kb) tma_load_2d_2sm(sB + kb * GP_TILE, &tmB, kb * GP_BK, n0 + (int)rank * 128, bfull_bar);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int mb = m_first; mb < m_blocks; mb += m_step) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (tp_elect_one()) {
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * GP_TILE);
          tma_load_2d_2sm(sA + stage * GP_TILE, &tmA, kb * GP_BK, mb * 256 + (int)rank * 128, full_bar(stage));
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---- MMA issuer
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      mbar_wait(bfull_bar, 0);
      for (int mb = m_first; mb < m_blocks; mb += m_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        mbar_wait_cluster(peer_tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (tp_elect_one()) {
#pragma unroll
            for (int kk = 0; kk < GP_BK / 16; ++kk) {
              const uint64_t da = umma_desc_sw128(sA + stage * GP_TILE + kk * 32);
              const uint64_t db = umma_desc_sw128(sB + kb * GP_TILE + kk * 32);
              umma_bf16_2sm(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit_2sm_mc(empty_bar(stage), (uint16_t)3);   // frees the slot in both CTAs when these MMAs retire
            if (kb == k_blocks - 1) umma_commit_2sm_mc(tfull_bar(acc), (uint16_t)3);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    } else {
      // ---- peer: tells the leader when this CTA's epilogue has drained an accumulator (relaxed: it publishes no data)
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t remote0 = mapa_u32(peer_tempty_bar(0), 0);
      for (int mb = m_first; mb < m_blocks; mb += m_step) {
        mbar_wait(tempty_bar(acc), acc_phase);
        if (tp_elect_one()) mbar_arrive_cluster_relaxed(remote0 + 8u * (uint32_t)acc);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ---- epilogue: 8 warps = 2 groups of 4; warp w owns TMEM lane quarter (w % 4), group g columns [128 g, 128 g + 128) of the
    // tile and its own 16 KB staging block; each 64-column pass is converted into the staging block and written with one store
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int rt = quarter * 32 + lane;   // row inside the CTA's 128-row block
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    uint8_t* cst = genC + grp * GP_TILE;
    const uint32_t cst_s = sC + grp * GP_TILE;
    int acc = 0; uint32_t acc_phase = 0;
    for (int mb = m_first; mb < m_blocks; mb += m_step) {
      const int mblk = 2 * mb + (int)rank;   // this CTA's 128-row block
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * 256u + grp * 128;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if ((ch & 1) == 0) {
          if (issuer) tma_store_wait_read();   // the group's staging block has been drained by its previous store
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
        tmem_ld_wait();
        if (ch + 1 < 4) tmem_ld32(taddr + (ch + 1) * 32, r[(ch + 1) & 1]);
        const uint32_t* rc = r[ch & 1];
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v0 = __uint_as_float(rc[2 * j]) + bias_s[grp * 128 + ch * 32 + 2 * j];
          const float v1 = __uint_as_float(rc[2 * j + 1]) + bias_s[grp * 128 + ch * 32 + 2 * j + 1];
          __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
          o[j] = *reinterpret_cast<uint32_t*>(&p);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t cidx = (uint32_t)((ch & 1) * 4 + q);   // 16-byte chunk inside the 64-column pass
          uint8_t* dstp = BLOCKED ? cst + cidx * 2048u + (uint32_t)rt * 16u : cst + sw128_chunk_off((uint32_t)rt, cidx);
          *reinterpret_cast<uint4*>(dstp) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        if ((ch & 1) == 1) {
          if (ch == 3) {   // all TMEM reads of this accumulator (by this thread) are done
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (issuer) {
            const int col = n0 + grp * 128 + (ch >> 1) * 64;
            if (mblk < m_blocks128) {   // (an odd number of row blocks: the peer's last block does not exist)
              if (BLOCKED) bulk_store_s2g(Cblk + ((size_t)mblk * N + (size_t)col) * 128, cst_s, GP_TILE);
              else tma_store_2d(&tmC, cst_s, col, mblk * 128);
            }
            tma_store_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();   // no CTA leaves while the pair's MMAs may still read its shared memory or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// BCI_GEMM_PAIR=off keeps the one-CTA kernels
static bool gp_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BCI_GEMM_PAIR");
    v = (e && e[0] == 'o') ? 0 : 1;
  }
  return v != 0;
}

static int gp_max_clusters() {
  static PerDeviceInt state_pd, max_pd;   // state: 0 = not tried, 1 = ok, -1 = unavailable
  int& state = state_pd.cur();
  int& mx = max_pd.cur();
  if (state == 0) {
    state = -1;
    const int smem = (int)gp_smem(GP_MAX_K);
    if (cudaFuncSetAttribute(proj_gemm_bf16_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(proj_gemm_bf16_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count(), 1, 1);
    cfg.blockDim = dim3(GP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, proj_gemm_bf16_pair<true>, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return 0;
    }
    mx = n;
    state = 1;
  }
  return state == 1 ? mx : 0;
}

// returns 1 when the shape is not for this kernel (the caller then runs the one-CTA kernel), BCI_OK / an error otherwise
int launch_proj_gemm_bf16_pair(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, __nv_bfloat16* C, int M, int N,
                               int K, bool blocked, cudaStream_t st) {
  if (!(gp_enabled() && M >= 512 && N % 256 == 0 && K % GP_BK == 0 && K >= GP_BK && K <= GP_MAX_K)) return 1;
  const int n_blocks = N / 256;
  const int mx = gp_max_clusters();
  if (mx < n_blocks) return 1;
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, 64, 128);
  if (rc) return rc;
  if ((rc = make_tmap_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tmC, C, (uint64_t)M, (uint64_t)N, 64, 128))) return rc;   // (unused for the blocked layout)
  const int m_blocks = ceil_div(M, 256);
  int per_n = mx / n_blocks;
  if (per_n > m_blocks) per_n = m_blocks;
  const int clusters = per_n * n_blocks;
  if (blocked) proj_gemm_bf16_pair<true><<<2 * clusters, GP_THREADS, gp_smem(K), st>>>(tmA, tmB, tmC, C, bias, M, N, K, gp_stages(K));
  else proj_gemm_bf16_pair<false><<<2 * clusters, GP_THREADS, gp_smem(K), st>>>(tmA, tmB, tmC, C, bias, M, N, K, gp_stages(K));
  BCI_LAUNCH_OK();
  return BCI_OK;
}

}  // namespace bci
