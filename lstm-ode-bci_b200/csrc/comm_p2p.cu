// Data-parallel optimizer step fused with its collective (BASELINE config 3; SURVEY.md §8 b `bci_fused_step`, §8 e).
//
// One process per GPU.  Every rank owns a gradient bucket allocated HERE with cudaMalloc and exported as a CUDA IPC
// handle, so after bci_comm_connect each rank holds peer pointers to all buckets over NVLink / NVSwitch.  A step is
//
//   p2p_reduce_kernel : cross-rank "ready" flags (system-scope release/acquire on a signal pad in the same
//                       allocation) -> every rank READS the same slice layout of all peer buckets (16-byte loads
//                       through NVLink), sums them in rank order (bit-identical on all ranks), writes the sum to a
//                       local buffer and a per-block partial of the squared norm -> "done" flags
//   adamw_p2p_kernel  : total norm from the partials in fixed order (bit-identical on all ranks, unlike atomics),
//                       clip_grad_norm_ + AdamW on the local replica, then waits for the peers' "done" flags so the
//                       next backward may overwrite the bucket
//
// i.e. the reduction is the load phase of the optimizer kernel pair -- no separate all-reduce pass over a
// staging buffer, no NCCL call on the step's critical path.  Replaces 04_lstm_model.py:490-507
// (loss.backward(); clip_grad_norm_; optimizer.step()) for the data-parallel case the reference does not have.
#include "common.cuh"
#include <cstring>
#include <new>

namespace bci {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_PAD_WORDS = 64;      // [0,16): ready flags by source rank; [16,32): done flags
constexpr int COMM_BLOCKS_MAX = 296;    // partials array size (2 x 148)

struct HandleBlob {           // BCI_COMM_HANDLE_BYTES
  cudaIpcMemHandle_t mem;     // 64 bytes
  int32_t rank, world;
  int64_t n_floats;
  uint64_t magic;
  char pad[BCI_COMM_HANDLE_BYTES - 64 - 4 - 4 - 8 - 8];
};
static_assert(sizeof(HandleBlob) == BCI_COMM_HANDLE_BYTES, "handle blob size");

struct PeerTable {
  const float* bucket[COMM_MAX_WORLD];
  uint32_t* pad[COMM_MAX_WORLD];
};

}  // namespace bci

struct bci_comm_s {
  int rank, world, device;
  int64_t n;            // floats in the bucket
  void* base;           // local allocation: [bucket n floats, padded][signal pad][reduced n floats][partials]
  float* bucket;
  uint32_t* pad;
  float* reduced;
  float* partials;
  void* peer_base[bci::COMM_MAX_WORLD];
  bci::PeerTable tab;
  bool connected;
  uint32_t epoch;
};

namespace bci {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data must never be served from a stale L1 line of the previous step (same addresses every step)
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// grid-stride over float4 groups; block b of every rank covers the same elements, partial[b] is its sum of squares
__global__ void __launch_bounds__(256)
p2p_reduce_kernel(PeerTable tab, int rank, int world, long long n, uint32_t epoch, float* __restrict__ reduced,
                  float* __restrict__ partials) {
  // ---- my bucket is complete (stream order): tell every peer, then wait until every peer said the same ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(tab.pad[threadIdx.x] + rank, epoch);  // ready[rank] in peer threadIdx.x's pad
  }
  if (threadIdx.x < world) {
    const uint32_t* f = tab.pad[rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
  float ss = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = ld_peer_v4(tab.bucket[r] + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(reduced)[i] = s;
    ss = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, fmaf(s.w, s.w, ss))));
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += ld_peer_f32(tab.bucket[r] + i);
      reduced[i] = s;
      ss = fmaf(s, s, ss);
    }
  }
  // deterministic block reduction (fixed tree)
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// signals "done reading" once ALL blocks of this rank's reduce kernel have finished (this kernel follows it in stream order)
__global__ void __launch_bounds__(256)
adamw_p2p_kernel(PeerTable tab, int rank, int world, uint32_t epoch, float* __restrict__ p, const float* __restrict__ g,
                 float* __restrict__ m, float* __restrict__ v, long long n, const float* __restrict__ partials, int n_partials,
                 float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float grad_scale, float max_norm,
                 float* __restrict__ norm_out) {
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(tab.pad[threadIdx.x] + 16 + rank, epoch);  // done[rank]
  // total of the partials in a fixed order (lane-strided sums, then the shuffle tree): same bits on every rank and block
  __shared__ float tot_s;
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += 32) t += partials[i];
    t = warp_sum(t);
    if (threadIdx.x == 0) tot_s = t;
  }
  __syncthreads();
  const float tot = tot_s;
  const float total = sqrtf(tot) * fabsf(grad_scale);
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(1.0f, max_norm / (total + 1e-6f));
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) { norm_out[0] = tot; norm_out[1] = total; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    pi -= (lr / bc1) * (mi / (sqrtf(vi) / sqrtf(bc2) + eps));
    p[i] = pi;
  }
  // the next backward overwrites my bucket: every peer must have finished reading it
  if (blockIdx.x == 0 && threadIdx.x < world) {
    const uint32_t* f = tab.pad[rank] + 16 + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
  }
}

}  // namespace bci

using namespace bci;

extern "C" int bci_comm_create(int32_t rank, int32_t world, int64_t n_floats, bci_comm_t* out) {
  BCI_REQUIRE(out && world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world && n_floats > 0, BCI_EINVAL,
              "bci_comm_create: bad arguments (rank %d, world %d, n %lld)", rank, world, (long long)n_floats);
  bci_comm_s* c = new (std::nothrow) bci_comm_s();
  BCI_REQUIRE(c, BCI_ENOMEM, "bci_comm_create: host allocation failed");
  std::memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->n = n_floats;
  cudaError_t e = cudaGetDevice(&c->device);
  const size_t nb = align_up((size_t)n_floats * 4, 256);
  const size_t total = nb + 256 + nb + COMM_BLOCKS_MAX * 4 + 256;
  if (e == cudaSuccess) e = cudaMalloc(&c->base, total);
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, total);
  if (e != cudaSuccess) { set_error("bci_comm_create: %s", cudaGetErrorString(e)); delete c; return BCI_ENOMEM; }
  c->bucket = reinterpret_cast<float*>(c->base);
  c->pad = reinterpret_cast<uint32_t*>((char*)c->base + nb);
  c->reduced = reinterpret_cast<float*>((char*)c->base + nb + 256);
  c->partials = reinterpret_cast<float*>((char*)c->base + nb + 256 + nb);
  c->tab.bucket[rank] = c->bucket;
  c->tab.pad[rank] = c->pad;
  c->connected = (world == 1);
  *out = c;
  return BCI_OK;
}

extern "C" int bci_comm_export(bci_comm_t c, void* handle_host) {
  BCI_REQUIRE(c && handle_host, BCI_EINVAL, "bci_comm_export: NULL argument");
  HandleBlob b;
  std::memset(&b, 0, sizeof(b));
  BCI_CUDA_OK(cudaIpcGetMemHandle(&b.mem, c->base));
  b.rank = c->rank; b.world = c->world; b.n_floats = c->n; b.magic = 0xB200C0331ull;
  std::memcpy(handle_host, &b, sizeof(b));
  return BCI_OK;
}

extern "C" int bci_comm_connect(bci_comm_t c, const void* handles_host) {
  BCI_REQUIRE(c && handles_host, BCI_EINVAL, "bci_comm_connect: NULL argument");
  const size_t nb = align_up((size_t)c->n * 4, 256);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    HandleBlob b;
    std::memcpy(&b, (const char*)handles_host + (size_t)r * BCI_COMM_HANDLE_BYTES, sizeof(b));
    BCI_REQUIRE(b.magic == 0xB200C0331ull && b.rank == r && b.world == c->world && b.n_floats == c->n, BCI_EINVAL,
                "bci_comm_connect: handle %d does not describe rank %d of a %d-rank, %lld-float communicator", r, r, c->world,
                (long long)c->n);
    BCI_CUDA_OK(cudaIpcOpenMemHandle(&c->peer_base[r], b.mem, cudaIpcMemLazyEnablePeerAccess));
    c->tab.bucket[r] = reinterpret_cast<const float*>(c->peer_base[r]);
    c->tab.pad[r] = reinterpret_cast<uint32_t*>((char*)c->peer_base[r] + nb);
  }
  c->connected = true;
  return BCI_OK;
}

extern "C" int bci_comm_bucket(bci_comm_t c, float** bucket, int64_t* n_floats) {
  BCI_REQUIRE(c && bucket, BCI_EINVAL, "bci_comm_bucket: NULL argument");
  *bucket = c->bucket;
  if (n_floats) *n_floats = c->n;
  return BCI_OK;
}

extern "C" int bci_comm_destroy(bci_comm_t c) {
  if (!c) return BCI_OK;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->base) cudaFree(c->base);
  delete c;
  return BCI_OK;
}

extern "C" int bci_fused_step(bci_comm_t c, float* p, float* m, float* v, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int32_t step, float max_norm, float* norm_out, void* stream) {
  bci::NvtxRange nvtx_range("bci_fused_step");
  BCI_REQUIRE(c && p && m && v && step >= 1, BCI_EINVAL, "bci_fused_step: bad arguments");
  BCI_REQUIRE(c->connected, BCI_ESTATE, "bci_fused_step: call bci_comm_connect first");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = c->n;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  if (blocks > COMM_BLOCKS_MAX) blocks = COMM_BLOCKS_MAX;
  if (blocks < 1) blocks = 1;
  const uint32_t epoch = ++c->epoch;
  p2p_reduce_kernel<<<blocks, 256, 0, st>>>(c->tab, c->rank, c->world, n, epoch, c->reduced, c->partials);
  BCI_LAUNCH_OK();
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_p2p_kernel<<<blocks, 256, 0, st>>>(c->tab, c->rank, c->world, epoch, p, c->reduced, m, v, n, c->partials, blocks, lr, beta1,
                                           beta2, eps, weight_decay, bc1, bc2, 1.0f / (float)c->world, max_norm, norm_out);
  BCI_LAUNCH_OK();
  return BCI_OK;
}
