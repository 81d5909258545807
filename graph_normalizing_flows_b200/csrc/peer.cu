// The one collective of the hot path over NVLink peer memory: all-reduce(SUM) of the fp64 4-vector
// (log_prob_zs, log_det_jacobian, log_prob_xs, num_nodes) across the ranks of one node.
//
// Why not NCCL for 32 bytes: an NCCL kernel cannot share an SM with the fused coupling kernel (one 203 KB CTA per SM),
// so it waits for a CTA to exit and then holds that SM while it spins for the slowest peer -- the next fused launch
// finishes late by that much (measured: +0.1 ms per 3.9 ms step at 2 GPUs).  The kernels here are one warp (or the
// tail of the log-prob kernel itself), a few registers, no shared memory to speak of: they co-reside with the fused
// CTAs, push 40 bytes to every peer with plain stores through the NVLink mapping (cudaIpc) and spin on their own
// memory.
//
// Protocol: every rank owns slots[2][world][8] 8-byte words.  Call number `seq` (1, 2, ...; the same on all ranks,
// they call in the same order) uses parity seq & 1.  Rank r writes its 4 doubles into slot [parity][r] of EVERY rank,
// then -- after a system-scope fence -- the flag word seq (st.release.sys).  Each rank waits until all `world` flags of
// its own slots read seq (ld.acquire.sys), then adds the vectors in RANK order: every rank computes bit-identical sums.
// A rank can be at most one call ahead of a peer (it needs the peer's flag of call seq to leave call seq), so two
// parities are enough.  The spin gives up after ~2 s (NaN result): a dead peer must not hang the device.
#include "common.cuh"

struct gnf_peer {
  int rank = 0, world = 1;
  unsigned long long* local = nullptr;          // slots[2][world][8]
  unsigned long long* remote[8] = {};           // remote[r] = rank r's slots, mapped into this process
  unsigned long long seq = 0;
  bool opened[8] = {};
};

namespace gnf {
namespace {

constexpr int kSlotWords = 8;

struct PeerArgs {
  unsigned long long* remote[8];
  unsigned long long* local;
  int rank, world;
  unsigned long long seq;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// lanes 0..world-1 of ONE warp: push `mine` to every peer, wait for every peer, sum in rank order into out[4]
__device__ __forceinline__ void peer_allreduce4(const PeerArgs& a, const double (&mine)[4], double* out, int lane) {
  const int parity = (int)(a.seq & 1ull);
  if (lane < a.world) {
    unsigned long long* dst = a.remote[lane] + ((size_t)parity * a.world + a.rank) * kSlotWords;
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<volatile double*>(dst)[k] = mine[k];
    __threadfence_system();
    st_release_sys(dst + 4, a.seq);
  }
  double got[4] = {0.0, 0.0, 0.0, 0.0};
  bool ok = true;
  if (lane < a.world) {
    const unsigned long long* src = a.local + ((size_t)parity * a.world + lane) * kSlotWords;
    const unsigned long long t0 = timer_ns();
    while (ld_acquire_sys(src + 4) != a.seq) {
      __nanosleep(64);
      if (timer_ns() - t0 > 2000000000ull) { ok = false; break; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) got[k] = reinterpret_cast<const volatile double*>(src)[k];
  }
  ok = __all_sync(0xffffffffu, ok);
  // fixed order: rank 0 + rank 1 + ... (lane r holds rank r's vector)
  double sum[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < a.world; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) sum[k] += __shfl_sync(0xffffffffu, got[k], r);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = ok ? sum[k] : __longlong_as_double(0x7ff8000000000000ll);
}

// stand-alone: vec[4] (device, this rank's values) -> all-reduced in place.  One warp.
__global__ void __launch_bounds__(32) k_peer_allreduce4(double* __restrict__ vec, const PeerArgs a) {
  const int lane = threadIdx.x;
  double mine[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) mine[k] = vec[k];
  __syncwarp();
  peer_allreduce4(a, mine, vec, lane);
}

// a8 + the collective in ONE kernel: the fixed-order fold of the |z|^2 partials, the log-prob assembly
// (run_grevnet.py:292-296) and the all-reduce of the resulting 4-vector.
__global__ void __launch_bounds__(256)
k_log_prob_allreduce(const double* __restrict__ partials, int n, const double* __restrict__ ldj, double n_nodes, int d,
                     double* __restrict__ out, const PeerArgs a) {
  __shared__ double red[256];
  double local = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) local += partials[i];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 32) {
    const double kLog2Pi = 1.8378770664093454835606594728112;
    const double lpz = -0.5 * red[0] - 0.5 * (double)d * kLog2Pi * n_nodes;
    const double l = ldj ? ldj[0] : 0.0;
    const double mine[4] = {lpz, l, lpz + l, n_nodes};
    peer_allreduce4(a, mine, out, threadIdx.x);
  }
}

PeerArgs make_args(gnf_peer* p) {
  PeerArgs a;
  for (int r = 0; r < 8; ++r) a.remote[r] = p->remote[r];
  a.local = p->local;
  a.rank = p->rank;
  a.world = p->world;
  a.seq = ++p->seq;
  return a;
}

}  // namespace
}  // namespace gnf

using namespace gnf;

extern "C" int gnf_peer_create(gnf_peer** out, int32_t rank, int32_t world, uint8_t* handle_out) {
  GNF_REQUIRE(out && handle_out, GNF_EINVAL, "gnf_peer_create: null pointer");
  GNF_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, GNF_EINVAL,
              "gnf_peer_create: need 1 <= world <= 8 (one node) and 0 <= rank < world");
  static_assert(sizeof(cudaIpcMemHandle_t) == GNF_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  gnf_peer* p = new gnf_peer();
  p->rank = rank;
  p->world = world;
  const size_t bytes = (size_t)2 * world * kSlotWords * sizeof(unsigned long long);
  cudaError_t e = cudaMalloc(&p->local, bytes);
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
  if (e != cudaSuccess) {
    set_error("gnf_peer_create: %s", cudaGetErrorString(e));
    cudaFree(p->local);
    delete p;
    return GNF_ECUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  p->remote[rank] = p->local;
  *out = p;
  return GNF_OK;
}

extern "C" int gnf_peer_connect(gnf_peer* p, const uint8_t* handles_host) {
  GNF_REQUIRE(p && handles_host, GNF_EINVAL, "gnf_peer_connect: null pointer");
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles_host + (size_t)r * GNF_PEER_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("gnf_peer_connect: cudaIpcOpenMemHandle(rank %d) failed: %s (no peer access between these GPUs?)", r,
                cudaGetErrorString(e));
      return GNF_ECUDA;
    }
    p->remote[r] = (unsigned long long*)ptr;
    p->opened[r] = true;
  }
  return GNF_OK;
}

extern "C" int gnf_peer_destroy(gnf_peer* p) {
  if (!p) return GNF_OK;
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->remote[r]);
  cudaFree(p->local);
  delete p;
  return GNF_OK;
}

extern "C" int gnf_peer_allreduce4(gnf_peer* p, double* vec, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(p && vec, GNF_EINVAL, "gnf_peer_allreduce4: null pointer");
  for (int r = 0; r < p->world; ++r) GNF_REQUIRE(p->remote[r], GNF_EINVAL, "gnf_peer_allreduce4: rank %d not connected", r);
  k_peer_allreduce4<<<1, 32, 0, stream>>>(vec, make_args(p));
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// defined in flow.cu: launches k_sumsq into the workspace and returns the number of partials
namespace gnf { int log_prob_partials(const float* z, int64_t n, int32_t d, void* ws, size_t ws_bytes, cudaStream_t stream, int* blocks); }

extern "C" int gnf_log_prob_allreduce(const float* z, int64_t n, int32_t d, const double* ldj, double* out, void* ws,
                                      size_t ws_bytes, gnf_peer* p, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(p && out, GNF_EINVAL, "gnf_log_prob_allreduce: null pointer");
  for (int r = 0; r < p->world; ++r)
    GNF_REQUIRE(p->remote[r], GNF_EINVAL, "gnf_log_prob_allreduce: rank %d not connected", r);
  int blocks = 0;
  int rc = log_prob_partials(z, n, d, ws, ws_bytes, stream, &blocks);
  if (rc) return rc;
  k_log_prob_allreduce<<<1, 256, 0, stream>>>((const double*)ws, blocks, ldj, (double)n, d, out, make_args(p));
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}
