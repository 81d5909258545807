// Batch-structure and segment kernels.
//
// Bit-exactness contract: the reference's aggregator is tf.unsorted_segment_sum, whose CPU
// kernel adds edge rows into their receiver segment serially in ascending edge index
// (gnn.py:103-104 via graph_nets ReceivedEdgesToNodesAggregator).  We therefore build a
// STABLE CSR-by-receiver once per batch and accumulate each (receiver, feature) serially in
// that order.  Parallelism is across (node, feature), never across one segment's edges.
#include "common.cuh"

namespace gnf {
namespace {

constexpr int kScanItems = 4;
constexpr int kScanThreads = 1024;
constexpr int kScanTile = kScanItems * kScanThreads;

__global__ void k_validate(const int32_t* __restrict__ senders, const int32_t* __restrict__ receivers,
                           int64_t n_nodes, int64_t n_edges, int32_t* __restrict__ bad) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int local = 0;
  for (; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    int32_t s = senders[e], r = receivers[e];
    local += (s < 0 || s >= n_nodes) + (r < 0 || r >= n_nodes);
  }
  if (local) atomicAdd(bad, local);
}

// ids outside [0, n_nodes) are skipped here and in k_fill (they are counted by k_validate and reported as a
// ValueError by the host; with deferred validation the CSR build runs before that count is read, so it must
// stay inside its buffers whatever the ids are)
__global__ void k_degree(const int32_t* __restrict__ receivers, int64_t n_edges, int64_t n_nodes,
                         int32_t* __restrict__ deg) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n_edges) {
    const int32_t r = receivers[e];
    if (r >= 0 && r < n_nodes) atomicAdd(&deg[r], 1);
  }
}

// exclusive scan, pass 1: per-tile exclusive scan + tile totals
__global__ void __launch_bounds__(kScanThreads)
k_scan_tiles(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out,
             int32_t* __restrict__ tile_sums) {
  __shared__ int32_t warp_sums[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    sum += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_sums[lane];
    int32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_sums[lane] = wi - w;  // exclusive
    if (lane == 31) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  int32_t excl = warp_sums[warp] + incl - sum;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
}

// pass 2: one block scans the tile totals in place (exclusive), serial carry over chunks
__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int n) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    int32_t v = (i < n) ? sums[i] : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_sums[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    int32_t carry = carry_s;
    int32_t excl = carry + warp_sums[warp] + incl - v;
    if (i < n) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}

// pass 3: add tile offsets over all n + 1 entries: rowptr[n] = number of edges with a valid receiver
// (= n_edges for validated input; never the raw n_edges, so the last segment cannot cover unfilled slots)
__global__ void k_scan_add(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ tile_sums) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i <= n) out[i] += tile_sums[i / kScanTile];
}

// unordered placement into segments
__global__ void k_fill(const int32_t* __restrict__ receivers, int64_t n_edges, int64_t n_nodes,
                       const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                       int32_t* __restrict__ tmp) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n_edges) {
    int32_t r = receivers[e];
    if (r < 0 || r >= n_nodes) return;
    int32_t pos = atomicAdd(&cursor[r], 1);
    tmp[rowptr[r] + pos] = (int32_t)e;
  }
}

// one warp per receiver: rank the segment's edge ids ascending (-> stable CSR), emit
// perm and csr_senders.  deg <= 32: in registers; otherwise O(deg^2/32) counting.
__global__ void k_sort_segments(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tmp,
                                const int32_t* __restrict__ senders, int64_t n_nodes,
                                int32_t* __restrict__ perm, int32_t* __restrict__ csr_senders) {
  const int lane = threadIdx.x & 31;
  int64_t node = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (node >= n_nodes) return;
  const int32_t beg = rowptr[node], end = rowptr[node + 1];
  const int deg = end - beg;
  if (deg <= 32) {
    int32_t mine = (lane < deg) ? tmp[beg + lane] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      int32_t other = __shfl_sync(0xffffffffu, mine, j);
      rank += (other < mine);
    }
    if (lane < deg) {
      perm[beg + rank] = mine;
      csr_senders[beg + rank] = senders[mine];
    }
  } else {
    for (int i = lane; i < deg; i += 32) {
      int32_t mine = tmp[beg + i];
      int rank = 0;
      for (int j = 0; j < deg; ++j) rank += (tmp[beg + j] < mine);
      perm[beg + rank] = mine;
      csr_senders[beg + rank] = senders[mine];
    }
  }
}

__global__ void k_gather_rows(const float* __restrict__ x, int h, const int32_t* __restrict__ senders,
                              int64_t total, float* __restrict__ edges) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) {
    int64_t e = i / h;
    int f = (int)(i - e * h);
    edges[i] = x[(int64_t)senders[e] * h + f];
  }
}

// a3 + a4 standalone: out[r, :] = reduce over the CSR segment of r of src[idx[e], :], SERIAL in ascending e per
// (receiver, feature) -- the bits of TF-CPU UnsortedSegmentSum.  HBM-bound sub-op.
//
// A CTA owns kSegNodes consecutive receivers = ONE contiguous CSR range.  What it needs is staged in shared memory
// with coalesced loads: rowptr, its index range (each index leaves DRAM once) and -- when h <= 8 and the senders of
// that range span at most kSegRows rows (graphs are contiguous node blocks, so they do) -- those rows, re-laid as
// padded rows (8 floats used of a 12-float stride).  Then ONE thread per receiver walks its segment: per edge one index read and two 128-bit
// shared-memory loads feed h serial adds, 7x fewer instructions and L1 wavefronts than a thread per (receiver,
// feature).  Results leave through a staging row: the CTA's h * 128 output floats are one contiguous, coalesced store.
// (ncu of the thread-per-feature kernel this replaces: issue slots 79 % busy, L1 data pipe 50 %, DRAM 13 % --
// instruction bound, not bandwidth bound.)
// CTAs that cannot stage (h > 8, materialised edge tensors whose rows are spread out, very high in-degree) fall back
// to a thread per (receiver, feature) with 4 loads in flight -- same order, same bits.
constexpr int kSegNodes = 128;
constexpr int kSegCap = 3072;      // staged indices per CTA
constexpr int kSegRows = 384;      // staged sender rows per CTA
constexpr int kSegRowStride = 12;  // floats between staged rows: 48 B keeps 16-byte alignment and spreads the rows over 8
                                   // bank groups (32 B rows only reach 4: ncu showed 40 % of the wavefronts were conflicts)

__global__ void __launch_bounds__(kSegNodes)
k_gather_segment(const float* __restrict__ src, int h, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ idx, int64_t n_nodes, int mean, float* __restrict__ out) {
  __shared__ int32_t s_row[kSegNodes + 1];
  __shared__ int32_t s_idx[kSegCap];
  __shared__ __align__(16) float s_rows[kSegRows * kSegRowStride];
  __shared__ float s_out[kSegNodes * 8];
  __shared__ int32_t s_red[2 * (kSegNodes / 32)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * kSegNodes;
  const int nn = (int)((n_nodes - n0) < kSegNodes ? (n_nodes - n0) : kSegNodes);
  for (int i = tid; i <= nn; i += kSegNodes) s_row[i] = rowptr[n0 + i];
  __syncthreads();
  const int32_t e0 = s_row[0];
  const int32_t ne = s_row[nn] - e0;
  const bool idx_staged = ne <= kSegCap;
  int32_t lo = 0x7fffffff, hi = -1;
  if (idx_staged)
    for (int i = tid; i < ne; i += kSegNodes) {
      const int32_t v = idx[e0 + i];
      s_idx[i] = v;
      lo = min(lo, v);
      hi = max(hi, v);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { s_red[warp] = lo; s_red[kSegNodes / 32 + warp] = hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSegNodes / 32; ++w) { lo = min(lo, s_red[w]); hi = max(hi, s_red[kSegNodes / 32 + w]); }
  const int nrows = ne > 0 ? hi - lo + 1 : 0;

  if (idx_staged && h <= 8 && nrows <= kSegRows) {
    // ---- staged rows: lane r reads h words at stride h, so a warp covers 32 consecutive rows of src ----------
    for (int r = tid; r < nrows; r += kSegNodes) {
      const float* __restrict__ g = src + (int64_t)(lo + r) * h;
      float v[8];
#pragma unroll
      for (int f = 0; f < 8; ++f) v[f] = f < h ? g[f] : 0.f;
      *reinterpret_cast<float4*>(&s_rows[r * kSegRowStride]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s_rows[r * kSegRowStride + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    if (tid < nn) {
      float acc[8];
#pragma unroll
      for (int f = 0; f < 8; ++f) acc[f] = 0.f;
      const int32_t beg = s_row[tid] - e0, end = s_row[tid + 1] - e0;
      int32_t e = beg;
      for (; e + 2 <= end; e += 2) {                      // two rows in flight, added in edge order
        const int r0 = s_idx[e] - lo, r1 = s_idx[e + 1] - lo;
        const float4 a0 = *reinterpret_cast<const float4*>(&s_rows[r0 * kSegRowStride]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_rows[r0 * kSegRowStride + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&s_rows[r1 * kSegRowStride]);
        const float4 b1 = *reinterpret_cast<const float4*>(&s_rows[r1 * kSegRowStride + 4]);
        acc[0] = __fadd_rn(__fadd_rn(acc[0], a0.x), b0.x); acc[1] = __fadd_rn(__fadd_rn(acc[1], a0.y), b0.y);
        acc[2] = __fadd_rn(__fadd_rn(acc[2], a0.z), b0.z); acc[3] = __fadd_rn(__fadd_rn(acc[3], a0.w), b0.w);
        acc[4] = __fadd_rn(__fadd_rn(acc[4], a1.x), b1.x); acc[5] = __fadd_rn(__fadd_rn(acc[5], a1.y), b1.y);
        acc[6] = __fadd_rn(__fadd_rn(acc[6], a1.z), b1.z); acc[7] = __fadd_rn(__fadd_rn(acc[7], a1.w), b1.w);
      }
      if (e < end) {
        const int r0 = s_idx[e] - lo;
        const float4 a0 = *reinterpret_cast<const float4*>(&s_rows[r0 * kSegRowStride]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_rows[r0 * kSegRowStride + 4]);
        acc[0] = __fadd_rn(acc[0], a0.x); acc[1] = __fadd_rn(acc[1], a0.y); acc[2] = __fadd_rn(acc[2], a0.z);
        acc[3] = __fadd_rn(acc[3], a0.w); acc[4] = __fadd_rn(acc[4], a1.x); acc[5] = __fadd_rn(acc[5], a1.y);
        acc[6] = __fadd_rn(acc[6], a1.z); acc[7] = __fadd_rn(acc[7], a1.w);
      }
      if (mean) {
        const float dv = fmaxf((float)(end - beg), 1.f);
#pragma unroll
        for (int f = 0; f < 8; ++f) acc[f] = __fdiv_rn(acc[f], dv);
      }
#pragma unroll
      for (int f = 0; f < 8; ++f)
        if (f < h) s_out[tid * h + f] = acc[f];
    }
    __syncthreads();
    float* __restrict__ o = out + n0 * h;
    for (int i = tid; i < nn * h; i += kSegNodes) o[i] = s_out[i];
    return;
  }

  // ---- generic: thread per (receiver, feature) ---------------------------------------------------------------
  const int items = nn * h;
  for (int item = tid; item < items; item += kSegNodes) {
    const int node = item / h;
    const int f = item - node * h;
    const int32_t beg = s_row[node] - e0, end = s_row[node + 1] - e0;
    const float* __restrict__ col = src + f;
    float acc = 0.f;
    int32_t e = beg;
    if (idx_staged) {
      for (; e + 4 <= end; e += 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = col[(int64_t)s_idx[e + j] * h];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc = __fadd_rn(acc, v[j]);
      }
      for (; e < end; ++e) acc = __fadd_rn(acc, col[(int64_t)s_idx[e] * h]);
    } else {
      const int32_t* __restrict__ gi = idx + e0;
      for (; e + 4 <= end; e += 4) {
        const int32_t i0 = gi[e], i1 = gi[e + 1], i2 = gi[e + 2], i3 = gi[e + 3];
        const float v0 = col[(int64_t)i0 * h], v1 = col[(int64_t)i1 * h], v2 = col[(int64_t)i2 * h],
                    v3 = col[(int64_t)i3 * h];
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v0), v1), v2), v3);
      }
      for (; e < end; ++e) acc = __fadd_rn(acc, col[(int64_t)gi[e] * h]);
    }
    if (mean) acc = __fdiv_rn(acc, fmaxf((float)(end - beg), 1.f));
    out[(n0 + node) * h + f] = acc;
  }
}

// f3  pred_adj(graph, scaled_hacky_sigmoid_l2) (loss.py:154-159,45-53,131-151,83-85), block-diagonal:
// one CTA per graph writes its [n, n] block  sigmoid(10 * (1 - D_ij / sqrt(dim))),  D_ij = r_i - 2 x_i.x_j + r_j,
// diagonal zeroed.  The reference builds the dense [N, N] matrix and multiplies by a block mask.
__global__ void __launch_bounds__(256)
k_pred_adj(const float* __restrict__ x, int d, const int64_t* __restrict__ node_off,
           const int64_t* __restrict__ adj_off, float temp, float shift, float* __restrict__ out) {
  const int g = blockIdx.x;
  const int64_t lo = node_off[g];
  const int n = (int)(node_off[g + 1] - lo);
  float* blk = out + adj_off[g];
  const float inv_sqrt_d = 1.f / sqrtf((float)d);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    const float* xi = x + (lo + i) * d;
    const float* xj = x + (lo + j) * d;
    float ri = 0.f, rj = 0.f, dot = 0.f;
    for (int k = 0; k < d; ++k) {
      const float a = xi[k], b = xj[k];
      ri = fmaf(a, a, ri);
      rj = fmaf(b, b, rj);
      dot = fmaf(a, b, dot);
    }
    float dist = (ri - 2.f * dot + rj) * inv_sqrt_d;
    float v = 1.f / (1.f + expf(-temp * (shift - dist)));
    blk[idx] = (i == j) ? 0.f : v;
  }
}

}  // namespace
}  // namespace gnf

using namespace gnf;

extern "C" int gnf_pred_adj(const float* nodes, int32_t d, const int64_t* node_off, const int64_t* adj_off,
                            int64_t n_graphs, float temp, float shift, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(d > 0 && n_graphs >= 0, GNF_EINVAL, "gnf_pred_adj: bad shape");
  if (n_graphs == 0) return GNF_OK;
  GNF_REQUIRE(nodes && node_off && adj_off && out, GNF_EINVAL, "gnf_pred_adj: null pointer");
  k_pred_adj<<<(unsigned)n_graphs, 256, 0, stream>>>(nodes, d, node_off, adj_off, temp, shift, out);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" size_t gnf_build_csr_workspace(int64_t n_nodes, int64_t n_edges) {
  size_t tiles = (size_t)ceil_div(n_nodes + 1, kScanTile);
  return align_up((size_t)(n_nodes + 1) * 4, 256)   // deg / cursor
         + align_up((size_t)n_edges * 4, 256)        // tmp
         + align_up(tiles * 4 + 4, 256);             // tile sums
}

extern "C" int gnf_build_csr(const int32_t* receivers, const int32_t* senders, int64_t n_nodes,
                             int64_t n_edges, int32_t* rowptr, int32_t* perm, int32_t* csr_senders,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(n_nodes >= 0 && n_edges >= 0, GNF_EINVAL, "gnf_build_csr: negative size");
  GNF_REQUIRE(n_nodes < (1ll << 31) - 1 && n_edges < (1ll << 31) - 1, GNF_EINVAL,
              "gnf_build_csr: int32 index space exceeded");
  GNF_REQUIRE(rowptr && (n_edges == 0 || (receivers && senders && perm && csr_senders)), GNF_EINVAL,
              "gnf_build_csr: null pointer");
  GNF_REQUIRE(n_edges == 0 || n_nodes > 0, GNF_EINVAL, "gnf_build_csr: edges without nodes");
  GNF_REQUIRE(workspace_bytes >= gnf_build_csr_workspace(n_nodes, n_edges) &&
                  (workspace || workspace_bytes == 0),
              GNF_EWORKSPACE, "gnf_build_csr: workspace too small");
  uint8_t* ws = (uint8_t*)workspace;
  int32_t* deg = (int32_t*)ws;
  ws += align_up((size_t)(n_nodes + 1) * 4, 256);
  int32_t* tmp = (int32_t*)ws;
  ws += align_up((size_t)n_edges * 4, 256);
  int32_t* tile_sums = (int32_t*)ws;

  const int64_t n1 = n_nodes + 1;
  GNF_CUDA(cudaMemsetAsync(deg, 0, (size_t)n1 * 4, stream));
  if (n_edges > 0) {
    k_degree<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(receivers, n_edges, n_nodes, deg);
    GNF_LAUNCH_CHECK();
  }
  const int tiles = (int)ceil_div(n1, kScanTile);
  k_scan_tiles<<<tiles, kScanThreads, 0, stream>>>(deg, n1, rowptr, tile_sums);
  GNF_LAUNCH_CHECK();
  k_scan_sums<<<1, 1024, 0, stream>>>(tile_sums, tiles);
  GNF_LAUNCH_CHECK();
  k_scan_add<<<(unsigned)ceil_div(n1, 256), 256, 0, stream>>>(rowptr, n_nodes, tile_sums);
  GNF_LAUNCH_CHECK();
  if (n_edges > 0) {
    GNF_CUDA(cudaMemsetAsync(deg, 0, (size_t)n1 * 4, stream));  // reuse as cursor
    k_fill<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(receivers, n_edges, n_nodes, rowptr, deg, tmp);
    GNF_LAUNCH_CHECK();
    k_sort_segments<<<(unsigned)ceil_div(n_nodes * 32, 256), 256, 0, stream>>>(
        rowptr, tmp, senders, n_nodes, perm, csr_senders);
    GNF_LAUNCH_CHECK();
  }
  return GNF_OK;
}

extern "C" int gnf_validate_indices(const int32_t* senders, const int32_t* receivers,
                                    int64_t n_nodes, int64_t n_edges, int32_t* bad_count,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(bad_count, GNF_EINVAL, "gnf_validate_indices: null bad_count");
  GNF_CUDA(cudaMemsetAsync(bad_count, 0, 4, stream));
  if (n_edges > 0) {
    GNF_REQUIRE(senders && receivers, GNF_EINVAL, "gnf_validate_indices: null pointer");
    int blocks = (int)(ceil_div(n_edges, 256) < 4096 ? ceil_div(n_edges, 256) : 4096);
    k_validate<<<blocks, 256, 0, stream>>>(senders, receivers, n_nodes, n_edges, bad_count);
    GNF_LAUNCH_CHECK();
  }
  return GNF_OK;
}

extern "C" int gnf_gather_rows(const float* x, int32_t h, const int32_t* senders, int64_t n_edges,
                               float* edges, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(h > 0 && n_edges >= 0, GNF_EINVAL, "gnf_gather_rows: bad shape");
  if (n_edges == 0) return GNF_OK;
  GNF_REQUIRE(x && senders && edges, GNF_EINVAL, "gnf_gather_rows: null pointer");
  int64_t total = n_edges * h;
  k_gather_rows<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x, h, senders, total, edges);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

static int segment_common(const float* src, int32_t h, const int32_t* rowptr, const int32_t* idx,
                          int64_t n_nodes, int32_t agg, float* out, cudaStream_t stream,
                          const char* who) {
  GNF_REQUIRE(h > 0 && n_nodes >= 0, GNF_EINVAL, "%s: bad shape", who);
  GNF_REQUIRE(agg == GNF_AGG_SUM || agg == GNF_AGG_MEAN, GNF_EINVAL, "%s: bad agg %d", who, agg);
  if (n_nodes == 0) return GNF_OK;
  GNF_REQUIRE(rowptr && out, GNF_EINVAL, "%s: null pointer", who);
  k_gather_segment<<<(unsigned)ceil_div(n_nodes, kSegNodes), kSegNodes, 0, stream>>>(
      src, h, rowptr, idx, n_nodes, agg == GNF_AGG_MEAN, out);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_segment_sum(const float* edges, int32_t h, const int32_t* rowptr,
                               const int32_t* perm, int64_t n_nodes, int32_t agg, float* out,
                               void* stream) {
  return segment_common(edges, h, rowptr, perm, n_nodes, agg, out, (cudaStream_t)stream,
                        "gnf_segment_sum");
}

extern "C" int gnf_gather_segment_sum(const float* x, int32_t h, const int32_t* rowptr,
                                      const int32_t* csr_senders, int64_t n_nodes, int32_t agg,
                                      float* out, void* stream) {
  return segment_common(x, h, rowptr, csr_senders, n_nodes, agg, out, (cudaStream_t)stream,
                        "gnf_gather_segment_sum");
}
