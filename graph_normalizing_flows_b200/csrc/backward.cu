// Analytic, reversible backward of the GRevNet density pass.
//
// loss = -scale * (sum_n log N(z_n; 0, I) + log_det_jacobian)           (run_grevnet.py:292-296,
//                                                                          per-node: scale = 1/N)
// Nothing is stored by the forward pass: starting from z, every half coupling step is undone with
// the inverse update (gnn.py:359,372) -- exactly what the missing GNFBlock(use_efficient_backprop)
// did by name (run_grevnet.py:46,288) -- its s/t MLPs are recomputed layer by layer with the
// activations kept for this half step only, and the gradients are propagated analytically:
//
//   forward half step:  s = S(xa), t = T(xa);  xb' = xb * exp(s) + t;  ldj += sum(s)
//   backward:           xb   = (xb' - t) * exp(-s)
//                       g_s  = g_xb' * xb * exp(s) - scale ;  g_t = g_xb' ;  g_xb = g_xb' * exp(s)
//                       g_xa += dS^T g_s + dT^T g_t  (through h = [xa, agg(xa)] or eps*xa + agg(xa);
//                                                     the aggregation's transpose walks a CSR by sender)
//                       dW_l += a_{l-1}^T delta_l ;  db_l += colsum(delta_l)
//
// fp32 FFMA kernels (split-K weight-gradient GEMM with a fixed-order reduction; no float atomics).
#include "common.cuh"

namespace gnf {
namespace {

static inline int pad_to(int x, int a) { return (x + a - 1) / a * a; }

constexpr int BM = 128, BK = 8;
constexpr int kSplit = 74;       // node ranges of the split-K weight-gradient GEMM (x tiles = multiple of 148 SMs)

// ---- dX: C[M,N] (+)= (A[M,K] @ Wt[K,N]) * act'(aux[M,N]) -----------------------------------
// mask: 0 none, 1 leaky_relu' (aux > 0 ? 1 : 0.2), 2 relu' (aux > 0 ? 1 : 0)
template <int BN>
__global__ void __launch_bounds__(256)
k_dx(const float* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ aux, float* __restrict__ C,
     int64_t M, int N, int K, int mask, int accumulate) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int a_row = tid >> 1, a_kq = (tid & 1) * 4;
  const bool a_ok = (row0 + a_row) < M;
  const float* a_ptr = A + (row0 + a_row) * K + a_kq;
  constexpr int B_F4 = BK * BN / 4;
  const int b_r = (tid * 4) / BN, b_c = (tid * 4) % BN;
  const bool b_ok = tid < B_F4 && (col0 + b_c) < N;
  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 av = a_ok ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0, 0, 0, 0);
    float4 bv = b_ok ? *reinterpret_cast<const float4*>(Wt + (int64_t)(k0 + b_r) * N + col0 + b_c)
                     : make_float4(0, 0, 0, 0);
    __syncthreads();
    As[a_kq + 0][a_row] = av.x;
    As[a_kq + 1][a_row] = av.y;
    As[a_kq + 2][a_row] = av.z;
    As[a_kq + 3][a_row] = av.w;
    if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[b_r][b_c]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int col = col0 + tx + 16 * j;
    if (col >= N) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = row0 + ty * 8 + i;
      if (row >= M) continue;
      float v = acc[i][j];
      if (mask) {
        const float a = aux[row * N + col];
        v *= (a > 0.f) ? 1.f : (mask == 1 ? 0.2f : 0.f);
      }
      if (accumulate) v += C[row * N + col];
      C[row * N + col] = v;
    }
  }
}

// ---- dW: part[z][m][n] = sum_{node in range z} A[node][m] * D[node][n] ------------------------
template <int BN>
__global__ void __launch_bounds__(256)
k_dw(const float* __restrict__ A, int lda, const float* __restrict__ D, int ldd, int64_t n_nodes, int Mdim, int Ndim,
     float* __restrict__ part) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t per = (n_nodes + gridDim.z - 1) / gridDim.z;
  const int64_t beg = (int64_t)blockIdx.z * per;
  const int64_t end = beg + per < n_nodes ? beg + per : n_nodes;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int a_k = tid >> 5, a_m = (tid & 31) * 4;              // 8 node rows x 128 features
  constexpr int B_F4 = BK * BN / 4;
  const int b_k = (tid * 4) / BN, b_n = (tid * 4) % BN;
  for (int64_t k0 = beg; k0 < end; k0 += BK) {
    float4 av = make_float4(0, 0, 0, 0), bv = make_float4(0, 0, 0, 0);
    if (k0 + a_k < end && m0 + a_m < Mdim) av = *reinterpret_cast<const float4*>(A + (k0 + a_k) * lda + m0 + a_m);
    if (tid < B_F4 && k0 + b_k < end && n0 + b_n < Ndim)
      bv = *reinterpret_cast<const float4*>(D + (k0 + b_k) * ldd + n0 + b_n);
    __syncthreads();
    *reinterpret_cast<float4*>(&As[a_k][a_m]) = av;
    if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* out = part + (int64_t)blockIdx.z * Mdim * Ndim;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= Mdim) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < Ndim) out[(int64_t)m * Ndim + n] = acc[i][j];
    }
  }
}

// dW for a NARROW left operand (A has MR <= 16 columns: the attention projections q / k / v when D/2 is small, where the
// 128 x 128 register tiles of k_dw multiply mostly zeros on a third of the SMs).  Block (z, c): node range z, 128
// columns of D, two row slots; a thread keeps MR sums for its column over every other row of the range (D read
// coalesced, the A row broadcast), the two slots are added in fixed order.  part[z][m][col] as k_dw writes it.
template <int MR>
__global__ void __launch_bounds__(256)
k_dw_skinny(const float* __restrict__ A, const float* __restrict__ D, int ldd, int64_t n_nodes, float* __restrict__ part) {
  __shared__ float red[MR][128];
  const int cl = threadIdx.x & 127, rs = threadIdx.x >> 7;
  const int col = blockIdx.y * 128 + cl;
  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per;
  const int64_t end = beg + per < n_nodes ? beg + per : n_nodes;
  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
  if (col < ldd) {
#pragma unroll 4
    for (int64_t r = beg + rs; r < end; r += 2) {
      const float d = D[r * ldd + col];
      const float4* ar = reinterpret_cast<const float4*>(A + r * MR);
#pragma unroll
      for (int m4 = 0; m4 < MR / 4; ++m4) {
        const float4 a = __ldg(ar + m4);
        acc[4 * m4] = fmaf(a.x, d, acc[4 * m4]);
        acc[4 * m4 + 1] = fmaf(a.y, d, acc[4 * m4 + 1]);
        acc[4 * m4 + 2] = fmaf(a.z, d, acc[4 * m4 + 2]);
        acc[4 * m4 + 3] = fmaf(a.w, d, acc[4 * m4 + 3]);
      }
    }
  }
  if (rs == 1) {
#pragma unroll
    for (int m = 0; m < MR; ++m) red[m][cl] = acc[m];
  }
  __syncthreads();
  if (rs == 0 && col < ldd) {
#pragma unroll
    for (int m = 0; m < MR; ++m) part[((int64_t)blockIdx.x * MR + m) * ldd + col] = acc[m] + red[m][cl];
  }
}

// column sums of D over a node range: part[z][n].  Block (z, c): 32 columns, 8 warps stride the rows of range z
// (one coalesced 128-byte read per warp and row); the 8 partial sums are combined in warp order (fixed order).
__global__ void __launch_bounds__(256)
k_db(const float* __restrict__ D, int ldd, int64_t n_nodes, int Ndim, float* __restrict__ part) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.y * 32 + lane;
  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per;
  const int64_t end = beg + per < n_nodes ? beg + per : n_nodes;
  float s0 = 0.f, s1 = 0.f;
  if (col < Ndim) {
    int64_t r = beg + warp;
    for (; r + 8 < end; r += 16) {
      s0 += D[r * ldd + col];
      s1 += D[(r + 8) * ldd + col];
    }
    if (r < end) s0 += D[r * ldd + col];
  }
  red[warp][lane] = s0 + s1;
  __syncthreads();
  if (warp == 0 && col < Ndim) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][lane];
    part[(int64_t)blockIdx.x * Ndim + col] = s;
  }
}

// grad[m*out + n] += sum_z part[z][m][n]  (fixed order; unpadded destination layout)
__global__ void k_reduce_split(const float* __restrict__ part, int splits, int Mpad, int Npad, int M, int N,
                               float* __restrict__ grad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[((int64_t)z * Mpad + m) * Npad + n];
  grad[i] += s;
}

// the same fold for MANY node ranges and few outputs (k_dw_skinny: up to 296 ranges, as few as 80 outputs): a warp per
// output element, lanes stride the ranges, fixed shuffle tree -- same bits on every run
__global__ void __launch_bounds__(256)
k_reduce_split_warp(const float* __restrict__ part, int splits, int Mpad, int Npad, int M, int N, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= M * N) return;
  const int m = w / N, n = w - m * N;
  float s = 0.f;
  for (int z = lane; z < splits; z += 32) s += part[((int64_t)z * Mpad + m) * Npad + n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) grad[w] += s;
}

// ---- element-wise backward of the affine update ---------------------------------------------
// xb: in = xb' (post-update), out = xb (pre-update).  gxb: in = dL/dxb', out = dL/dxb.
__global__ void k_coupling_bwd(float* __restrict__ xb, float* __restrict__ gxb, const float* __restrict__ s,
                               const float* __restrict__ t, int64_t n, int h, int hp, int sp, int gp, float scale,
                               float* __restrict__ gs, float* __restrict__ gt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  const float sv = s[node * sp + f], tv = t[node * sp + f];
  const float es = expf(sv);
  const float xbp = xb[node * hp + f];
  const float x = __fmul_rn(__fsub_rn(xbp, tv), expf(-sv));      // gnn.py:359
  const float g = gxb[node * hp + f];
  xb[node * hp + f] = x;
  gs[node * gp + f] = g * x * es - scale;                          // d(-scale*ldj)/ds = -scale
  gt[node * gp + f] = g;
  gxb[node * hp + f] = g * es;
}

// ---- transpose of gather + segment reduce: g_xa += [g_h direct] + scatter over out-edges ------
__global__ void __launch_bounds__(256)
k_agg_bwd(const float* __restrict__ gh, int in_pad, int h, int hp, const int32_t* __restrict__ rowptr_s,
          const int32_t* __restrict__ csr_receivers, const int32_t* __restrict__ rowptr_r, int64_t n, int mean,
          int concat, float eps, float* __restrict__ gxa) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  const int col = concat ? h + f : f;                              // where d/d agg lives in g_h
  float acc = 0.f;
  const int32_t end = rowptr_s[node + 1];
  for (int32_t e = rowptr_s[node]; e < end; ++e) {
    const int32_t r = csr_receivers[e];
    float v = gh[(int64_t)r * in_pad + col];
    if (mean) v = v / fmaxf((float)(rowptr_r[r + 1] - rowptr_r[r]), 1.f);
    acc += v;
  }
  const float direct = concat ? gh[node * in_pad + f] : eps * gh[node * in_pad + f];
  gxa[node * hp + f] += direct + acc;
}

__global__ void k_scale_rows(const float* __restrict__ x, int64_t total, float scale, float* __restrict__ g) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) g[i] = scale * x[i];
}

__global__ void k_split_p(const float* __restrict__ x, int64_t n, int d, int h, int hp, float* __restrict__ x0,
                          float* __restrict__ x1) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * hp) return;
  int64_t node = i / hp;
  int f = (int)(i - node * hp);
  x0[i] = (f < h) ? x[node * d + f] : 0.f;
  x1[i] = (f < h) ? x[node * d + h + f] : 0.f;
}

__global__ void k_merge_p(const float* __restrict__ x0, const float* __restrict__ x1, int64_t n, int d, int h, int hp,
                          float* __restrict__ z) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  int64_t node = i / d;
  int f = (int)(i - node * d);
  z[i] = (f < h) ? x0[node * hp + f] : x1[node * hp + f - h];
}


// ---- f1: backward of DMSelfAttention (gnn.py:419-477) -----------------------------------------
// keys = project_q (at the sender), queries = project_k (at the receiver), as in k_dm_attn.
// Pass 1, thread per (receiver r, head): with the segment max / sum the recomputed forward left in `stats`,
// dot = sum_e w_e g_w_e with g_w_e = <g_att[r,h,:], values[s_e,:]>, and
// g_queries[r,h,:] = sum_e g_l_e keys[s_e,h,:], g_l_e = w_e (g_w_e - dot) * inv_scale.
__global__ void __launch_bounds__(128)
k_attn_bwd_recv(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                const float* __restrict__ gatt, int qk_pad, int v_pad, int hv_pad, int heads, int kq, int vd,
                float inv_scale, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_senders, int64_t n,
                float* __restrict__ stats, float* __restrict__ gqueries, const int32_t* __restrict__ only) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * heads) return;
  const int64_t r = i / heads;
  if (only && !only[r >> 5]) return;       // second pass: only the 32-receiver groups k_attn_bwd_block handed back
  const int h = (int)(i - r * heads);
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  const float* qr = queries + r * qk_pad + h * kq;
  const float* ga = gatt + r * hv_pad + h * vd;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  const float mx = stats[i * 3], rsum = 1.f / stats[i * 3 + 1];      // segment max / sum from the recomputed forward
  float dot = 0.f;
  for (int32_t e = beg; e < end; ++e) {
    const int32_t s = csr_senders[e];
    const float* ks = keys + (int64_t)s * qk_pad + h * kq;
    float l = 0.f;
    for (int d = 0; d < kq; ++d) l = fmaf(ks[d], qr[d], l);
    const float w = expf(l * inv_scale - mx) * rsum;
    const float* vs = vals + (int64_t)s * v_pad;
    float gw = 0.f;
    for (int c = 0; c < vd; ++c) gw = fmaf(ga[c], vs[c], gw);
    dot = fmaf(w, gw, dot);
  }
  for (int32_t e = beg; e < end; ++e) {
    const int32_t s = csr_senders[e];
    const float* ks = keys + (int64_t)s * qk_pad + h * kq;
    float l = 0.f;
    for (int d = 0; d < kq; ++d) l = fmaf(ks[d], qr[d], l);
    const float w = expf(l * inv_scale - mx) * rsum;
    const float* vs = vals + (int64_t)s * v_pad;
    float gw = 0.f;
    for (int c = 0; c < vd; ++c) gw = fmaf(ga[c], vs[c], gw);
    const float gl = w * (gw - dot) * inv_scale;
#pragma unroll
    for (int d = 0; d < 64; ++d)
      if (d < kq) acc[d] = fmaf(gl, ks[d], acc[d]);
  }
  stats[i * 3 + 2] = dot;
#pragma unroll
  for (int d = 0; d < 64; ++d)
    if (d < kq) gqueries[r * qk_pad + h * kq + d] = acc[d];
}

// Pass 2, thread per (sender s, head): walks the out-edges (CSR by sender) and rebuilds each edge's
// weight from the receiver's statistics:  g_keys[s,h,:] = sum g_l_e queries[r_e,h,:],
// g_vh[s,h,:] = sum w_e g_att[r_e,h,:]  (the value projection is shared by the heads: summed after).
template <int VMAX>
__global__ void __launch_bounds__(128)
k_attn_bwd_send(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                const float* __restrict__ gatt, const float* __restrict__ stats, int qk_pad, int v_pad, int hv_pad,
                int heads, int kq, int vd, float inv_scale, const int32_t* __restrict__ rowptr_s,
                const int32_t* __restrict__ csr_receivers, int64_t n, float* __restrict__ gkeys,
                float* __restrict__ gvh, const int32_t* __restrict__ only) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * heads) return;
  const int64_t s = i / heads;
  if (only && !only[s >> 5]) return;
  const int h = (int)(i - s * heads);
  const float* ks = keys + s * qk_pad + h * kq;
  const float* vs = vals + s * v_pad;
  float acc[64], accv[VMAX];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
#pragma unroll
  for (int c = 0; c < VMAX; ++c) accv[c] = 0.f;
  const int32_t end = rowptr_s[s + 1];
  for (int32_t e = rowptr_s[s]; e < end; ++e) {
    const int64_t r = csr_receivers[e];
    const float* qr = queries + r * qk_pad + h * kq;
    const float* ga = gatt + r * hv_pad + h * vd;
    const float* st = stats + (r * heads + h) * 3;
    float l = 0.f;
    for (int d = 0; d < kq; ++d) l = fmaf(ks[d], qr[d], l);
    const float w = expf(l * inv_scale - st[0]) / st[1];
    float gw = 0.f;
    for (int c = 0; c < vd; ++c) gw = fmaf(ga[c], vs[c], gw);
    const float gl = w * (gw - st[2]) * inv_scale;
#pragma unroll
    for (int d = 0; d < 64; ++d)
      if (d < kq) acc[d] = fmaf(gl, qr[d], acc[d]);
#pragma unroll
    for (int c = 0; c < VMAX; ++c)
      if (c < vd) accv[c] = fmaf(w, ga[c], accv[c]);
  }
#pragma unroll
  for (int d = 0; d < 64; ++d)
    if (d < kq) gkeys[s * qk_pad + h * kq + d] = acc[d];
#pragma unroll
  for (int c = 0; c < VMAX; ++c)
    if (c < vd) gvh[s * hv_pad + h * vd + c] = accv[c];
}

// Warp per (node, head) forms of the two passes for wide keys / values (the embedding flow: 1 head, kq = v = 64, fully
// connected graphs of 100..400 nodes -- a thread per (node, head) is N threads in all, each walking hundreds of edges x
// 256 scalar loads, and took two thirds of that flow's training step).  32 edges at a time: lane = edge for the per-edge
// terms (l, w, g_w, g_l: the own row broadcast out of shared memory, the other end's row as 16-byte loads), then lane =
// output column for the sums over the edges, g_l / w and the other end's index handed round by shuffle so that every
// row is one coalesced read.  Same formulas as k_attn_bwd_recv / k_attn_bwd_send; sums in ascending edge order per column.
__device__ __forceinline__ float attn_row_dot(const float* __restrict__ g, const float* s, int len, bool vec) {
  float a = 0.f;
  if (vec) {
    for (int d = 0; d < len; d += 4) {
      const float4 x = *reinterpret_cast<const float4*>(g + d);
      const float4 y = *reinterpret_cast<const float4*>(s + d);
      a = fmaf(x.x, y.x, a); a = fmaf(x.y, y.y, a); a = fmaf(x.z, y.z, a); a = fmaf(x.w, y.w, a);
    }
  } else {
    for (int d = 0; d < len; ++d) a = fmaf(g[d], s[d], a);
  }
  return a;
}

__global__ void __launch_bounds__(256)
k_attn_bwd_recv_warp(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                     const float* __restrict__ gatt, int qk_pad, int v_pad, int hv_pad, int heads, int kq, int vd,
                     float inv_scale, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_senders,
                     int64_t n, float* __restrict__ stats, float* __restrict__ gqueries, const int32_t* __restrict__ only) {
  __shared__ __align__(16) float q_s[8][64];
  __shared__ __align__(16) float g_s[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= n * heads) return;
  const int64_t r = i / heads;
  if (only && !only[r >> 5]) return;
  const int h = (int)(i - r * heads);
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  const float* qr = queries + r * qk_pad + h * kq;
  const float* ga = gatt + r * hv_pad + h * vd;
  for (int d = lane; d < kq; d += 32) q_s[warp][d] = qr[d];
  for (int c = lane; c < vd; c += 32) g_s[warp][c] = ga[c];
  __syncwarp();
  const bool veck = (kq & 3) == 0 && (qk_pad & 3) == 0, vecv = (vd & 3) == 0 && (v_pad & 3) == 0;
  const float mx = stats[i * 3], rsum = 1.f / stats[i * 3 + 1];      // segment max / sum from the recomputed forward
  float dot = 0.f;
  for (int32_t c0 = beg; c0 < end; c0 += 32) {
    const int32_t e = c0 + lane;
    if (e < end) {
      const int64_t s = csr_senders[e];
      const float w = expf(attn_row_dot(keys + s * qk_pad + h * kq, q_s[warp], kq, veck) * inv_scale - mx) * rsum;
      dot = fmaf(w, attn_row_dot(vals + s * v_pad, g_s[warp], vd, vecv), dot);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  float acc0 = 0.f, acc1 = 0.f;
  for (int32_t c0 = beg; c0 < end; c0 += 32) {
    const int32_t e = c0 + lane;
    const bool valid = e < end;
    const int32_t s = valid ? csr_senders[e] : 0;
    float gl = 0.f;
    if (valid) {
      const float w = expf(attn_row_dot(keys + (int64_t)s * qk_pad + h * kq, q_s[warp], kq, veck) * inv_scale - mx) * rsum;
      const float gw = attn_row_dot(vals + (int64_t)s * v_pad, g_s[warp], vd, vecv);
      gl = w * (gw - dot) * inv_scale;
    }
    const int cnt = min(32, end - c0);
    for (int j = 0; j < cnt; ++j) {
      const float glj = __shfl_sync(0xffffffffu, gl, j);
      const int32_t sj = __shfl_sync(0xffffffffu, s, j);
      const float* kr = keys + (int64_t)sj * qk_pad + h * kq;
      if (lane < kq) acc0 = fmaf(glj, kr[lane], acc0);
      if (lane + 32 < kq) acc1 = fmaf(glj, kr[lane + 32], acc1);
    }
  }
  if (lane == 0) stats[i * 3 + 2] = dot;
  if (lane < kq) gqueries[r * qk_pad + h * kq + lane] = acc0;
  if (lane + 32 < kq) gqueries[r * qk_pad + h * kq + lane + 32] = acc1;
}

__global__ void __launch_bounds__(256)
k_attn_bwd_send_warp(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                     const float* __restrict__ gatt, const float* __restrict__ stats, int qk_pad, int v_pad, int hv_pad,
                     int heads, int kq, int vd, float inv_scale, const int32_t* __restrict__ rowptr_s,
                     const int32_t* __restrict__ csr_receivers, int64_t n, float* __restrict__ gkeys,
                     float* __restrict__ gvh, const int32_t* __restrict__ only) {
  __shared__ __align__(16) float k_s[8][64];
  __shared__ __align__(16) float v_s[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= n * heads) return;
  const int64_t s = i / heads;
  if (only && !only[s >> 5]) return;
  const int h = (int)(i - s * heads);
  const float* ks = keys + s * qk_pad + h * kq;
  const float* vs = vals + s * v_pad;
  for (int d = lane; d < kq; d += 32) k_s[warp][d] = ks[d];
  for (int c = lane; c < vd; c += 32) v_s[warp][c] = vs[c];
  __syncwarp();
  const bool veck = (kq & 3) == 0 && (qk_pad & 3) == 0, vecg = (vd & 3) == 0 && (hv_pad & 3) == 0;
  float acck0 = 0.f, acck1 = 0.f, accv0 = 0.f, accv1 = 0.f;
  const int32_t beg = rowptr_s[s], end = rowptr_s[s + 1];
  for (int32_t c0 = beg; c0 < end; c0 += 32) {
    const int32_t e = c0 + lane;
    const bool valid = e < end;
    const int32_t r = valid ? csr_receivers[e] : 0;
    float w = 0.f, gl = 0.f;
    if (valid) {
      const float* st = stats + ((int64_t)r * heads + h) * 3;
      const float l = attn_row_dot(queries + (int64_t)r * qk_pad + h * kq, k_s[warp], kq, veck);
      w = expf(l * inv_scale - st[0]) / st[1];
      const float gw = attn_row_dot(gatt + (int64_t)r * hv_pad + h * vd, v_s[warp], vd, vecg);
      gl = w * (gw - st[2]) * inv_scale;
    }
    const int cnt = min(32, end - c0);
    for (int j = 0; j < cnt; ++j) {
      const float glj = __shfl_sync(0xffffffffu, gl, j);
      const float wj = __shfl_sync(0xffffffffu, w, j);
      const int32_t rj = __shfl_sync(0xffffffffu, r, j);
      const float* qrow = queries + (int64_t)rj * qk_pad + h * kq;
      const float* grow = gatt + (int64_t)rj * hv_pad + h * vd;
      if (lane < kq) acck0 = fmaf(glj, qrow[lane], acck0);
      if (lane + 32 < kq) acck1 = fmaf(glj, qrow[lane + 32], acck1);
      if (lane < vd) accv0 = fmaf(wj, grow[lane], accv0);
      if (lane + 32 < vd) accv1 = fmaf(wj, grow[lane + 32], accv1);
    }
  }
  if (lane < kq) gkeys[s * qk_pad + h * kq + lane] = acck0;
  if (lane + 32 < kq) gkeys[s * qk_pad + h * kq + lane + 32] = acck1;
  if (lane < vd) gvh[s * hv_pad + h * vd + lane] = accv0;
  if (lane + 32 < vd) gvh[s * hv_pad + h * vd + lane + 32] = accv1;
}

// Block-staged attention backward (round 2), the mirror of k_dm_attn_block.  One template, two passes:
//   SEND = 0 : a CTA owns 32 consecutive RECEIVERS r (CSR by receiver); "other" end of an edge = its sender s.
//              dot[r,h] = sum_e w_e g_w_e  -> stats[.., 2];   g_queries[r,h,:] = sum_e g_l_e keys[s_e,h,:]
//   SEND = 1 : a CTA owns 32 consecutive SENDERS s (CSR by sender); other end = the receiver r.
//              g_keys[s,h,:] = sum_e g_l_e queries[r_e,h,:];  g_vh[s,h,:] = sum_e w_e g_att[r_e,h,:]
// with, per edge and head,  l = <keys[s,h,:], queries[r,h,:]>,  w = exp(l * inv_scale - max[r,h]) / sum[r,h],
// g_w = <g_att[r,h,:], values[s,:]>,  g_l = w (g_w - dot[r,h]) inv_scale      (k_attn_bwd_recv / k_attn_bwd_send).
// The CTA stages its own rows and -- graphs being contiguous node blocks -- the contiguous range [lo, hi] of "other"
// rows with coalesced loads; a warp takes own nodes in turn, 32 edges at a time: lane = edge for the per-edge terms
// (all heads at once, odd row stride: conflict free), then lane = output column for the sums over the edges (the
// per-edge factors pass through shared memory).  Fixed summation order, no atomics.  CTAs whose other rows are not one
// compact range hand their 32 nodes back to the thread-per-head kernels (`fallback`).
constexpr int kAbRecv = 32;              // own nodes per CTA (= 1 << 5: the `only` shift of the thread-per-head kernels)
constexpr int kAbWarps = 16;
constexpr int kAbRows = 128;             // staged "other" rows per CTA
constexpr int kAbIdx = 2048;             // staged CSR indices per CTA
constexpr int kAbH = 8;                  // heads (all in one register pass)
constexpr int kAbCols = 4;               // output columns per lane: heads * kq <= 128, heads * vd <= 128

// "other" row stride: odd (scalar loads, lane = row: conflict free); specialised shapes (vec): 4 mod 8 words, which
// keeps 16-byte loads of 8 consecutive rows on 8 different bank groups
__host__ __device__ inline int ab_ld(int x, bool vec) { return vec ? ((x + 3) / 4 * 4) + (((x + 3) / 4) % 2 ? 0 : 4) : (x | 1); }
template <int SEND>
__host__ __device__ inline size_t attn_bwd_block_bytes(int heads, int kq, int vd, bool vec) {
  const int qk = heads * kq, hv = heads * vd;
  const int other = SEND ? ab_ld(qk + hv + 3 * heads, vec) : ab_ld(qk + (vec ? (vd + 3) / 4 * 4 : vd), vec);
  const int own = SEND ? qk + (vec ? (vd + 3) / 4 * 4 : vd) : qk + hv + 2 * heads;
  return ((size_t)kAbRows * other + (size_t)kAbRecv * own + (size_t)kAbWarps * 32 * kAbH * (SEND ? 2 : 1)) * sizeof(float) +
         ((size_t)kAbIdx + kAbRecv + 1 + 2 * kAbWarps + 2) * sizeof(int32_t);
}

// KQ > 0: specialised on (heads, kq, vd) = (NH, KQ, VD) -- the run_grevnet.py defaults 8 x 10, 10 -- with the per-edge
// loops fully unrolled over 16-byte shared-memory loads; the sender-side value row is then padded to a multiple of 4.
template <int SEND, int KQ, int NH, int VD>
__global__ void __launch_bounds__(kAbWarps * 32)
k_attn_bwd_block(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                 const float* __restrict__ gatt, float* __restrict__ stats, int qk_pad, int v_pad, int hv_pad, int heads,
                 int kq, int vd, float inv_scale, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr,
                 int64_t n, float* __restrict__ out1, float* __restrict__ out2, int32_t* __restrict__ fallback) {
  extern __shared__ float sm_ab[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool VEC = KQ > 0;
  const int qk = heads * kq, hv = heads * vd;
  const int vdp = VEC ? (vd + 3) / 4 * 4 : vd;            // sender-side rows: [keys qk | values vdp]
  const int o_ld = SEND ? ab_ld(qk + hv + 3 * heads, VEC) : ab_ld(qk + vdp, VEC);     // "other" row stride
  const int w_ld = SEND ? qk + vdp : qk + hv + 2 * heads;                               // own row stride
  float* o_s = sm_ab;                                     // [kAbRows][o_ld]
  float* w_s = o_s + kAbRows * o_ld;                      // [kAbRecv][w_ld]
  float* p_s = w_s + kAbRecv * w_ld;                      // [warps][32][kAbH] (x2 when SEND)
  int32_t* idx_s = reinterpret_cast<int32_t*>(p_s + kAbWarps * 32 * kAbH * (SEND ? 2 : 1));   // [kAbIdx]
  int32_t* row_s = idx_s + kAbIdx;                        // [kAbRecv + 1]
  int32_t* red_s = row_s + kAbRecv + 1;                   // [2 * warps + 2]
  const int64_t r0 = (int64_t)blockIdx.x * kAbRecv;
  const int nr = (int)((n - r0) < kAbRecv ? (n - r0) : kAbRecv);
  for (int i = tid; i <= nr; i += kAbWarps * 32) row_s[i] = rowptr[r0 + i];
  __syncthreads();
  const int32_t e0 = row_s[0], ne = row_s[nr] - e0;
  int32_t lo = 0x7fffffff, hi = -1;
  if (ne <= kAbIdx)
    for (int i = tid; i < ne; i += kAbWarps * 32) {
      const int32_t v = csr[e0 + i];
      idx_s[i] = v;
      lo = min(lo, v);
      hi = max(hi, v);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { red_s[warp] = lo; red_s[kAbWarps + warp] = hi; }
  __syncthreads();
  for (int w = 0; w < kAbWarps; ++w) { lo = min(lo, red_s[w]); hi = max(hi, red_s[kAbWarps + w]); }
  const int nrows = ne > 0 ? hi - lo + 1 : 0;
  if (ne > kAbIdx || nrows > kAbRows) {                   // uniform: hand the CTA's nodes to the thread-per-head kernels
    if (tid == 0) fallback[blockIdx.x] = 1;
    return;
  }
  // ---- stage (a warp copies whole rows: coalesced, no index arithmetic per element) ---------------------------------
  // receiver-side row: [queries qk | g_att hv | max heads | 1/sum heads | (SEND: dot heads)]; sender-side row: [keys qk | values vd]
  for (int row = warp; row < nrows + nr; row += kAbWarps) {
    const bool is_other = row < nrows;
    const int64_t node = is_other ? (int64_t)lo + row : r0 + (row - nrows);
    float* dst = is_other ? o_s + row * o_ld : w_s + (row - nrows) * w_ld;
    const bool recv_side = is_other ? (SEND != 0) : (SEND == 0);
    if (recv_side) {
      for (int c = lane; c < qk; c += 32) dst[c] = queries[node * qk_pad + c];
      for (int c = lane; c < hv; c += 32) dst[qk + c] = gatt[node * hv_pad + c];
      for (int c = lane; c < heads; c += 32) {
        const float* st = stats + (node * heads + c) * 3;
        dst[qk + hv + c] = st[0];
        dst[qk + hv + heads + c] = 1.f / st[1];
        if (SEND) dst[qk + hv + 2 * heads + c] = st[2];
      }
    } else {
      for (int c = lane; c < qk; c += 32) dst[c] = keys[node * qk_pad + c];
      for (int c = lane; c < vdp; c += 32) dst[qk + c] = c < vd ? vals[node * v_pad + c] : 0.f;
    }
  }
  __syncthreads();
  float* pw = p_s + warp * 32 * kAbH * (SEND ? 2 : 1);   // [32][kAbH] g_l (and, SEND, [32][kAbH] w behind it)
  int col_h1[kAbCols], col_h2[kAbCols];                   // head of this lane's output columns (out1: / kq, out2: / vd)
#pragma unroll
  for (int j = 0; j < kAbCols; ++j) {
    col_h1[j] = min((lane + 32 * j) / kq, kAbH - 1);
    col_h2[j] = min((lane + 32 * j) / vd, kAbH - 1);
  }
  for (int rl = warp; rl < nr; rl += kAbWarps) {
    const int32_t beg = row_s[rl] - e0, end = row_s[rl + 1] - e0;
    const float* own = w_s + rl * w_ld;
    float w[kAbH], gw[kAbH];
    // per-edge terms of edge e (lane = edge): w[h], gw[h]; invalid lanes get zeros
    auto edge_terms = [&](int32_t e, bool valid) {
      const float* oth = o_s + (valid ? idx_s[e] - lo : 0) * o_ld;
      const float* kp = SEND ? own : oth;                 // keys / values live on the sender side
      const float* rp = SEND ? oth : own;                 // queries / g_att / statistics on the receiver side
      // (opaque to the optimiser: hoisting the own row's ~170 loop-invariant loads out of the edge loop spills)
      asm volatile("" : "+l"(kp), "+l"(rp));
      if constexpr (VEC) {
        static_assert(!VEC || ((KQ * NH) % 4 == 0 && (VD * NH) % 4 == 0 && NH <= kAbH && NH % 4 == 0), "");
        float l[kAbH], g[kAbH];
#pragma unroll
        for (int h = 0; h < kAbH; ++h) { l[h] = 0.f; g[h] = 0.f; }
        const float4* k4 = reinterpret_cast<const float4*>(kp);
        const float4* q4 = reinterpret_cast<const float4*>(rp);
#pragma unroll
        for (int i4 = 0; i4 < KQ * NH / 4; ++i4) {            // same FMA order per head as the generic loop
          const float4 kv = k4[i4], qv = q4[i4];
          l[(4 * i4) / KQ] = fmaf(kv.x, qv.x, l[(4 * i4) / KQ]);
          l[(4 * i4 + 1) / KQ] = fmaf(kv.y, qv.y, l[(4 * i4 + 1) / KQ]);
          l[(4 * i4 + 2) / KQ] = fmaf(kv.z, qv.z, l[(4 * i4 + 2) / KQ]);
          l[(4 * i4 + 3) / KQ] = fmaf(kv.w, qv.w, l[(4 * i4 + 3) / KQ]);
        }
        float vv[(VD + 3) / 4 * 4];
        const float4* v4 = reinterpret_cast<const float4*>(kp + KQ * NH);
#pragma unroll
        for (int i4 = 0; i4 < (VD + 3) / 4; ++i4) {
          const float4 x = v4[i4];
          vv[4 * i4] = x.x; vv[4 * i4 + 1] = x.y; vv[4 * i4 + 2] = x.z; vv[4 * i4 + 3] = x.w;
        }
        const float4* g4 = reinterpret_cast<const float4*>(rp + KQ * NH);
#pragma unroll
        for (int i4 = 0; i4 < VD * NH / 4; ++i4) {
          const float4 x = g4[i4];
          g[(4 * i4) / VD] = fmaf(x.x, vv[(4 * i4) % VD], g[(4 * i4) / VD]);
          g[(4 * i4 + 1) / VD] = fmaf(x.y, vv[(4 * i4 + 1) % VD], g[(4 * i4 + 1) / VD]);
          g[(4 * i4 + 2) / VD] = fmaf(x.z, vv[(4 * i4 + 2) % VD], g[(4 * i4 + 2) / VD]);
          g[(4 * i4 + 3) / VD] = fmaf(x.w, vv[(4 * i4 + 3) % VD], g[(4 * i4 + 3) / VD]);
        }
        const float4* s4 = reinterpret_cast<const float4*>(rp + KQ * NH + VD * NH);    // max[NH] then 1/sum[NH]
        float mxv[kAbH], rsv[kAbH];
#pragma unroll
        for (int i4 = 0; i4 < NH / 4; ++i4) {
          const float4 a = s4[i4], b = s4[NH / 4 + i4];
          mxv[4 * i4] = a.x; mxv[4 * i4 + 1] = a.y; mxv[4 * i4 + 2] = a.z; mxv[4 * i4 + 3] = a.w;
          rsv[4 * i4] = b.x; rsv[4 * i4 + 1] = b.y; rsv[4 * i4 + 2] = b.z; rsv[4 * i4 + 3] = b.w;
        }
#pragma unroll
        for (int h = 0; h < kAbH; ++h) {
          const bool on = valid && h < NH;
          w[h] = on ? expf(l[h] * inv_scale - mxv[h]) * rsv[h] : 0.f;
          gw[h] = on ? g[h] : 0.f;
        }
      } else {
#pragma unroll
        for (int h = 0; h < kAbH; ++h) {
          float l = 0.f, g = 0.f;
          if (h < heads) {
            for (int d = 0; d < kq; ++d) l = fmaf(kp[h * kq + d], rp[h * kq + d], l);
            for (int c = 0; c < vd; ++c) g = fmaf(rp[qk + h * vd + c], kp[qk + c], g);
          }
          const bool on = valid && h < heads;
          w[h] = on ? expf(l * inv_scale - rp[qk + hv + h]) * rp[qk + hv + heads + h] : 0.f;
          gw[h] = on ? g : 0.f;
        }
      }
    };
    float acc1[kAbCols], acc2[kAbCols];
#pragma unroll
    for (int j = 0; j < kAbCols; ++j) { acc1[j] = 0.f; acc2[j] = 0.f; }
    float dot[kAbH];
#pragma unroll
    for (int h = 0; h < kAbH; ++h) dot[h] = 0.f;
    if (!SEND) {
      // pass A over the receiver's in-edges: dot[h] = sum_e w_e g_w_e
      for (int32_t c0 = beg; c0 < end; c0 += 32) {
        edge_terms(c0 + lane, c0 + lane < end);
#pragma unroll
        for (int h = 0; h < kAbH; ++h) dot[h] = fmaf(w[h], gw[h], dot[h]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int h = 0; h < kAbH; ++h) dot[h] += __shfl_xor_sync(0xffffffffu, dot[h], o);
      if (lane < heads) {
        float mine = 0.f;
#pragma unroll
        for (int h = 0; h < kAbH; ++h) mine = (lane == h) ? dot[h] : mine;
        stats[((r0 + rl) * heads + lane) * 3 + 2] = mine;
      }
    }
    const bool one_chunk = end - beg <= 32;
    for (int32_t c0 = beg; c0 < end; c0 += 32) {
      const bool valid = c0 + lane < end;
      if (SEND || !one_chunk) edge_terms(c0 + lane, valid);          // (a single chunk still holds pass A's terms)
      const float* oth = o_s + (valid ? idx_s[c0 + lane] - lo : 0) * o_ld;
      float gl[kAbH];
#pragma unroll
      for (int h = 0; h < kAbH; ++h) {
        const float d = SEND ? ((valid && h < heads) ? oth[qk + hv + 2 * heads + h] : 0.f) : dot[h];
        gl[h] = w[h] * (gw[h] - d) * inv_scale;
      }
      *reinterpret_cast<float4*>(pw + lane * kAbH) = make_float4(gl[0], gl[1], gl[2], gl[3]);
      *reinterpret_cast<float4*>(pw + lane * kAbH + 4) = make_float4(gl[4], gl[5], gl[6], gl[7]);
      if (SEND) {
        *reinterpret_cast<float4*>(pw + 32 * kAbH + lane * kAbH) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(pw + 32 * kAbH + lane * kAbH + 4) = make_float4(w[4], w[5], w[6], w[7]);
      }
      __syncwarp();
      // lane = output column: sums over the chunk's edges, in edge order
      const int cnt = min(32, end - c0);
      for (int j = 0; j < cnt; ++j) {
        const float* orow = o_s + (idx_s[c0 + j] - lo) * o_ld;
        const float* pj = pw + j * kAbH;
#pragma unroll
        for (int u = 0; u < kAbCols; ++u) {
          const int c = lane + 32 * u;
          if (c < qk) acc1[u] = fmaf(pj[col_h1[u]], orow[c], acc1[u]);               // keys (SEND 0) / queries (SEND 1)
          if (SEND && c < hv) acc2[u] = fmaf(pj[32 * kAbH + col_h2[u]], orow[qk + c], acc2[u]);   // g_att
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < kAbCols; ++u) {
      const int c = lane + 32 * u;
      if (c < qk) out1[(r0 + rl) * qk_pad + c] = acc1[u];
      if (SEND && c < hv) out2[(r0 + rl) * hv_pad + c] = acc2[u];
    }
  }
}

// g_v[s, c] = sum over heads of g_vh[s, h, c]   (keras.backend.repeat, gnn.py:528)
__global__ void k_sum_heads(const float* __restrict__ gvh, int hv_pad, int heads, int vd, int v_pad, int64_t n,
                            float* __restrict__ gv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * v_pad) return;
  const int64_t s = i / v_pad;
  const int c = (int)(i - s * v_pad);
  float a = 0.f;
  if (c < vd)
    for (int h = 0; h < heads; ++h) a += gvh[s * hv_pad + h * vd + c];
  gv[i] = a;
}

// dst[n, dpad] = src[n, off : off + cols] zero padded
__global__ void k_take_cols(const float* __restrict__ src, int spad, int off, int cols, int dpad, int64_t n,
                            float* __restrict__ dst) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * dpad) return;
  const int64_t r = i / dpad;
  const int c = (int)(i - r * dpad);
  dst[i] = c < cols ? src[r * spad + off + c] : 0.f;
}

// g_xa += [concat: g_h direct columns] + [residual: top gradient] + g_xq
__global__ void k_attn_gxa(float* __restrict__ gxa, int h, int hp, const float* __restrict__ gh, int in_pad, int concat,
                           const float* __restrict__ gtop, int gp, int residual, const float* __restrict__ gxq, int hp8,
                           int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  float v = gxq[node * hp8 + f];
  if (concat) v += gh[node * in_pad + f];
  if (residual) v += gtop[node * gp + f];
  gxa[node * hp + f] += v;
}

// backward of snt.LayerNorm (gnn.py:554-556) per node row: gy (in: dL/d LN output) becomes dL/d(pre-LN input);
// t = gy * xhat is left for the gamma gradient (column sums), beta's gradient is the column sum of gy (taken before).
__global__ void k_layer_norm_bwd(const float* __restrict__ pre, float* __restrict__ gy, float* __restrict__ t, int64_t n,
                                 int h, int hp, int gp, const float* __restrict__ gb) {
  int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= n) return;
  const float* r = pre + node * hp;
  float* g = gy + node * gp;
  float* tr = t + node * gp;
  float mean = 0.f;
  for (int f = 0; f < h; ++f) mean += r[f];
  mean /= (float)h;
  float var = 0.f;
  for (int f = 0; f < h; ++f) var += (r[f] - mean) * (r[f] - mean);
  var /= (float)h;
  const float inv = 1.f / sqrtf(var + 1e-5f);
  float m1 = 0.f, m2 = 0.f;
  for (int f = 0; f < h; ++f) {
    const float xh = (r[f] - mean) * inv, a = g[f] * gb[f];
    m1 += a;
    m2 += a * xh;
  }
  m1 /= (float)h;
  m2 /= (float)h;
  for (int f = 0; f < h; ++f) {
    const float xh = (r[f] - mean) * inv, gyf = g[f];
    tr[f] = gyf * xh;
    g[f] = inv * (gyf * gb[f] - m1 - xh * m2);
  }
  for (int f = h; f < gp; ++f) tr[f] = 0.f;
}

__global__ void k_add_rows_p(float* __restrict__ out, const float* __restrict__ x, int h, int hp, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  out[node * hp + f] = __fadd_rn(out[node * hp + f], x[node * hp + f]);
}

struct BwdWs {
  float *x0, *x1, *g0, *g1, *hbuf, *gh, *sbuf, *tbuf, *gs, *gt, *d0, *d1, *part;
  float* act[2][kMaxLayers];
  // f1 attention block: per-GNN forward intermediates, shared gradient temporaries
  AttnBufs ab[2];
  float *hbuf2, *gproj, *gatt, *gkeys, *gqueries, *gvh, *gv, *gxq;
  float *preln[2], *lnt;       // GNF_ATTN_LAYER_NORM: pre-LayerNorm GNN outputs, gy * xhat
  size_t part_floats;          // capacity of `part`
  size_t bytes;
};

BwdWs carve_bwd(const Flow& f, int64_t n, void* base) {
  BwdWs w{};
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? (void*)(p + off) : nullptr;
    off += align_up(bytes, 256);
    return (float*)r;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const int lp = pad_to(f.L, 8);
  const int gp = pad_to(f.HP, 8);
  w.x0 = take(nn * f.HP * 4);
  w.x1 = take(nn * f.HP * 4);
  w.g0 = take(nn * f.HP * 4);
  w.g1 = take(nn * f.HP * 4);
  w.hbuf = take(nn * f.in_pad * 4);
  w.gh = take(nn * f.in_pad * 4);
  w.sbuf = take(nn * f.HP * 4);
  w.tbuf = take(nn * f.HP * 4);
  w.gs = take(nn * gp * 4);
  w.gt = take(nn * gp * 4);
  w.d0 = take(nn * lp * 4);
  w.d1 = take(nn * lp * 4);
  for (int m = 0; m < 2; ++m)
    for (int l = 0; l < f.K - 1; ++l) w.act[m][l] = take(nn * lp * 4);
  const int mmax = lp > f.in_pad ? lp : f.in_pad;
  size_t part_elems = (size_t)mmax * lp;
  if (f.attn) {
    for (int m = 0; m < 2; ++m) {
      w.ab[m].xq = take(nn * f.hp8 * 4);
      w.ab[m].qbuf = take(nn * f.qk_pad * 4);
      w.ab[m].kbuf = take(nn * f.qk_pad * 4);
      w.ab[m].vbuf = take(nn * f.v_pad * 4);
      w.ab[m].att = take(nn * f.hv_pad * 4);
      w.ab[m].proj = take(nn * f.cho_pad * 4);
      w.ab[m].stats = take(nn * f.heads * 3 * 4);
      w.ab[m].fallback = (int32_t*)take((nn / 32 + 2) * 4);
    }
    w.hbuf2 = take(nn * f.in_pad * 4);
    w.gproj = take(nn * f.cho_pad * 4);
    w.gatt = take(nn * f.hv_pad * 4);
    w.gkeys = take(nn * f.qk_pad * 4);
    w.gqueries = take(nn * f.qk_pad * 4);
    w.gvh = take(nn * f.hv_pad * 4);
    w.gv = take(nn * f.v_pad * 4);
    w.gxq = take(nn * f.hp8 * 4);
    w.preln[0] = take(nn * f.HP * 4);
    w.preln[1] = take(nn * f.HP * 4);
    w.lnt = take(nn * gp * 4);
    const size_t a1 = (size_t)f.hp8 * f.qk_pad, a2 = (size_t)f.hv_pad * f.cho_pad;
    if (a1 > part_elems) part_elems = a1;
    if (a2 > part_elems) part_elems = a2;
  }
  w.part = take((size_t)kSplit * part_elems * 4);
  w.part_floats = (size_t)kSplit * part_elems;
  w.bytes = off;
  return w;
}

int run_dx(const float* A, const float* Wt, const float* aux, float* C, int64_t M, int N, int K, int mask,
           int accumulate, cudaStream_t stream) {
  if (N <= 16) {
    dim3 grid((unsigned)ceil_div(M, BM), 1);
    k_dx<16><<<grid, 256, 0, stream>>>(A, Wt, aux, C, M, N, K, mask, accumulate);
  } else {
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, 128));
    k_dx<128><<<grid, 256, 0, stream>>>(A, Wt, aux, C, M, N, K, mask, accumulate);
  }
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// dW (+db) of one layer into the flat gradient
// ---- dW of a wide layer on the tensor cores (layered flows): dW = a^T delta as ONE k_gemm_tc call -------------------
// k_gemm_tc multiplies a row-major A [M, K] by a weight IMAGE [K, N]; here M = input features, K = nodes, N = output
// features, so a is transposed into aT [in, n] (rows padded to 4 with zeros) and delta [n, out] is packed into the
// kernel's bf16 hi/lo image format (the same arithmetic k_dw_tc uses for the fused flows: both operands bf16 hi/lo).
__global__ void __launch_bounds__(256)
k_transpose_pad(const float* __restrict__ a, int lda, int64_t n, int cols, float* __restrict__ aT, int64_t n4) {
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                 // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t r = r0 + ty + 8 * j;
    t[ty + 8 * j][tx] = (r < n && c0 + tx < cols) ? a[r * lda + c0 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j;
    const int64_t r = r0 + tx;
    if (c < cols && r < n4) aT[(int64_t)c * n4 + r] = t[tx][ty + 8 * j];
  }
}

// delta [n, >= cols] row-major -> bf16 hi/lo image of the logical weight [k = node, n' = column] in k_gemm_tc's geometry
// (pack.cu kPackTc: column blocks of nb, k16 slabs, K-major 8 x 16-byte core matrices); zero beyond n / cols
__global__ void __launch_bounds__(256)
k_pack_rows_image(const float* __restrict__ D, int ldd, int64_t n, int cols, int kpad, int npad, int nb,
                  uint8_t* __restrict__ img) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)kpad * npad) return;
  const int64_t k = i / npad;
  const int c = (int)(i - k * npad);
  const float v = (k < n && c < cols) ? D[k * ldd + c] : 0.f;
  const int ph = c / nb, nl = c - ph * nb, kl = (int)(k & 15);
  const size_t mat_bytes = (size_t)nb * 32;
  const size_t off = ((size_t)ph * (kpad / 16) + (size_t)(k >> 4)) * (2 * mat_bytes) + (size_t)(kl >> 3) * (nb * 16) +
                     (nl >> 3) * 128 + (nl & 7) * 16 + (kl & 7) * 2;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(h));
  *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(img + off + mat_bytes) = lo;
}

__global__ void k_add_mat(float* __restrict__ grad, const float* __restrict__ c, int rows, int cols, int ldc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), col = (int)(i - (int64_t)r * cols);
  grad[i] += c[(int64_t)r * ldc + col];
}

// returns 1 when the layer was done here, 0 when the caller should use the FFMA kernels, < 0 on error
int run_dw_tc(const Flow& f, int l, const float* a_in, int lda, const float* delta, int ldd, int64_t n, float* part,
              size_t part_floats, float* grad_mlp, cudaStream_t stream) {
  const int in = f.ins[l], out = f.outs[l];
  if (in < 64 || out < 64 || n < 16 || n > (1ll << 30)) return 0;      // narrow layers: the FFMA kernels are cheap there
  const int64_t n4 = (n + 3) / 4 * 4;
  int kpad, npad, nb;
  tc_gemm_geometry((int)n, out, kpad, npad, nb);
  const int outp = (out + 3) / 4 * 4;
  const size_t at_floats = align_up((size_t)in * n4, 64), img_floats = align_up((size_t)kpad * npad, 64),
               c_floats = align_up((size_t)in * outp, 64);
  if (at_floats + img_floats + c_floats > part_floats) return 0;       // scratch = the split-K partial buffer
  float* aT = part;
  uint8_t* img = (uint8_t*)(part + at_floats);
  float* ctmp = part + at_floats + img_floats;
  k_transpose_pad<<<dim3((unsigned)ceil_div(n4, 32), (unsigned)ceil_div(in, 32)), 256, 0, stream>>>(a_in, lda, n, in, aT, n4);
  GNF_LAUNCH_CHECK();
  k_pack_rows_image<<<(unsigned)ceil_div((int64_t)kpad * npad, 256), 256, 0, stream>>>(delta, ldd, n, out, kpad, npad, nb, img);
  GNF_LAUNCH_CHECK();
  int rc = tc_gemm(f, GNF_MATH_TC3X_BF16, aT, (int)n4, (int)n4, img, img, (int)n, out, f.zeros, 2, ctmp, outp, outp, in, stream);
  if (rc) return rc;
  k_add_mat<<<(unsigned)ceil_div((int64_t)in * out, 256), 256, 0, stream>>>(grad_mlp + f.flat_w_off[l], ctmp, in, out, outp);
  GNF_LAUNCH_CHECK();
  return 1;
}

int run_dw(const Flow& f, int l, const float* a_in, int lda, const float* delta, int ldd, int64_t n, float* part,
           float* grad_mlp, cudaStream_t stream, int math = GNF_MATH_FP32, size_t part_floats = 0) {
  const int Mdim = f.in_pads[l], Ndim = ldd;
  int done = 0;
  if (math != GNF_MATH_FP32 && f.tc_layered) {
    done = run_dw_tc(f, l, a_in, lda, delta, ldd, n, part, part_floats, grad_mlp, stream);
    if (done < 0) return done;
  }
  if (done) {
    // weight gradient done on the tensor cores; the bias gradient below
  } else if (Ndim <= 16) {
    dim3 grid((unsigned)ceil_div(Mdim, BM), 1, kSplit);
    k_dw<16><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, Mdim, Ndim, part);
  } else {
    dim3 grid((unsigned)ceil_div(Mdim, BM), (unsigned)ceil_div(Ndim, 128), kSplit);
    k_dw<128><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, Mdim, Ndim, part);
  }
  if (!done) {
    GNF_LAUNCH_CHECK();
    k_reduce_split<<<(unsigned)ceil_div((int64_t)f.ins[l] * f.outs[l], 256), 256, 0, stream>>>(
        part, kSplit, Mdim, Ndim, f.ins[l], f.outs[l], grad_mlp + f.flat_w_off[l]);
    GNF_LAUNCH_CHECK();
  }
  k_db<<<dim3(kSplit, (unsigned)ceil_div(Ndim, 32)), 256, 0, stream>>>(delta, ldd, n, Ndim, part);
  GNF_LAUNCH_CHECK();
  k_reduce_split<<<(unsigned)ceil_div(f.outs[l], 256), 256, 0, stream>>>(part, kSplit, 1, Ndim, 1, f.outs[l],
                                                                        grad_mlp + f.flat_b_off[l]);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// part/grad of one bias-free projection: grad[in, out] += a_in^T delta
int run_dw_generic(const float* a_in, int lda, int in_real, const float* delta, int ldd, int out_real, int64_t n,
                   float* part, float* grad, cudaStream_t stream, size_t part_capacity = 0) {
  if ((lda == 8 || lda == 16) && part_capacity >= (size_t)lda * ldd) {
    // narrow left operand: many small node ranges (the part buffer holds part_capacity floats)
    int64_t nz = (int64_t)(part_capacity / ((size_t)lda * ldd));
    if (nz > 296) nz = 296;
    if (nz > ceil_div(n, 32)) nz = ceil_div(n, 32);
    if (nz < 1) nz = 1;
    dim3 grid((unsigned)nz, (unsigned)ceil_div(ldd, 128));
    if (lda == 8) k_dw_skinny<8><<<grid, 256, 0, stream>>>(a_in, delta, ldd, n, part);
    else k_dw_skinny<16><<<grid, 256, 0, stream>>>(a_in, delta, ldd, n, part);
    GNF_LAUNCH_CHECK();
    k_reduce_split_warp<<<(unsigned)ceil_div((int64_t)in_real * out_real, 8), 256, 0, stream>>>(part, (int)nz, lda, ldd,
                                                                                               in_real, out_real, grad);
    GNF_LAUNCH_CHECK();
    return GNF_OK;
  }
  if (ldd <= 16) {
    dim3 grid((unsigned)ceil_div(lda, BM), 1, kSplit);
    k_dw<16><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, lda, ldd, part);
  } else {
    dim3 grid((unsigned)ceil_div(lda, BM), (unsigned)ceil_div(ldd, 128), kSplit);
    k_dw<128><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, lda, ldd, part);
  }
  GNF_LAUNCH_CHECK();
  k_reduce_split<<<(unsigned)ceil_div((int64_t)in_real * out_real, 256), 256, 0, stream>>>(part, kSplit, lda, ldd,
                                                                                          in_real, out_real, grad);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// forward of one MLP keeping every hidden activation
int mlp_forward_keep(const Flow& f, int mlp, const BwdWs& w, int m, const float* h_in, float* out, int64_t n,
                     cudaStream_t stream, int math = GNF_MATH_FP32) {
  const float* base = f.w32 + (int64_t)mlp * f.w32_per_mlp;
  const float* in = h_in;
  const bool tc = math != GNF_MATH_FP32 && f.tc_layered;      // layered flows: the recompute in k_gemm_tc, as the forward
  for (int l = 0; l < f.K; ++l) {
    const bool last = l == f.K - 1;
    float* dst = last ? out : w.act[m][l];
    int rc;
    if (tc) {
      const size_t off = (size_t)mlp * f.wgemm_per_mlp + f.gemm_off[l];
      rc = tc_gemm(f, math, in, f.in_pads[l], f.in_pads[l], f.wgemm[0] + off, f.wgemm[1] + off, f.ins[l], f.outs[l],
                   base + f.b32_layer_off[l], last ? 2 : f.d.act, dst, f.out_pads[l], f.out_pads[l], n, stream);
    } else
    rc = fwd_linear(in, base + f.w32_layer_off[l], base + f.b32_layer_off[l], dst, n, f.out_pads[l],
                        f.in_pads[l], last ? 2 : f.d.act, stream);
    if (rc) return rc;
    in = dst;
  }
  return GNF_OK;
}

// x *= act'(a) elementwise (a = the stored activation): what k_dx does in its epilogue, for the dX GEMMs run in k_gemm_tc
__global__ void k_mask_rows(float* __restrict__ x, const float* __restrict__ a, int64_t total, int mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total && !(a[i] > 0.f)) x[i] *= (mask == 1 ? 0.2f : 0.f);
}

// backward of one MLP: top gradient g_top [n, gp]; accumulates into gh and the flat grads.  `math` != FP32 (layered
// flows): dX = delta W_l^T in k_gemm_tc from the transposed bf16 hi/lo images (gradients would flush in fp16), the
// activation derivative applied by k_mask_rows; dW stays on the FFMA kernels
int mlp_backward(const Flow& f, int mlp, const BwdWs& w, int m, const float* h_in, const float* g_top, int gp, int64_t n,
                 int accumulate_gh, float* grads, cudaStream_t stream, int math = GNF_MATH_FP32) {
  const float* wt = f.w32T + (int64_t)mlp * f.w32T_per_mlp;
  float* grad_mlp = grads + (int64_t)mlp * f.params_per_mlp;
  const int mask = f.d.act == GNF_ACT_LEAKY_RELU ? 1 : 2;
  const float* delta = g_top;
  int ldd = gp;
  for (int l = f.K - 1; l >= 0; --l) {
    const float* a_in = l == 0 ? h_in : w.act[m][l - 1];
    const int lda = f.in_pads[l];
    // weight gradient needs delta with leading dimension == out_pads[l]: the top gradient buffer is
    // [n, gp] with gp = pad8(HP) >= HP; its extra columns are zero, so Ndim = ldd works for both
    int rc = run_dw(f, l, a_in, lda, delta, ldd, n, w.part, grad_mlp, stream, math, w.part_floats);
    if (rc) return rc;
    float* dst = l == 0 ? w.gh : ((l & 1) ? w.d1 : w.d0);
    if (math != GNF_MATH_FP32 && f.wgemmT && !(l == 0 && accumulate_gh)) {
      const uint8_t* img = f.wgemmT + (size_t)mlp * f.wgemmT_per_mlp + f.gemmT_off[l];
      rc = tc_gemm(f, GNF_MATH_TC3X_BF16, delta, ldd, ldd, img, img, f.outs[l], f.ins[l], f.zeros, 2, dst, f.in_pads[l],
                   f.in_pads[l], n, stream);
      if (rc) return rc;
      if (l > 0) {
        const int64_t total = n * f.in_pads[l];
        k_mask_rows<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(dst, a_in, total, mask);
        GNF_LAUNCH_CHECK();
      }
      delta = dst;
      ldd = f.in_pads[l];
      continue;
    }
    // K of this GEMM = leading dimension of delta (zero-padded rows of W^T beyond out)
    rc = run_dx(delta, wt + f.w32T_layer_off[l], l == 0 ? nullptr : a_in, dst, n, f.in_pads[l],
                l == f.K - 1 ? f.out_pad8[l] : ldd, l == 0 ? 0 : mask, l == 0 ? accumulate_gh : 0, stream);
    if (rc) return rc;
    delta = dst;
    ldd = f.in_pads[l];
  }
  return GNF_OK;
}

// backward of the attention front end of GNN `mlp` (m = 0: s, 1: t): w.gh holds dL/d(MLP input)
int attn_backward(const Flow& f, int mlp, const BwdWs& w, int m, const float* g_top, int gp, int64_t n,
                  const int32_t* rowptr, const int32_t* csr_senders, const int32_t* rowptr_s,
                  const int32_t* csr_receivers, float* ga, float* grads, cudaStream_t stream, int math = GNF_MATH_FP32) {
  const AttnBufs& b = w.ab[m];
  const float* wt = f.wattnT + (int64_t)mlp * f.wattnT_per_mlp;
  float* gm = grads + (int64_t)mlp * f.params_per_mlp;
  const int concat = (f.attn_flags & GNF_ATTN_CONCAT) ? 1 : 0, residual = (f.attn_flags & GNF_ATTN_RESIDUAL) ? 1 : 0;
  const int qk = f.heads * f.kq, hv = f.heads * f.vd;
  const float inv_scale = (f.attn_flags & GNF_ATTN_KQ_DIV) ? 1.f / sqrtf((float)f.kq) : 1.f;
  // g_proj = columns of g_h that multiplied new_node_proj's output (gnn.py:547-548)
  k_take_cols<<<(unsigned)ceil_div(n * f.cho_pad, 256), 256, 0, stream>>>(w.gh, f.in_pad, concat ? f.H : 0, f.cho,
                                                                         f.cho_pad, n, w.gproj);
  GNF_LAUNCH_CHECK();
  int rc = run_dw_generic(b.att, f.hv_pad, hv, w.gproj, f.cho_pad, f.cho, n, w.part,
                          gm + 2ll * f.H * qk + (int64_t)f.H * f.vd, stream);                    // dWo
  if (rc) return rc;
  if (math != GNF_MATH_FP32 && f.lin_off[6] >= 0) {                                              // g_att = g_proj Wo^T
    const size_t off = (size_t)mlp * f.wlin_per_mlp + f.lin_off[6];
    rc = tc_linear(f, GNF_MATH_TC3X_BF16, w.gproj, f.cho_pad, f.cho_pad, f.wlin[0] + off, f.wlin[1] + off, f.cho, hv,
                   f.zeros, 2, w.gatt, f.hv_pad, f.hv_pad, n, stream);
  } else {
    rc = run_dx(w.gproj, wt + f.woT_off, nullptr, w.gatt, n, f.hv_pad, f.cho_pad, 0, 0, stream);
  }
  if (rc) return rc;
  const unsigned blocks = (unsigned)ceil_div(n * f.heads, 128);
  if (f.qk_pad != qk) {       // dX reads qk_pad columns of these (zero weight rows under the pad): pad columns = 0, not NaN
    GNF_CUDA(cudaMemsetAsync(w.gkeys, 0, (size_t)n * f.qk_pad * 4, stream));
    GNF_CUDA(cudaMemsetAsync(w.gqueries, 0, (size_t)n * f.qk_pad * 4, stream));
  }
  // block-staged kernels first (compact graphs); the thread-per-head kernels then serve the groups handed back, or
  // everything when the shape does not fit the staged layout.  The receiver pass leaves dot[r,h] in stats for the
  // sender pass, so each pass completes (staged + handed-back groups) before the next starts.
  const bool vec = f.heads == 8 && f.kq == 10 && f.vd == 10;      // the specialised instantiation (run_grevnet.py defaults)
  const size_t sm0 = attn_bwd_block_bytes<0>(f.heads, f.kq, f.vd, vec), sm1 = attn_bwd_block_bytes<1>(f.heads, f.kq, f.vd, vec);
  const bool staged = b.fallback && f.heads <= kAbH && qk <= 32 * kAbCols && hv <= 32 * kAbCols && sm0 <= 200 * 1024 &&
                      sm1 <= 200 * 1024;
  const unsigned nblk = (unsigned)ceil_div(n, kAbRecv);
  const int32_t* only = staged ? b.fallback : nullptr;
  if (staged) {
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
      GNF_CUDA(cudaFuncSetAttribute(k_attn_bwd_block<0, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GNF_CUDA(cudaFuncSetAttribute(k_attn_bwd_block<1, 0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GNF_CUDA(cudaFuncSetAttribute(k_attn_bwd_block<0, 10, 8, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GNF_CUDA(cudaFuncSetAttribute(k_attn_bwd_block<1, 10, 8, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    GNF_CUDA(cudaMemsetAsync(b.fallback, 0, (size_t)nblk * 4, stream));
    auto k0 = vec ? k_attn_bwd_block<0, 10, 8, 10> : k_attn_bwd_block<0, 0, 0, 0>;
    k0<<<nblk, kAbWarps * 32, sm0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, b.stats, f.qk_pad, f.v_pad, f.hv_pad, f.heads,
                                             f.kq, f.vd, inv_scale, rowptr, csr_senders, n, w.gqueries, nullptr, b.fallback);
    GNF_LAUNCH_CHECK();
  }
  // wide keys / values: warp per (node, head) for whatever the staged kernels do not serve (everything, or the groups
  // they handed back); otherwise the thread-per-(node, head) kernels
  const bool warp_k = (f.vd >= 16 || f.kq >= 32) && f.vd <= 64 && f.kq <= 64;
  const unsigned wblocks = (unsigned)ceil_div(n * f.heads, 8);
  if (warp_k)
    k_attn_bwd_recv_warp<<<wblocks, 256, 0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, f.qk_pad, f.v_pad, f.hv_pad, f.heads,
                                                      f.kq, f.vd, inv_scale, rowptr, csr_senders, n, b.stats, w.gqueries, only);
  else
  k_attn_bwd_recv<<<blocks, 128, 0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, f.qk_pad, f.v_pad, f.hv_pad, f.heads, f.kq,
                                              f.vd, inv_scale, rowptr, csr_senders, n, b.stats, w.gqueries, only);
  GNF_LAUNCH_CHECK();
  if (staged) {
    GNF_CUDA(cudaMemsetAsync(b.fallback, 0, (size_t)nblk * 4, stream));
    auto k1 = vec ? k_attn_bwd_block<1, 10, 8, 10> : k_attn_bwd_block<1, 0, 0, 0>;
    k1<<<nblk, kAbWarps * 32, sm1, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, b.stats, f.qk_pad, f.v_pad, f.hv_pad, f.heads,
                                             f.kq, f.vd, inv_scale, rowptr_s, csr_receivers, n, w.gkeys, w.gvh, b.fallback);
    GNF_LAUNCH_CHECK();
  }
  if (warp_k)
    k_attn_bwd_send_warp<<<wblocks, 256, 0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, b.stats, f.qk_pad, f.v_pad, f.hv_pad,
                                                      f.heads, f.kq, f.vd, inv_scale, rowptr_s, csr_receivers, n, w.gkeys,
                                                      w.gvh, only);
  else if (f.vd <= 32)
    k_attn_bwd_send<32><<<blocks, 128, 0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, b.stats, f.qk_pad, f.v_pad, f.hv_pad,
                                                    f.heads, f.kq, f.vd, inv_scale, rowptr_s, csr_receivers, n, w.gkeys,
                                                    w.gvh, only);
  else
    k_attn_bwd_send<64><<<blocks, 128, 0, stream>>>(b.qbuf, b.kbuf, b.vbuf, w.gatt, b.stats, f.qk_pad, f.v_pad, f.hv_pad,
                                                    f.heads, f.kq, f.vd, inv_scale, rowptr_s, csr_receivers, n, w.gkeys,
                                                    w.gvh, only);
  GNF_LAUNCH_CHECK();
  k_sum_heads<<<(unsigned)ceil_div(n * f.v_pad, 256), 256, 0, stream>>>(w.gvh, f.hv_pad, f.heads, f.vd, f.v_pad, n, w.gv);
  GNF_LAUNCH_CHECK();
  // keys = project_q, queries = project_k (gnn.py:531-532)
  rc = run_dw_generic(b.xq, f.hp8, f.H, w.gkeys, f.qk_pad, qk, n, w.part, gm, stream, w.part_floats);           // dWq
  if (rc) return rc;
  rc = run_dw_generic(b.xq, f.hp8, f.H, w.gqueries, f.qk_pad, qk, n, w.part, gm + (int64_t)f.H * qk, stream, w.part_floats);   // dWk
  if (rc) return rc;
  rc = run_dw_generic(b.xq, f.hp8, f.H, w.gv, f.v_pad, f.vd, n, w.part, gm + 2ll * f.H * qk, stream, w.part_floats);           // dWv
  if (rc) return rc;
  rc = run_dx(w.gkeys, wt + f.wqT_off, nullptr, w.gxq, n, f.hp8, f.qk_pad, 0, 0, stream);
  if (rc) return rc;
  rc = run_dx(w.gqueries, wt + f.wkT_off, nullptr, w.gxq, n, f.hp8, f.qk_pad, 0, 1, stream);
  if (rc) return rc;
  rc = run_dx(w.gv, wt + f.wvT_off, nullptr, w.gxq, n, f.hp8, f.v_pad, 0, 1, stream);
  if (rc) return rc;
  k_attn_gxa<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(ga, f.H, f.HP, w.gh, f.in_pad, concat, g_top, gp,
                                                                  residual, w.gxq, f.hp8, n);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace

int bwd_agg_transpose(const Flow& f, const float* gh, int gh_stride, const int32_t* rowptr_s,
                      const int32_t* csr_receivers, const int32_t* rowptr_r, int64_t n, float* gxa,
                      cudaStream_t stream) {
  k_agg_bwd<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(gh, gh_stride, f.H, f.HP, rowptr_s, csr_receivers,
                                                                 rowptr_r, n, f.d.agg == GNF_AGG_MEAN,
                                                                 f.d.block == GNF_BLOCK_CONCAT, f.d.eps, gxa);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

int bwd_split_scale(const float* z, int64_t n, int d, int h, int hp, float scale, float* x0, float* x1, float* g0,
                    float* g1, cudaStream_t stream) {
  const unsigned blocks = (unsigned)ceil_div(n * hp, 256);
  k_split_p<<<blocks, 256, 0, stream>>>(z, n, d, h, hp, x0, x1);
  GNF_LAUNCH_CHECK();
  k_scale_rows<<<blocks, 256, 0, stream>>>(x0, n * hp, scale, g0);      // dL/dz = scale * z
  GNF_LAUNCH_CHECK();
  k_scale_rows<<<blocks, 256, 0, stream>>>(x1, n * hp, scale, g1);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

int bwd_merge(const float* x0, const float* x1, int64_t n, int d, int h, int hp, float* x, cudaStream_t stream) {
  k_merge_p<<<(unsigned)ceil_div(n * d, 256), 256, 0, stream>>>(x0, x1, n, d, h, hp, x);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace gnf

using namespace gnf;

// one reversed half step on the fp32 FFMA path: (xa, xb', g_xa, g_xb') -> (xb, g_xa += ..., g_xb), grads += ...
// (`math` != FP32, layered flows only: the forward recompute -- attention projections and MLP layers -- runs in
//  k_linear_tc / k_gemm_tc exactly as the density pass did; the dX / dW GEMMs and the attention backward stay here)
static int bwd_half_fp32(const Flow& f, const BwdWs& w, int half, int i, const float* xa, float* xb, float* ga, float* gb,
                         int64_t n, const int32_t* rowptr, const int32_t* csr_senders, const int32_t* rowptr_by_sender,
                         const int32_t* csr_receivers, float scale, float* grads, cudaStream_t stream,
                         int math = GNF_MATH_FP32) {
  const int H = f.H, HP = f.HP, gp = pad_to(HP, 8);
  const unsigned eb = (unsigned)ceil_div(n * H, 256);
  const int ms = f.mlp_index(0, half, i), mt = f.mlp_index(1, half, i);
  int rc;
  if (f.attn) {
    // f1: every GNN has its own attention front end, so s and t have different MLP inputs
    float* hin[2] = {w.hbuf, w.hbuf2};
    float* outs[2] = {w.sbuf, w.tbuf};
    const int mm[2] = {ms, mt};
    for (int m = 0; m < 2; ++m) {
      GNF_CUDA(cudaMemsetAsync(hin[m], 0, (size_t)n * f.in_pad * 4, stream));
      rc = fwd_attn_input(f, mm[m], xa, n, rowptr, csr_senders, w.ab[m], hin[m], stream, math);
      if (rc) return rc;
      rc = mlp_forward_keep(f, mm[m], w, m, hin[m], outs[m], n, stream, math);
      if (rc) return rc;
      if (f.attn_flags & GNF_ATTN_RESIDUAL) {                                     // gnn.py:551-552
        k_add_rows_p<<<eb, 256, 0, stream>>>(outs[m], xa, H, HP, n);
        GNF_LAUNCH_CHECK();
      }
      if (f.attn_flags & GNF_ATTN_LAYER_NORM) {                                   // gnn.py:554-556
        GNF_CUDA(cudaMemcpyAsync(w.preln[m], outs[m], (size_t)n * HP * 4, cudaMemcpyDeviceToDevice, stream));
        rc = fwd_layer_norm(f, mm[m], outs[m], n, stream);
        if (rc) return rc;
      }
    }
    k_coupling_bwd<<<eb, 256, 0, stream>>>(xb, gb, w.sbuf, w.tbuf, n, H, HP, HP, gp, scale, w.gs, w.gt);
    GNF_LAUNCH_CHECK();
    float* gtop[2] = {w.gs, w.gt};
    for (int m = 0; m < 2; ++m) {
      if (f.attn_flags & GNF_ATTN_LAYER_NORM) {
        float* gln = grads + (int64_t)mm[m] * f.params_per_mlp + f.ln_off;      // gamma[H] then beta[H]
        k_db<<<dim3(kSplit, (unsigned)ceil_div(gp, 32)), 256, 0, stream>>>(gtop[m], gp, n, gp, w.part);                                  // d beta
        GNF_LAUNCH_CHECK();
        k_reduce_split<<<1, 256, 0, stream>>>(w.part, kSplit, 1, gp, 1, H, gln + H);
        GNF_LAUNCH_CHECK();
        k_layer_norm_bwd<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(w.preln[m], gtop[m], w.lnt, n, H, HP, gp,
                                                                        f.wln + (int64_t)mm[m] * 2 * HP);
        GNF_LAUNCH_CHECK();
        k_db<<<dim3(kSplit, (unsigned)ceil_div(gp, 32)), 256, 0, stream>>>(w.lnt, gp, n, gp, w.part);                                    // d gamma
        GNF_LAUNCH_CHECK();
        k_reduce_split<<<1, 256, 0, stream>>>(w.part, kSplit, 1, gp, 1, H, gln);
        GNF_LAUNCH_CHECK();
      }
      rc = mlp_backward(f, mm[m], w, m, hin[m], gtop[m], gp, n, 0, grads, stream, math);
      if (rc) return rc;
      rc = attn_backward(f, mm[m], w, m, gtop[m], gp, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers, ga,
                         grads, stream);
      if (rc) return rc;
    }
    return GNF_OK;
  }
  rc = fwd_agg_input(f, xa, n, rowptr, csr_senders, w.hbuf, stream);
  if (rc) return rc;
  rc = mlp_forward_keep(f, ms, w, 0, w.hbuf, w.sbuf, n, stream, math);
  if (rc) return rc;
  rc = mlp_forward_keep(f, mt, w, 1, w.hbuf, w.tbuf, n, stream, math);
  if (rc) return rc;
  k_coupling_bwd<<<eb, 256, 0, stream>>>(xb, gb, w.sbuf, w.tbuf, n, H, HP, HP, gp, scale, w.gs, w.gt);
  GNF_LAUNCH_CHECK();
  rc = mlp_backward(f, ms, w, 0, w.hbuf, w.gs, gp, n, 0, grads, stream, math);
  if (rc) return rc;
  rc = mlp_backward(f, mt, w, 1, w.hbuf, w.gt, gp, n, 1, grads, stream, math);
  if (rc) return rc;
  k_agg_bwd<<<eb, 256, 0, stream>>>(w.gh, f.in_pad, H, HP, rowptr_by_sender, csr_receivers, rowptr, n,
                                    f.d.agg == GNF_AGG_MEAN, f.d.block == GNF_BLOCK_CONCAT, f.d.eps, ga);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// One reversed half step of an INJECT flow (MLP input wider than the fused kernels' 16 columns: dm_self_attn without
// residual / layer_norm, message passing with D > 16) with the MLPs on the tensor cores:
//   input assembly (attention front end or aggregation)      -> hin[m]          the forward's own kernels, `math`
//   layer 0 pre-activations                                   -> pre0[m]         k_linear_tc
//   layers 1..K-1 forward, coupling inverse, dX chain, images -> xb, g_xb, d0[m] k_bwd_chain<BINJ>
//   dW / db of every layer                                    -> grads           k_dw_tc + k_dw_reduce
//   g_h = d0[m] W_0^T                                         -> w.gh            k_linear_tc (K = L, bf16 hi/lo split)
//   below layer 0: attention backward (per GNN) or the transposed aggregation (once, on the sum)
static int bwd_half_inject(const Flow& f, const BwdWs& w, void* tc_ws, int half, int i, const float* xa, float* xb,
                           float* ga, float* gb, int64_t n, const int32_t* rowptr, const int32_t* csr_senders,
                           const int32_t* rowptr_by_sender, const int32_t* csr_receivers, float scale, float* grads,
                           int math, cudaStream_t stream) {
  const int H = f.H, gp = pad_to(f.HP, 8);
  const int mm[2] = {f.mlp_index(0, half, i), f.mlp_index(1, half, i)};
  const int dw_parts = (math == GNF_MATH_TC3X || math == GNF_MATH_TC3X_BF16) ? 2 : 1;
  const int fwd_f16 = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 1 : 0;
  float *pre0[2], *d0[2];
  tc_bwd_inject_buffers(f, n, tc_ws, pre0, d0);
  float* hin[2] = {w.hbuf, f.attn ? w.hbuf2 : w.hbuf};
  int rc;
  for (int m = 0; m < 2; ++m) {
    if (f.attn) {
      GNF_CUDA(cudaMemsetAsync(hin[m], 0, (size_t)n * f.in_pad * 4, stream));
      rc = fwd_attn_input(f, mm[m], xa, n, rowptr, csr_senders, w.ab[m], hin[m], stream, math);
      if (rc) return rc;
    } else if (m == 0) {
      GNF_CUDA(cudaMemsetAsync(hin[0], 0, (size_t)n * f.in_pad * 4, stream));
      rc = fwd_agg_input(f, xa, n, rowptr, csr_senders, hin[0], stream);
      if (rc) return rc;
    }
    const float* base = f.w32 + (int64_t)mm[m] * f.w32_per_mlp;
    const size_t off = (size_t)mm[m] * f.wlin_per_mlp + f.lin_off[4];
    rc = tc_linear(f, math, hin[m], f.in_pads[0], f.in_pads[0], f.wlin[0] + off, f.wlin[1] + off, f.in_dim, f.L,
                   base + f.b32_layer_off[0], 2, pre0[m], f.L, f.L, n, stream);
    if (rc) return rc;
  }
  const float* const hin_c[2] = {hin[0], hin[1]};
  rc = tc_half_backward_inject(f, mm[0], mm[1], hin_c, f.in_pads[0], xb, gb, n, scale, grads, tc_ws, dw_parts, fwd_f16,
                               stream);
  if (rc) return rc;
  for (int m = 0; m < 2; ++m) {
    const size_t off = (size_t)mm[m] * f.wlin_per_mlp + f.lin_off[5];
    // gradients are bf16 hi/lo operands everywhere (fp16 would flush them at scale = 1/N)
    rc = tc_linear(f, GNF_MATH_TC3X_BF16, d0[m], f.L, f.L, f.wlin[0] + off, f.wlin[1] + off, f.L, f.in_dim, f.zeros, 2, w.gh,
                   f.in_pad, f.in_pad, n, stream, (!f.attn && m == 1) ? 1 : 0);
    if (rc) return rc;
    if (f.attn) {
      // g_top of attn_backward only feeds the residual connection, which inject flows do not have
      rc = attn_backward(f, mm[m], w, m, w.gs, gp, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers, ga, grads,
                         stream, math);
      if (rc) return rc;
    }
  }
  if (!f.attn) {
    k_agg_bwd<<<(unsigned)ceil_div(n * H, 256), 256, 0, stream>>>(w.gh, f.in_pad, H, f.HP, rowptr_by_sender, csr_receivers,
                                                                 rowptr, n, f.d.agg == GNF_AGG_MEAN,
                                                                 f.d.block == GNF_BLOCK_CONCAT, f.d.eps, ga);
    GNF_LAUNCH_CHECK();
  }
  return GNF_OK;
}

static bool bwd_use_tc(const Flow& f, int math) { return math != GNF_MATH_FP32 && tc_bwd_supported(f); }
static bool bwd_use_inject(const Flow& f, int math) { return math != GNF_MATH_FP32 && tc_bwd_inject_supported(f); }
// inject flows: the fp32 path's buffers (input assembly, attention backward) followed by the tensor-core side's
static size_t inject_fp32_part(const Flow& f, int64_t n) { return align_up(carve_bwd(f, n, nullptr).bytes, 1024); }

extern "C" size_t gnf_grevnet_backward_workspace(const gnf_flow* h, int64_t n_nodes, int32_t math) {
  if (!h || n_nodes < 0) return 0;
  if (bwd_use_tc(h->f, math)) return tc_bwd_workspace(h->f, n_nodes);
  if (bwd_use_inject(h->f, math)) return inject_fp32_part(h->f, n_nodes) + tc_bwd_inject_workspace(h->f, n_nodes);
  return carve_bwd(h->f, n_nodes, nullptr).bytes;
}

extern "C" int gnf_debug_bwd_layout(const gnf_flow* h, int64_t n_nodes, int64_t* out8) {
  GNF_REQUIRE(h && out8 && n_nodes > 0 && tc_bwd_supported(h->f), GNF_EINVAL, "gnf_debug_bwd_layout: bad argument");
  tc_bwd_layout(h->f, n_nodes, out8);
  return GNF_OK;
}

extern "C" int gnf_debug_dw_gemm(const float* a, const float* b, int64_t n, int32_t fa, int32_t fb, int32_t parts,
                                 int32_t n_splits, float* out, void* ws, size_t ws_bytes, void* stream) {
  GNF_REQUIRE(a && b && out && n > 0, GNF_EINVAL, "gnf_debug_dw_gemm: null pointer / empty");
  return tc_dw_gemm_test(a, b, n, fa, fb, parts, n_splits, out, ws, ws_bytes, stream);
}


extern "C" int gnf_grevnet_backward(const gnf_flow* h, const float* z, int64_t n, int64_t e,
                                    const int32_t* rowptr, const int32_t* csr_senders,
                                    const int32_t* rowptr_by_sender, const int32_t* csr_receivers, double loss_scale,
                                    float* grads, float* x_out, int32_t math, void* ws, size_t ws_bytes,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(h && grads, GNF_EINVAL, "gnf_grevnet_backward: null flow/grads");
  GNF_REQUIRE(n >= 0 && e >= 0, GNF_EINVAL, "gnf_grevnet_backward: negative size");
  GNF_REQUIRE(math >= GNF_MATH_FP32 && math <= GNF_MATH_TC2X, GNF_EINVAL, "gnf_grevnet_backward: bad math %d", math);
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(z && rowptr && rowptr_by_sender && (e == 0 || (csr_senders && csr_receivers)), GNF_EINVAL,
              "gnf_grevnet_backward: null pointer");
  const Flow& f = h->f;
  const bool inject = bwd_use_inject(f, math);
  const bool layered = math != GNF_MATH_FP32 && !inject && !tc_bwd_supported(f) && f.tc_layered;
  if (math != GNF_MATH_FP32 && !inject && !layered) {
    GNF_REQUIRE(tc_bwd_supported(f), GNF_EUNSUPPORTED,
                "gnf_grevnet_backward: the tensor-core backward needs a flow shape the fused kernel supports "
                "(gnf_flow_supports_backward); use GNF_MATH_FP32");
    const int dw_parts = (math == GNF_MATH_TC3X || math == GNF_MATH_TC3X_BF16) ? 2 : 1;
    const int fwd_f16 = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 1 : 0;
    return tc_grevnet_backward(f, z, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers, loss_scale, grads, x_out,
                               ws, ws_bytes, dw_parts, fwd_f16, stream_);
  }
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= gnf_grevnet_backward_workspace(h, n, math), GNF_EWORKSPACE,
              "gnf_grevnet_backward: workspace too small or misaligned");
  BwdWs w = carve_bwd(f, n, ws);
  void* tc_ws = inject ? (void*)((uint8_t*)ws + inject_fp32_part(f, n)) : nullptr;
  const int D = f.d.node_embedding_dim, H = f.H, HP = f.HP, gp = pad_to(HP, 8);
  const float scale = (float)loss_scale;
  const unsigned eb = (unsigned)ceil_div(n * H, 256);

  GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));
  GNF_CUDA(cudaMemsetAsync(w.gs, 0, (size_t)n * gp * 4, stream));
  GNF_CUDA(cudaMemsetAsync(w.gt, 0, (size_t)n * gp * 4, stream));
  k_split_p<<<(unsigned)ceil_div(n * HP, 256), 256, 0, stream>>>(z, n, D, H, HP, w.x0, w.x1);
  GNF_LAUNCH_CHECK();
  // dL/dz = scale * z   (L = -scale * sum log N(z;0,I) - scale * ldj)
  k_scale_rows<<<(unsigned)ceil_div(n * HP, 256), 256, 0, stream>>>(w.x0, n * HP, scale, w.g0);
  GNF_LAUNCH_CHECK();
  k_scale_rows<<<(unsigned)ceil_div(n * HP, 256), 256, 0, stream>>>(w.x1, n * HP, scale, w.g1);
  GNF_LAUNCH_CHECK();

  for (int i = f.d.num_timesteps - 1; i >= 0; --i) {
    for (int half = 1; half >= 0; --half) {                 // forward order was half 0 then half 1
      float* xa = half == 0 ? w.x0 : w.x1;
      float* xb = half == 0 ? w.x1 : w.x0;
      float* ga = half == 0 ? w.g0 : w.g1;
      float* gb = half == 0 ? w.g1 : w.g0;
      int rc = inject ? bwd_half_inject(f, w, tc_ws, half, i, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_by_sender,
                                        csr_receivers, scale, grads, math, stream)
                      : bwd_half_fp32(f, w, half, i, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_by_sender,
                                      csr_receivers, scale, grads, stream, layered ? math : GNF_MATH_FP32);
      if (rc) return rc;
    }
  }
  if (x_out) {
    k_merge_p<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, H, HP, x_out);
    GNF_LAUNCH_CHECK();
  }
  return GNF_OK;
}

extern "C" int gnf_coupling_half_backward(const gnf_flow* h, int32_t half, int32_t step, const float* xa, float* xb,
                                          float* ga, float* gb, int64_t n, int64_t e, const int32_t* rowptr,
                                          const int32_t* csr_senders, const int32_t* rowptr_by_sender,
                                          const int32_t* csr_receivers, double loss_scale, float* grads, int32_t math,
                                          void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(h && grads, GNF_EINVAL, "gnf_coupling_half_backward: null flow/grads");
  GNF_REQUIRE(n >= 0 && e >= 0 && (half == 0 || half == 1), GNF_EINVAL, "gnf_coupling_half_backward: bad argument");
  GNF_REQUIRE(step >= 0 && step < h->f.d.num_timesteps, GNF_EINVAL, "gnf_coupling_half_backward: bad step %d", step);
  GNF_REQUIRE(math >= GNF_MATH_FP32 && math <= GNF_MATH_TC2X, GNF_EINVAL, "gnf_coupling_half_backward: bad math %d", math);
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(xa && xb && ga && gb && rowptr && rowptr_by_sender && (e == 0 || (csr_senders && csr_receivers)),
              GNF_EINVAL, "gnf_coupling_half_backward: null pointer");
  const Flow& f = h->f;
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= gnf_grevnet_backward_workspace(h, n, math), GNF_EWORKSPACE,
              "gnf_coupling_half_backward: workspace too small or misaligned");
  if (bwd_use_inject(f, math)) {
    BwdWs w = carve_bwd(f, n, ws);
    GNF_CUDA(cudaMemsetAsync(w.gs, 0, (size_t)n * pad_to(f.HP, 8) * 4, stream));
    return bwd_half_inject(f, w, (uint8_t*)ws + inject_fp32_part(f, n), half, step, xa, xb, ga, gb, n, rowptr, csr_senders,
                           rowptr_by_sender, csr_receivers, (float)loss_scale, grads, math, stream);
  }
  const bool layered = math != GNF_MATH_FP32 && !tc_bwd_supported(f) && f.tc_layered;
  if (math != GNF_MATH_FP32 && !layered) {
    GNF_REQUIRE(tc_bwd_supported(f), GNF_EUNSUPPORTED, "gnf_coupling_half_backward: flow shape needs GNF_MATH_FP32");
    const int dw_parts = (math == GNF_MATH_TC3X || math == GNF_MATH_TC3X_BF16) ? 2 : 1;
    const int fwd_f16 = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 1 : 0;
    return tc_half_backward(f, half, step, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers,
                            loss_scale, grads, ws, dw_parts, fwd_f16, stream_);
  }
  BwdWs w = carve_bwd(f, n, ws);
  const int gp = pad_to(f.HP, 8);
  GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));
  GNF_CUDA(cudaMemsetAsync(w.gs, 0, (size_t)n * gp * 4, stream));
  GNF_CUDA(cudaMemsetAsync(w.gt, 0, (size_t)n * gp * 4, stream));
  return bwd_half_fp32(f, w, half, step, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers,
                       (float)loss_scale, grads, stream, layered ? math : GNF_MATH_FP32);
}
