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

// column sums of D over a node range: part[z][n]
__global__ void k_db(const float* __restrict__ D, int ldd, int64_t n_nodes, int Ndim, float* __restrict__ part) {
  const int64_t per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t beg = (int64_t)blockIdx.x * per;
  const int64_t end = beg + per < n_nodes ? beg + per : n_nodes;
  for (int n = threadIdx.x; n < Ndim; n += blockDim.x) {
    float s = 0.f;
    for (int64_t r = beg; r < end; ++r) s += D[r * ldd + n];
    part[(int64_t)blockIdx.x * Ndim + n] = s;
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

struct BwdWs {
  float *x0, *x1, *g0, *g1, *hbuf, *gh, *sbuf, *tbuf, *gs, *gt, *d0, *d1, *part;
  float* act[2][kMaxLayers];
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
  w.part = take((size_t)kSplit * mmax * lp * 4);
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
int run_dw(const Flow& f, int l, const float* a_in, int lda, const float* delta, int ldd, int64_t n, float* part,
           float* grad_mlp, cudaStream_t stream) {
  const int Mdim = f.in_pads[l], Ndim = ldd;
  if (Ndim <= 16) {
    dim3 grid((unsigned)ceil_div(Mdim, BM), 1, kSplit);
    k_dw<16><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, Mdim, Ndim, part);
  } else {
    dim3 grid((unsigned)ceil_div(Mdim, BM), (unsigned)ceil_div(Ndim, 128), kSplit);
    k_dw<128><<<grid, 256, 0, stream>>>(a_in, lda, delta, ldd, n, Mdim, Ndim, part);
  }
  GNF_LAUNCH_CHECK();
  k_reduce_split<<<(unsigned)ceil_div((int64_t)f.ins[l] * f.outs[l], 256), 256, 0, stream>>>(
      part, kSplit, Mdim, Ndim, f.ins[l], f.outs[l], grad_mlp + f.flat_w_off[l]);
  GNF_LAUNCH_CHECK();
  k_db<<<kSplit, 256, 0, stream>>>(delta, ldd, n, Ndim, part);
  GNF_LAUNCH_CHECK();
  k_reduce_split<<<(unsigned)ceil_div(f.outs[l], 256), 256, 0, stream>>>(part, kSplit, 1, Ndim, 1, f.outs[l],
                                                                        grad_mlp + f.flat_b_off[l]);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// forward of one MLP keeping every hidden activation
int mlp_forward_keep(const Flow& f, int mlp, const BwdWs& w, int m, float* out, int64_t n, cudaStream_t stream) {
  const float* base = f.w32 + (int64_t)mlp * f.w32_per_mlp;
  const float* in = w.hbuf;
  for (int l = 0; l < f.K; ++l) {
    const bool last = l == f.K - 1;
    float* dst = last ? out : w.act[m][l];
    int rc = fwd_linear(in, base + f.w32_layer_off[l], base + f.b32_layer_off[l], dst, n, f.out_pads[l],
                        f.in_pads[l], last ? 2 : f.d.act, stream);
    if (rc) return rc;
    in = dst;
  }
  return GNF_OK;
}

// backward of one MLP: top gradient g_top [n, gp]; accumulates into gh and the flat grads
int mlp_backward(const Flow& f, int mlp, const BwdWs& w, int m, const float* g_top, int gp, int64_t n,
                 int accumulate_gh, float* grads, cudaStream_t stream) {
  const float* wt = f.w32T + (int64_t)mlp * f.w32T_per_mlp;
  float* grad_mlp = grads + (int64_t)mlp * f.params_per_mlp;
  const int mask = f.d.act == GNF_ACT_LEAKY_RELU ? 1 : 2;
  const float* delta = g_top;
  int ldd = gp;
  for (int l = f.K - 1; l >= 0; --l) {
    const float* a_in = l == 0 ? w.hbuf : w.act[m][l - 1];
    const int lda = f.in_pads[l];
    // weight gradient needs delta with leading dimension == out_pads[l]: the top gradient buffer is
    // [n, gp] with gp = pad8(HP) >= HP; its extra columns are zero, so Ndim = ldd works for both
    int rc = run_dw(f, l, a_in, lda, delta, ldd, n, w.part, grad_mlp, stream);
    if (rc) return rc;
    float* dst = l == 0 ? w.gh : ((l & 1) ? w.d1 : w.d0);
    // K of this GEMM = leading dimension of delta (zero-padded rows of W^T beyond out)
    rc = run_dx(delta, wt + f.w32T_layer_off[l], l == 0 ? nullptr : a_in, dst, n, f.in_pads[l],
                l == f.K - 1 ? f.out_pad8[l] : ldd, l == 0 ? 0 : mask, l == 0 ? accumulate_gh : 0, stream);
    if (rc) return rc;
    delta = dst;
    ldd = f.in_pads[l];
  }
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

static bool bwd_use_tc(const Flow& f, int math) { return math != GNF_MATH_FP32 && tc_bwd_supported(f); }

extern "C" size_t gnf_grevnet_backward_workspace(const gnf_flow* h, int64_t n_nodes, int32_t math) {
  if (!h || n_nodes < 0) return 0;
  if (bwd_use_tc(h->f, math)) return tc_bwd_workspace(h->f, n_nodes);
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
  GNF_REQUIRE(!f.attn, GNF_EUNSUPPORTED, "gnf_grevnet_backward: the dm_self_attn block has no backward yet");
  if (math != GNF_MATH_FP32) {
    GNF_REQUIRE(tc_bwd_supported(f), GNF_EUNSUPPORTED,
                "gnf_grevnet_backward: the tensor-core backward needs a flow shape the fused kernel supports "
                "(gnf_flow_supports); use GNF_MATH_FP32");
    const int dw_parts = (math == GNF_MATH_TC3X || math == GNF_MATH_TC3X_BF16) ? 2 : 1;
    const int fwd_f16 = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 1 : 0;
    return tc_grevnet_backward(f, z, n, rowptr, csr_senders, rowptr_by_sender, csr_receivers, loss_scale, grads, x_out,
                               ws, ws_bytes, dw_parts, fwd_f16, stream_);
  }
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= carve_bwd(f, n, nullptr).bytes, GNF_EWORKSPACE,
              "gnf_grevnet_backward: workspace too small or misaligned");
  BwdWs w = carve_bwd(f, n, ws);
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
      const int ms = f.mlp_index(0, half, i), mt = f.mlp_index(1, half, i);
      int rc = fwd_agg_input(f, xa, n, rowptr, csr_senders, w.hbuf, stream);
      if (rc) return rc;
      rc = mlp_forward_keep(f, ms, w, 0, w.sbuf, n, stream);
      if (rc) return rc;
      rc = mlp_forward_keep(f, mt, w, 1, w.tbuf, n, stream);
      if (rc) return rc;
      k_coupling_bwd<<<eb, 256, 0, stream>>>(xb, gb, w.sbuf, w.tbuf, n, H, HP, HP, gp, scale, w.gs, w.gt);
      GNF_LAUNCH_CHECK();
      rc = mlp_backward(f, ms, w, 0, w.gs, gp, n, 0, grads, stream);
      if (rc) return rc;
      rc = mlp_backward(f, mt, w, 1, w.gt, gp, n, 1, grads, stream);
      if (rc) return rc;
      k_agg_bwd<<<eb, 256, 0, stream>>>(w.gh, f.in_pad, H, HP, rowptr_by_sender, csr_receivers, rowptr, n,
                                        f.d.agg == GNF_AGG_MEAN, f.d.block == GNF_BLOCK_CONCAT, f.d.eps, ga);
      GNF_LAUNCH_CHECK();
    }
  }
  if (x_out) {
    k_merge_p<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, H, HP, x_out);
    GNF_LAUNCH_CHECK();
  }
  return GNF_OK;
}
