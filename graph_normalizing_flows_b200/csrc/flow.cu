// GRevNet flow object, fp32 layer-by-layer kernels, forward / inverse drivers, log-prob.
// Reference path: GRevNet.f / GRevNet.g (gnn.py:304-373) over NodeBlockGNN (gnn.py:143-156)
// with ConcatThenMLPBlock / AggThenMLPBlock (gnn.py:100-126) and make_mlp_model (gnn.py:159-180).
#include <stdarg.h>

#include "common.cuh"

namespace gnf {

// ---- error / bookkeeping -------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() {
  static thread_local int64_t c = 0;
  return c;
}
int num_sms() {
  static int cached[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const bool slot = dev >= 0 && dev < kMaxDevices;
  if (slot && cached[dev]) return cached[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (slot) cached[dev] = n;
  return n;
}

namespace {

static inline int pad_to(int x, int a) { return (x + a - 1) / a * a; }

// ---- planar split / merge ------------------------------------------------------------------
__global__ void k_split(const float* __restrict__ x, int64_t n, int d, int h, int hp,
                        float* __restrict__ x0, float* __restrict__ x1) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * hp) return;
  int64_t node = i / hp;
  int f = (int)(i - node * hp);
  x0[i] = (f < h) ? x[node * d + f] : 0.f;
  x1[i] = (f < h) ? x[node * d + h + f] : 0.f;
}

__global__ void k_merge(const float* __restrict__ x0, const float* __restrict__ x1, int64_t n,
                        int d, int h, int hp, float* __restrict__ z) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  int64_t node = i / d;
  int f = (int)(i - node * d);
  z[i] = (f < h) ? x0[node * hp + f] : x1[node * hp + f - h];
}

__global__ void k_pad_rows(const float* __restrict__ x, int64_t n, int h, int hp, float* __restrict__ xp) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * hp) return;
  int64_t node = i / hp;
  int f = (int)(i - node * hp);
  xp[i] = (f < h) ? x[node * h + f] : 0.f;
}

__global__ void k_unpad_rows(const float* __restrict__ xp, int64_t n, int h, int hp, float* __restrict__ x) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  int64_t node = i / h;
  int f = (int)(i - node * h);
  x[i] = xp[node * hp + f];
}

// ---- a3+a4+a5: aggregate and assemble the MLP input ------------------------------------------
// thread per (node, feature<H); serial in-order accumulation over the stable CSR segment.
__global__ void __launch_bounds__(256)
k_agg_input(const float* __restrict__ xa, int h, int hp, const int32_t* __restrict__ rowptr,
            const int32_t* __restrict__ csr_senders, int64_t n, int mean, int concat, float eps,
            int in_pad, float* __restrict__ hbuf) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  int32_t e = rowptr[node];
  const int32_t end = rowptr[node + 1];
  const int32_t cnt = end - e;
  float acc = 0.f;
  for (; e + 4 <= end; e += 4) {
    int32_t i0 = csr_senders[e], i1 = csr_senders[e + 1], i2 = csr_senders[e + 2], i3 = csr_senders[e + 3];
    float v0 = xa[(int64_t)i0 * hp + f];
    float v1 = xa[(int64_t)i1 * hp + f];
    float v2 = xa[(int64_t)i2 * hp + f];
    float v3 = xa[(int64_t)i3 * hp + f];
    acc = __fadd_rn(acc, v0);
    acc = __fadd_rn(acc, v1);
    acc = __fadd_rn(acc, v2);
    acc = __fadd_rn(acc, v3);
  }
  for (; e < end; ++e) acc = __fadd_rn(acc, xa[(int64_t)csr_senders[e] * hp + f]);
  if (mean) acc = __fdiv_rn(acc, fmaxf((float)cnt, 1.f));
  const float self = xa[node * hp + f];
  if (concat) {
    hbuf[node * in_pad + f] = self;          // tf.concat([nodes, agg], 1)   gnn.py:108-109
    hbuf[node * in_pad + h + f] = acc;
  } else {
    hbuf[node * in_pad + f] = __fadd_rn(__fmul_rn(eps, self), acc);   // gnn.py:123-124
  }
}

// ---- f1: DMSelfAttention (gnn.py:385-477) -----------------------------------------------------
// thread per (receiver, head).  keys = project_q at the SENDER, queries = project_k at the RECEIVER
// (the reference passes (values, project_q, project_k) into (values, keys, queries), gnn.py:531-532);
// logits_e = <keys[s_e], queries[r]> (/ sqrt(kq)); softmax over the receiver's in-edges as
// gn.modules._unsorted_segment_softmax (subtract segment max, exp, divide by segment sum);
// attended = sum_e values[s_e] * w_e in ascending edge order.  Empty segments give 0.
template <int VMAX>
__global__ void __launch_bounds__(128)
k_dm_attn(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
          int qk_pad, int v_pad, int hv_pad, int heads, int kq, int vd, float inv_scale,
          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_senders, int64_t n,
          float* __restrict__ att, float* __restrict__ stats, const int32_t* __restrict__ only, int only_shift) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * heads) return;
  const int64_t r = i / heads;
  if (only && !only[r >> only_shift]) return;      // second pass: only the receiver groups k_dm_attn_block handed back
  const int h = (int)(i - r * heads);
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  const float* qr = queries + r * qk_pad + h * kq;
  float acc[VMAX];
#pragma unroll
  for (int c = 0; c < VMAX; ++c) acc[c] = 0.f;
  // one pass over the in-edges with a running maximum (the segment softmax of graph_nets subtracts the segment
  // max, exponentiates and divides by the segment sum; rescaling the running sums when the max moves gives the
  // same value up to rounding)
  float mx = -INFINITY, sum = 0.f;
  for (int32_t e = beg; e < end; ++e) {
    const int32_t s = csr_senders[e];
    const float* ks = keys + (int64_t)s * qk_pad + h * kq;
    float l = 0.f;
    for (int d = 0; d < kq; ++d) l = fmaf(ks[d], qr[d], l);
    l *= inv_scale;
    float p = 1.f;
    if (l > mx) {
      const float sc = expf(mx - l);            // exp(-inf) = 0 on the first edge
      sum *= sc;
#pragma unroll
      for (int c = 0; c < VMAX; ++c)
        if (c < vd) acc[c] *= sc;
      mx = l;
    } else {
      p = expf(l - mx);
    }
    sum += p;
    const float* vs = vals + (int64_t)s * v_pad;
#pragma unroll
    for (int c = 0; c < VMAX; ++c)
      if (c < vd) acc[c] = fmaf(vs[c], p, acc[c]);
  }
  const float inv = end > beg ? 1.f / sum : 0.f;       // empty segments give 0
#pragma unroll
  for (int c = 0; c < VMAX; ++c)
    if (c < vd) att[r * hv_pad + h * vd + c] = acc[c] * inv;
  if (stats) {                                         // kept for the backward: segment max and sum
    stats[i * 3] = end > beg ? mx : 0.f;
    stats[i * 3 + 1] = end > beg ? sum : 1.f;
  }
}

// Warp per (receiver, head): the shapes of the embedding flow (train_grevnet_with_data.py:41-47: 1 head, kq = v = 64, fully
// connected graphs of 100..400 nodes), where a thread per (receiver, head) leaves the GPU almost empty (N threads in
// all, each walking hundreds of edges x 128 scalar loads).  32 in-edges at a time:
//   logits : lane = edge, <keys[s_e, h, :], queries[r, h, :]> with the query row broadcast out of shared memory
//   softmax: running (max, sum), two warp reductions per chunk
//   values : lane = value column (and column + 32); the chunk's edges in ascending order, p_e and s_e by shuffle -- every
//            value row is one coalesced read
// Same mathematics and summation order over a segment as k_dm_attn (serial in edge order, rescaled when the max moves).
__global__ void __launch_bounds__(256)
k_dm_attn_warp(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
               int qk_pad, int v_pad, int hv_pad, int heads, int kq, int vd, float inv_scale,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_senders, int64_t n,
               float* __restrict__ att, float* __restrict__ stats, const int32_t* __restrict__ only, int only_shift) {
  __shared__ __align__(16) float q_s[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= n * heads) return;
  const int64_t r = i / heads;
  if (only && !only[r >> only_shift]) return;
  const int h = (int)(i - r * heads);
  const int32_t beg = rowptr[r], end = rowptr[r + 1];
  const float* qr = queries + r * qk_pad + h * kq;
  for (int d = lane; d < kq; d += 32) q_s[warp][d] = qr[d];
  __syncwarp();
  const bool vec = (kq & 3) == 0 && (qk_pad & 3) == 0;      // key rows of head h start 16-byte aligned
  float mx = -INFINITY, sum = 0.f, acc0 = 0.f, acc1 = 0.f;
  for (int32_t c0 = beg; c0 < end; c0 += 32) {
    const int32_t e = c0 + lane;
    const bool valid = e < end;
    const int32_t s = valid ? csr_senders[e] : 0;
    float l = -INFINITY;
    if (valid) {
      const float* ks = keys + (int64_t)s * qk_pad + h * kq;
      float a = 0.f;
      if (vec) {
        for (int d = 0; d < kq; d += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(ks + d);
          const float4 qv = *reinterpret_cast<const float4*>(&q_s[warp][d]);
          a = fmaf(kv.x, qv.x, a); a = fmaf(kv.y, qv.y, a); a = fmaf(kv.z, qv.z, a); a = fmaf(kv.w, qv.w, a);
        }
      } else {
        for (int d = 0; d < kq; ++d) a = fmaf(ks[d], q_s[warp][d], a);
      }
      l = a * inv_scale;
    }
    float m = l;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float m_new = fmaxf(mx, m);
    const float sc = expf(mx - m_new);                   // exp(-inf) = 0 on the first chunk
    const float p = valid ? expf(l - m_new) : 0.f;
    // the segment sums in ascending edge order, like the thread-per-head kernel (every lane keeps the same `sum`)
    sum *= sc; acc0 *= sc; acc1 *= sc;
    mx = m_new;
    const int cnt = min(32, end - c0);
    for (int j = 0; j < cnt; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const int32_t sj = __shfl_sync(0xffffffffu, s, j);
      const float* vs = vals + (int64_t)sj * v_pad;
      sum += pj;
      if (lane < vd) acc0 = fmaf(vs[lane], pj, acc0);
      if (lane + 32 < vd) acc1 = fmaf(vs[lane + 32], pj, acc1);
    }
  }
  const float inv = end > beg ? 1.f / sum : 0.f;          // empty segments give 0
  if (lane < vd) att[r * hv_pad + h * vd + lane] = acc0 * inv;
  if (lane + 32 < vd) att[r * hv_pad + h * vd + lane + 32] = acc1 * inv;
  if (stats && lane == 0) {
    stats[i * 3] = end > beg ? mx : 0.f;
    stats[i * 3 + 1] = end > beg ? sum : 1.f;
  }
}

// Block-staged attention (round 2).  A CTA owns kAttnRecv consecutive receivers; like k_gather_segment it stages what
// they need ONCE in shared memory with coalesced loads: their CSR index range and -- graphs being contiguous node
// blocks -- the sender rows [lo, hi] of the key and value matrices, plus their own query rows.  Every warp then takes
// receivers in turn, 32 in-edges at a time:
//   logits   : lane = edge; <keys[s_e, h, :], queries[r, h, :]> for up to 8 heads at once out of shared memory (rows
//              padded to an odd stride: conflict free)
//   softmax  : running (max, sum) per head, kept in registers on every lane; the 8 heads' shuffle reductions are
//              interleaved, so their latencies overlap
//   values   : lane = (value column, edge residue): acc[h] += p[e][h] * v[s_e][c]; one 3-way shuffle fold per chunk
// Same mathematics as k_dm_attn (segment softmax = subtract the segment max, exponentiate, divide by the segment sum);
// the summation order over a segment differs, as any parallel reduction's does (fp32 tolerance, not a bit-exact op).
// CTAs whose senders are too spread out, and shapes beyond heads <= 8 / vd <= 10 per pass, use k_dm_attn.
constexpr int kAttnRecv = 32;            // receivers per CTA (= 1 << 5: k_dm_attn's only_shift)
constexpr int kAttnWarps = 8;
constexpr int kAttnRows = 128;           // staged sender rows per CTA
constexpr int kAttnIdx = 2048;           // staged CSR indices per CTA
constexpr int kAttnH = 8;                // heads per register pass

__host__ __device__ inline int attn_odd(int x) { return x | 1; }
// key row stride: odd (scalar loads, lane = row: conflict free), or -- specialised shapes, qk a multiple of 4 -- 4 mod 8
// words, which keeps 16-byte loads of 8 consecutive rows on 8 different bank groups
__host__ __device__ inline int attn_key_ld(int qk, bool vec) { return vec ? ((qk + 3) / 4 * 4) + (((qk + 3) / 4) % 2 ? 0 : 4) : attn_odd(qk); }
__host__ __device__ inline size_t attn_block_bytes(int heads, int kq, int vd, bool vec = false) {
  const int qk = heads * kq;
  return ((size_t)kAttnRows * attn_key_ld(qk, vec) + (size_t)kAttnRows * attn_odd(vd) + (size_t)kAttnRecv * qk +
          (size_t)kAttnWarps * 32 * kAttnH) * sizeof(float) +
         ((size_t)kAttnIdx + kAttnRecv + 1 + 2 * kAttnWarps + 2) * sizeof(int32_t);
}

// KQ > 0: specialised on (heads, kq) = (NH, KQ) -- the run_grevnet.py defaults 8 x 10 -- with the logit loop fully
// unrolled over 16-byte shared-memory loads (40 loads + 80 FMAs per edge instead of ~500 instructions)
template <int KQ, int NH, int MINB>
__global__ void __launch_bounds__(kAttnWarps * 32, MINB)
k_dm_attn_block(const float* __restrict__ keys, const float* __restrict__ queries, const float* __restrict__ vals,
                int qk_pad, int v_pad, int hv_pad, int heads, int kq, int vd, float inv_scale,
                const int32_t* __restrict__ rowptr, const int32_t* __restrict__ csr_senders, int64_t n,
                float* __restrict__ att, float* __restrict__ stats, int32_t* __restrict__ fallback) {
  extern __shared__ float sm_attn[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qk = heads * kq;
  const int ks_ld = attn_key_ld(qk, KQ > 0), v_ld = attn_odd(vd);
  float* k_s = sm_attn;                                   // [kAttnRows][ks_ld]
  float* v_s = k_s + kAttnRows * ks_ld;                   // [kAttnRows][v_ld]
  float* q_s = v_s + kAttnRows * v_ld;                    // [kAttnRecv][qk]
  float* p_s = q_s + kAttnRecv * qk;                      // [warps][32][kAttnH]
  int32_t* idx_s = reinterpret_cast<int32_t*>(p_s + kAttnWarps * 32 * kAttnH);   // [kAttnIdx]
  int32_t* row_s = idx_s + kAttnIdx;                      // [kAttnRecv + 1]
  int32_t* red_s = row_s + kAttnRecv + 1;                 // [2 * warps + 2]
  const int64_t r0 = (int64_t)blockIdx.x * kAttnRecv;
  const int nr = (int)((n - r0) < kAttnRecv ? (n - r0) : kAttnRecv);
  for (int i = tid; i <= nr; i += kAttnWarps * 32) row_s[i] = rowptr[r0 + i];
  __syncthreads();
  const int32_t e0 = row_s[0], ne = row_s[nr] - e0;
  int32_t lo = 0x7fffffff, hi = -1;
  if (ne <= kAttnIdx)
    for (int i = tid; i < ne; i += kAttnWarps * 32) {
      const int32_t v = csr_senders[e0 + i];
      idx_s[i] = v;
      lo = min(lo, v);
      hi = max(hi, v);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { red_s[warp] = lo; red_s[kAttnWarps + warp] = hi; }
  __syncthreads();
  for (int w = 0; w < kAttnWarps; ++w) { lo = min(lo, red_s[w]); hi = max(hi, red_s[kAttnWarps + w]); }
  const int nrows = ne > 0 ? hi - lo + 1 : 0;
  if (ne > kAttnIdx || nrows > kAttnRows) {               // uniform: hand the CTA's receivers to k_dm_attn
    if (tid == 0) fallback[blockIdx.x] = 1;
    return;
  }
  // ---- stage: contiguous global rows [lo, hi] of keys / values, rows [r0, r0 + nr) of queries ------------------
  for (int row = warp; row < nrows + nr; row += kAttnWarps) {      // a warp copies whole rows: coalesced, no division
    if (row < nrows) {
      const int64_t node = (int64_t)lo + row;
      for (int c = lane; c < qk; c += 32) k_s[row * ks_ld + c] = keys[node * qk_pad + c];
      for (int c = lane; c < vd; c += 32) v_s[row * v_ld + c] = vals[node * v_pad + c];
    } else {
      const int rr = row - nrows;
      for (int c = lane; c < qk; c += 32) q_s[rr * qk + c] = queries[(r0 + rr) * qk_pad + c];
    }
  }
  __syncthreads();
  float* pw = p_s + warp * 32 * kAttnH;
  const int vc = lane % 10, vg = lane / 10;               // value phase: column vc (< vd <= 10), edge residue vg (0..2)
  for (int rl = warp; rl < nr; rl += kAttnWarps) {
    const int32_t beg = row_s[rl] - e0, end = row_s[rl + 1] - e0;
    const float* qr = q_s + rl * qk;   // (not const-qualified pointer value: see the asm below)
    for (int h0 = 0; h0 < heads; h0 += kAttnH) {          // up to 8 heads per register pass
      const int nh = min(kAttnH, heads - h0);
      float mx[kAttnH], sum[kAttnH], acc[kAttnH];
#pragma unroll
      for (int h = 0; h < kAttnH; ++h) { mx[h] = -INFINITY; sum[h] = 0.f; acc[h] = 0.f; }
      for (int32_t c0 = beg; c0 < end; c0 += 32) {
        const int32_t e = c0 + lane;
        const bool valid = e < end;
        const int row = valid ? idx_s[e] - lo : 0;
        const float* kr = k_s + row * ks_ld + h0 * kq;
        if constexpr (KQ > 0) asm volatile("" : "+l"(qr));     // keep the query row's loads inside the loop (registers)
        float l[kAttnH];
        if constexpr (KQ > 0) {
          static_assert((KQ * NH) % 4 == 0 && NH <= kAttnH, "");
          float a[kAttnH];
#pragma unroll
          for (int h = 0; h < kAttnH; ++h) a[h] = 0.f;
          const float4* k4 = reinterpret_cast<const float4*>(kr);
          const float4* q4 = reinterpret_cast<const float4*>(qr);
#pragma unroll
          for (int i4 = 0; i4 < KQ * NH / 4; ++i4) {          // same FMA order per head as the generic loop
            const float4 kv = k4[i4], qv = q4[i4];
            a[(4 * i4) / KQ] = fmaf(kv.x, qv.x, a[(4 * i4) / KQ]);
            a[(4 * i4 + 1) / KQ] = fmaf(kv.y, qv.y, a[(4 * i4 + 1) / KQ]);
            a[(4 * i4 + 2) / KQ] = fmaf(kv.z, qv.z, a[(4 * i4 + 2) / KQ]);
            a[(4 * i4 + 3) / KQ] = fmaf(kv.w, qv.w, a[(4 * i4 + 3) / KQ]);
          }
#pragma unroll
          for (int h = 0; h < kAttnH; ++h) l[h] = (valid && h < NH) ? a[h] * inv_scale : -INFINITY;
        } else {
#pragma unroll
          for (int h = 0; h < kAttnH; ++h) {
            float a = 0.f;
            if (h < nh)
              for (int d = 0; d < kq; ++d) a = fmaf(kr[h * kq + d], qr[(h0 + h) * kq + d], a);
            l[h] = (valid && h < nh) ? a * inv_scale : -INFINITY;
          }
        }
        float m[kAttnH];
#pragma unroll
        for (int h = 0; h < kAttnH; ++h) m[h] = l[h];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int h = 0; h < kAttnH; ++h) m[h] = fmaxf(m[h], __shfl_xor_sync(0xffffffffu, m[h], o));
        float pe[kAttnH], ps[kAttnH], sc[kAttnH];
#pragma unroll
        for (int h = 0; h < kAttnH; ++h) {
          const float m_new = fmaxf(mx[h], m[h]);
          sc[h] = h < nh ? expf(mx[h] - m_new) : 0.f;      // exp(-inf) = 0 on the first chunk
          pe[h] = (valid && h < nh) ? expf(l[h] - m_new) : 0.f;
          ps[h] = pe[h];
          mx[h] = m_new;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int h = 0; h < kAttnH; ++h) ps[h] += __shfl_xor_sync(0xffffffffu, ps[h], o);
#pragma unroll
        for (int h = 0; h < kAttnH; ++h) {
          sum[h] = sum[h] * sc[h] + ps[h];
          acc[h] *= sc[h];
        }
        *reinterpret_cast<float4*>(pw + lane * kAttnH) = make_float4(pe[0], pe[1], pe[2], pe[3]);
        *reinterpret_cast<float4*>(pw + lane * kAttnH + 4) = make_float4(pe[4], pe[5], pe[6], pe[7]);
        __syncwarp();
        // values: lane (vc, vg) adds the edges j = vg, vg + 3, ... of this chunk for its column
        if (vg < 3 && vc < vd) {
          const int cnt = min(32, end - c0);
          for (int j = vg; j < cnt; j += 3) {
            const float v = v_s[(idx_s[c0 + j] - lo) * v_ld + vc];
            const float4 p0 = *reinterpret_cast<const float4*>(pw + j * kAttnH);
            const float4 p1 = *reinterpret_cast<const float4*>(pw + j * kAttnH + 4);
            acc[0] = fmaf(p0.x, v, acc[0]); acc[1] = fmaf(p0.y, v, acc[1]); acc[2] = fmaf(p0.z, v, acc[2]);
            acc[3] = fmaf(p0.w, v, acc[3]); acc[4] = fmaf(p1.x, v, acc[4]); acc[5] = fmaf(p1.y, v, acc[5]);
            acc[6] = fmaf(p1.z, v, acc[6]); acc[7] = fmaf(p1.w, v, acc[7]);
          }
        }
        __syncwarp();
      }
      // fold the three edge residues (lanes vc, vc + 10, vc + 20) and write head h0 + h, column vc
#pragma unroll
      for (int h = 0; h < kAttnH; ++h) {
        const float a1 = __shfl_sync(0xffffffffu, acc[h], (lane + 10) & 31);
        const float a2 = __shfl_sync(0xffffffffu, acc[h], (lane + 20) & 31);
        if (lane < 10 && lane < vd && h < nh)
          att[(r0 + rl) * hv_pad + (h0 + h) * vd + lane] = end > beg ? ((acc[h] + a1) + a2) / sum[h] : 0.f;
        if (stats && lane == 0 && h < nh) {
          stats[((r0 + rl) * heads + h0 + h) * 3] = end > beg ? mx[h] : 0.f;
          stats[((r0 + rl) * heads + h0 + h) * 3 + 1] = end > beg ? sum[h] : 1.f;
        }
      }
    }
  }
}

// MLP input of DMSelfAttentionMLP: concat([nodes, proj]) or proj (gnn.py:547-548)
__global__ void k_attn_input(const float* __restrict__ xa, int h, int hp, const float* __restrict__ proj, int cho,
                             int cho_pad, int concat, int in_pad, int64_t n, float* __restrict__ hbuf) {
  const int width = concat ? h + cho : cho;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t node = i / width;
  const int c = (int)(i - node * width);
  float v;
  if (concat) v = c < h ? xa[node * hp + c] : proj[node * cho_pad + c - h];
  else v = proj[node * cho_pad + c];
  hbuf[node * in_pad + c] = v;
}

__global__ void k_add_rows(float* __restrict__ out, const float* __restrict__ x, int h, int hp, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * h) return;
  const int64_t node = i / h;
  const int f = (int)(i - node * h);
  out[node * hp + f] = __fadd_rn(out[node * hp + f], x[node * hp + f]);
}

// snt.LayerNorm over the H features of every node row, in place (gnn.py:554-556): eps 1e-5 [upstream]
__global__ void k_layer_norm(float* __restrict__ x, int64_t n, int h, int hp, const float* __restrict__ gb) {
  int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (node >= n) return;
  float* r = x + node * hp;
  float mean = 0.f;
  for (int f = 0; f < h; ++f) mean += r[f];
  mean /= (float)h;
  float var = 0.f;
  for (int f = 0; f < h; ++f) var += (r[f] - mean) * (r[f] - mean);
  var /= (float)h;
  const float inv = 1.f / sqrtf(var + 1e-5f);
  for (int f = 0; f < h; ++f) r[f] = (r[f] - mean) * inv * gb[f] + gb[hp + f];
}

// ---- a6: one Sonnet Linear (+ activation), fp32 FFMA -----------------------------------------
// C[M,N] = act(A[M,K] @ W[K,N] + b);  K % 8 == 0, N % 4 == 0 (zero-padded operands).
constexpr int BM = 128, BK = 8;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 0) return fmaxf(v, 0.2f * v);   // tf.nn.leaky_relu, alpha = 0.2
  if (act == 1) return fmaxf(v, 0.f);        // tf.nn.relu
  return v;                                  // activate_final=False
}

template <int BN>
__global__ void __launch_bounds__(256)
k_linear(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
         float* __restrict__ C, int64_t M, int N, int K, int act) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
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
  // B tile: BK x BN floats = 8*BN; one float4 per thread covers BN=128; BN=16 needs 32 threads
  constexpr int B_F4 = BK * BN / 4;
  const int b_r = (tid * 4) / BN, b_c = (tid * 4) % BN;
  const bool b_ok = tid < B_F4 && (col0 + b_c) < N;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 av = a_ok ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0, 0, 0, 0);
    float4 bv = b_ok ? *reinterpret_cast<const float4*>(W + (int64_t)(k0 + b_r) * N + col0 + b_c)
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
    const float bj = bias[col];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = row0 + ty * 8 + i;
      if (row < M) C[row * N + col] = apply_act(acc[i][j] + bj, act);
    }
  }
}

// ---- a7: affine coupling update + log-det partials -------------------------------------------
// forward: xb <- xb * exp(s) + t, ldj += sum(s)    (gnn.py:322-323,337-338)
// inverse: xb <- (xb - t) * exp(-s)                (gnn.py:359,372)
__global__ void __launch_bounds__(256)
k_coupling_update(float* __restrict__ xb, const float* __restrict__ s, const float* __restrict__ t,
                  int64_t n, int h, int hp, int sp, int inverse, double* __restrict__ partials) {
  __shared__ double red[8];
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double local = 0.0;
  if (i < n * h) {
    int64_t node = i / h;
    int f = (int)(i - node * h);
    float sv = s[node * sp + f], tv = t[node * sp + f];
    float x = xb[node * hp + f];
    if (!inverse) {
      xb[node * hp + f] = __fadd_rn(__fmul_rn(x, expf(sv)), tv);
      local = (double)sv;
    } else {
      xb[node * hp + f] = __fmul_rn(__fsub_rn(x, tv), expf(-sv));
    }
  }
  if (partials) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w];
      partials[blockIdx.x] = tot;
    }
  }
}

// fixed-order reduction of partials; accum[0] (+)= total
__global__ void __launch_bounds__(256)
k_reduce_partials(const double* __restrict__ partials, int n, double* __restrict__ accum, int add) {
  __shared__ double red[256];
  double local = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) local += partials[i];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) accum[0] = (add ? accum[0] : 0.0) + red[0];
}

// ---- a9: batch-norm bijector pieces (per-feature moments, in-place affine) ------------------
// thread t owns feature t % hp (blockDim is a multiple of hp); fp64 partials, fixed order
__global__ void k_bn_partial(const float* __restrict__ x, int64_t total, int hp, double* __restrict__ partials) {
  extern __shared__ double sm[];                 // [2][blockDim]
  double s = 0.0, ss = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double v = (double)x[i];
    s += v;
    ss += v * v;
  }
  sm[threadIdx.x] = s;
  sm[blockDim.x + threadIdx.x] = ss;
  __syncthreads();
  if ((int)threadIdx.x < hp) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < (int)blockDim.x; k += hp) {
      a += sm[k];
      b += sm[blockDim.x + k];
    }
    partials[(int64_t)blockIdx.x * 2 * hp + threadIdx.x] = a;
    partials[(int64_t)blockIdx.x * 2 * hp + hp + threadIdx.x] = b;
  }
}

__global__ void k_bn_final(const double* __restrict__ partials, int nblocks, int h, int hp, double* __restrict__ out,
                           double count) {
  const int f = threadIdx.x;
  if (f == 0 && count >= 0.0) out[2 * h] = count;          // sums[2H] = N (ranks all-reduce it with the sums)
  if (f >= h) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblocks; ++k) {
    a += partials[(int64_t)k * 2 * hp + f];
    b += partials[(int64_t)k * 2 * hp + hp + f];
  }
  out[f] = a;
  out[h + f] = b;
}

__global__ void k_affine_rows(float* __restrict__ x, int64_t n, int h, int hp, const float* __restrict__ scale,
                              const float* __restrict__ shift) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * hp) return;
  const int f = (int)(i % hp);
  if (f < h) x[i] = __fadd_rn(__fmul_rn(x[i], scale[f]), shift[f]);
}

// bn.forward (sampling direction, gnn.py:356-358,369-371): de-normalise with the MOVING statistics.
// scale = sqrt(moving_var + eps) / gamma, shift = moving_mean - beta * scale, each a separately rounded fp32 operation
__global__ void k_bn_moving_scale_shift(const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ mm, const float* __restrict__ mv, float eps, int h,
                                        float* __restrict__ scale_shift) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= h) return;
  const float sc = __fdiv_rn(__fsqrt_rn(__fadd_rn(mv[f], eps)), gamma[f]);
  scale_shift[f] = sc;
  scale_shift[h + f] = __fsub_rn(mm[f], __fmul_rn(beta[f], sc));
}

// backward of the batch-norm bijector, phase 1: per-feature sums of G_y and G_y * xhat, xhat = (y - beta) / gamma
__global__ void k_bn_bwd_partial(const float* __restrict__ y, const float* __restrict__ g, const float* __restrict__ beta,
                                 const float* __restrict__ inv_gamma, int64_t total, int h, int hp,
                                 double* __restrict__ partials) {
  extern __shared__ double sm[];                 // [2][blockDim]
  double s = 0.0, ss = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int f = threadIdx.x % hp;                // stride and blockDim are multiples of hp
  const float b = f < h ? beta[f] : 0.f, ig = f < h ? inv_gamma[f] : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double gv = (double)g[i];
    s += gv;
    ss += gv * (double)((y[i] - b) * ig);
  }
  sm[threadIdx.x] = s;
  sm[blockDim.x + threadIdx.x] = ss;
  __syncthreads();
  if ((int)threadIdx.x < hp) {
    double a = 0.0, c = 0.0;
    for (int k = threadIdx.x; k < (int)blockDim.x; k += hp) {
      a += sm[k];
      c += sm[blockDim.x + k];
    }
    partials[(int64_t)blockIdx.x * 2 * hp + threadIdx.x] = a;
    partials[(int64_t)blockIdx.x * 2 * hp + hp + threadIdx.x] = c;
  }
}

// phase 2: xhat = (y - beta)/gamma;  g <- c1*g + c2*xhat + c3;  y <- xhat*s + mu   (coef rows: beta, 1/gamma, c1, c2, c3, s, mu)
__global__ void k_bn_bwd_apply(float* __restrict__ y, float* __restrict__ g, int64_t n, int h, int hp,
                               const float* __restrict__ coef) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * hp) return;
  const int f = (int)(i % hp);
  if (f >= h) return;
  const float xh = (y[i] - coef[f]) * coef[h + f];
  g[i] = fmaf(coef[2 * h + f], g[i], fmaf(coef[3 * h + f], xh, coef[4 * h + f]));
  y[i] = fmaf(xh, coef[5 * h + f], coef[6 * h + f]);
}

// [H]-sized bookkeeping of the bijector, density direction: batch moments -> scale/shift of the normalisation, the
// N-tiled log-det term, the moving averages, and the statistics the backward needs.  One block.
__global__ void k_bn_finalize(const double* __restrict__ sums, int h, const float* __restrict__ gamma,
                              const float* __restrict__ beta, double eps, double n_local, double* __restrict__ ldj,
                              float* __restrict__ scale_shift, double* __restrict__ stats, float* __restrict__ moving_mean,
                              float* __restrict__ moving_var, float momentum) {
  __shared__ double red[256];
  const int f = threadIdx.x;
  double term = 0.0;
  if (f < h) {
    const double n = sums[2 * h];
    const double mean = sums[f] / n;
    double var = sums[h + f] / n - mean * mean;          // biased, as tf.nn.moments
    if (var < 0.0) var = 0.0;
    const double g = (double)gamma[f], b = (double)beta[f];
    const double inv = 1.0 / sqrt(var + eps);
    const double sc = g * inv;
    scale_shift[f] = (float)sc;
    scale_shift[h + f] = (float)(b - mean * sc);
    stats[f] = mean;
    stats[h + f] = var;
    if (f == 0) stats[2 * h] = n;
    term = log(g) - 0.5 * log(var + eps);
    if (momentum >= 0.f) {
      moving_mean[f] = moving_mean[f] * momentum + (1.f - momentum) * (float)mean;
      moving_var[f] = moving_var[f] * momentum + (1.f - momentum) * (float)var;
    }
  }
  red[f] = term;
  __syncthreads();
  if (f == 0 && ldj) {
    double s = 0.0;
    for (int k = 0; k < h; ++k) s += red[k];
    ldj[0] += n_local * s;       // scalar ildj tiled over the node axis (event_ndims=2): x N (this rank's share)
  }
}

// backward bookkeeping: coefficient rows of gnf_bn_backward_apply and the gamma / beta gradients
__global__ void k_bn_bwd_coef(const double* __restrict__ sums, const double* __restrict__ stats, int h,
                              const float* __restrict__ gamma, const float* __restrict__ beta, double eps,
                              double loss_scale, float* __restrict__ coef, float* __restrict__ inv_gamma_out,
                              double* __restrict__ g_gamma, double* __restrict__ g_beta) {
  const int f = threadIdx.x;
  if (f >= h) return;
  const double g = (double)gamma[f], b = (double)beta[f];
  if (!sums) {                      // phase 0: only 1/gamma (input of gnf_bn_backward_sums)
    inv_gamma_out[f] = (float)(1.0 / g);
    return;
  }
  const double n = stats[2 * h], mean = stats[f], var = stats[h + f];
  const double s = sqrt(var + eps), s1 = sums[f], s2 = sums[h + f];
  coef[f] = (float)b;
  coef[h + f] = (float)(1.0 / g);
  coef[2 * h + f] = (float)(g / s);
  coef[3 * h + f] = (float)((loss_scale - g * s2 / n) / s);
  coef[4 * h + f] = (float)(-g * s1 / (n * s));
  coef[5 * h + f] = (float)s;
  coef[6 * h + f] = (float)mean;
  g_beta[f] += s1;
  g_gamma[f] += s2 - loss_scale * n / g;
}

// ---- a8: log-prob ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sumsq(const float* __restrict__ z, int64_t total, double* __restrict__ partials) {
  __shared__ double red[8];
  double local = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = (double)z[i];
    local += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    partials[blockIdx.x] = tot;
  }
}

__global__ void k_log_prob_final(const double* __restrict__ partials, int n, const double* __restrict__ ldj,
                                 double n_nodes, int d, double* __restrict__ out) {
  __shared__ double red[256];
  double local = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) local += partials[i];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double kLog2Pi = 1.8378770664093454835606594728112;
    double lpz = -0.5 * red[0] - 0.5 * (double)d * kLog2Pi * n_nodes;
    double l = ldj ? ldj[0] : 0.0;
    out[0] = lpz;
    out[1] = l;
    out[2] = lpz + l;
    out[3] = n_nodes;
  }
}

constexpr int kLogProbBlocks = 1184;  // 148 SMs x 8

// ---- workspace carve-up --------------------------------------------------------------------
struct Workspace {
  float *x0, *x1, *hbuf, *act0, *act1, *sbuf, *tbuf;
  float *xq, *qbuf, *kbuf, *vbuf, *att, *proj;     // attention block only
  int32_t* attn_fallback;                          // per 32-receiver group: staged attention kernel handed it back
  double* partials;
  unsigned int* counter;       // arrival counter of the fused kernel's log-det hand-off (zeroed per entry point)
  int n_partials_cap;
  size_t bytes;
};

Workspace carve(const Flow& f, int64_t n, int math, void* base) {
  Workspace w{};
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? (void*)(p + off) : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.x0 = (float*)take(nn * f.HP * 4);
  w.x1 = (float*)take(nn * f.HP * 4);
  w.n_partials_cap = (int)ceil_div((int64_t)nn * f.H, 256);
  if (w.n_partials_cap < 1024) w.n_partials_cap = 1024;
  w.partials = (double*)take((size_t)w.n_partials_cap * 8);
  w.counter = (unsigned int*)take(256);
  const bool inject = math != GNF_MATH_FP32 && f.tc_inject;
  const bool layered = math != GNF_MATH_FP32 && f.tc_layered;       // the fp32 path's buffers, MLP layers in k_gemm_tc
  if (math == GNF_MATH_FP32 || inject || layered) {
    const int lp = pad_to(f.L, 8);
    w.hbuf = (float*)take(nn * f.in_pad * 4);
    w.act0 = (float*)take(nn * lp * 4);        // inject: layer-0 pre-activations of the s MLP
    w.act1 = (float*)take(nn * lp * 4);        //         ... and of the t MLP
    if (!inject) {
      w.sbuf = (float*)take(nn * f.HP * 4);
      w.tbuf = (float*)take(nn * f.HP * 4);
    }
    if (f.attn) {
      w.xq = (float*)take(nn * f.hp8 * 4);
      w.qbuf = (float*)take(nn * f.qk_pad * 4);
      w.kbuf = (float*)take(nn * f.qk_pad * 4);
      w.vbuf = (float*)take(nn * f.v_pad * 4);
      w.att = (float*)take(nn * f.hv_pad * 4);
      w.proj = (float*)take(nn * f.cho_pad * 4);
      w.attn_fallback = (int32_t*)take((nn / 32 + 2) * 4);
    }
  }
  w.bytes = off;
  return w;
}

int run_linear(const float* A, const float* W, const float* b, float* C, int64_t M, int N, int K,
               int act, cudaStream_t stream);
}  // namespace
static int linear_any(const Flow& f, int mlp, int math, int slot, const float* A, int lda, int k, const float* w32,
                      const float* bias, float* C, int ldc, int n, int64_t M, cudaStream_t stream);
namespace {
int run_linear(const float* A, const float* W, const float* b, float* C, int64_t M, int N, int K,
               int act, cudaStream_t stream) {
  if (N <= 16) {
    dim3 grid((unsigned)ceil_div(M, BM), 1);
    k_linear<16><<<grid, 256, 0, stream>>>(A, W, b, C, M, N, K, act);
  } else {
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, 128));
    k_linear<128><<<grid, 256, 0, stream>>>(A, W, b, C, M, N, K, act);
  }
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// fp32 layered MLP: hbuf -> out (uses act0/act1 ping-pong)
int run_mlp32(const Flow& f, int mlp, const Workspace& w, float* out, int64_t n, cudaStream_t stream,
              int math = GNF_MATH_FP32) {
  const float* base = f.w32 + (int64_t)mlp * f.w32_per_mlp;
  const float* in = w.hbuf;
  const bool tc = math != GNF_MATH_FP32 && f.tc_layered;      // layer by layer on the tensor cores (gemm_tc.cu)
  for (int l = 0; l < f.K; ++l) {
    const bool last = (l == f.K - 1);
    float* dst = last ? out : ((l & 1) ? w.act1 : w.act0);
    int rc;
    if (tc) {
      const size_t off = (size_t)mlp * f.wgemm_per_mlp + f.gemm_off[l];
      rc = tc_gemm(f, math, in, f.in_pads[l], f.in_pads[l], f.wgemm[0] + off, f.wgemm[1] + off, f.ins[l], f.outs[l],
                   base + f.b32_layer_off[l], last ? 2 : f.d.act, dst, f.out_pads[l], f.out_pads[l], n, stream);
    } else
    rc = run_linear(in, base + f.w32_layer_off[l], base + f.b32_layer_off[l], dst, n,
                        f.out_pads[l], f.in_pads[l], last ? 2 : f.d.act, stream);
    if (rc) return rc;
    in = dst;
  }
  return GNF_OK;
}

// MLP input of GNN `mlp` from the half xa: aggregation blocks (shared by s and t) or attention
int build_attn_input(const Flow& f, int mlp, const float* xa, int64_t n, const int32_t* rowptr,
                     const int32_t* csr_senders, const Workspace& w, cudaStream_t stream, int math = GNF_MATH_FP32) {
  AttnBufs b{w.xq, w.qbuf, w.kbuf, w.vbuf, w.att, w.proj, nullptr, w.attn_fallback};
  return fwd_attn_input(f, mlp, xa, n, rowptr, csr_senders, b, w.hbuf, stream, math);
}

int gnn_forward32(const Flow& f, int mlp, bool build_agg, const float* xa, int64_t n, const int32_t* rowptr,
                  const int32_t* csr_senders, const Workspace& w, float* out, cudaStream_t stream,
                  int math = GNF_MATH_FP32) {
  int rc = GNF_OK;
  if (f.attn) {
    rc = build_attn_input(f, mlp, xa, n, rowptr, csr_senders, w, stream, math);
  } else if (build_agg) {
    k_agg_input<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(
        xa, f.H, f.HP, rowptr, csr_senders, n, f.d.agg == GNF_AGG_MEAN, f.d.block == GNF_BLOCK_CONCAT, f.d.eps,
        f.in_pad, w.hbuf);
    GNF_LAUNCH_CHECK();
  }
  if (rc) return rc;
  rc = run_mlp32(f, mlp, w, out, n, stream, math);
  if (rc) return rc;
  if (f.attn && (f.attn_flags & GNF_ATTN_RESIDUAL)) {                                          // gnn.py:551-552
    k_add_rows<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(out, xa, f.H, f.HP, n);
    GNF_LAUNCH_CHECK();
  }
  if (f.attn && (f.attn_flags & GNF_ATTN_LAYER_NORM)) return fwd_layer_norm(f, mlp, out, n, stream);   // gnn.py:554-556
  return GNF_OK;
}

// one half coupling step: (xa, xb) with the s/t GNNs of (half, step)
int coupling_half(const Flow& f, int half, int step, int inverse, const float* xa, float* xb,
                  int64_t n, const int32_t* rowptr, const int32_t* csr_senders, double* ldj_accum,
                  int math, const Workspace& w, cudaStream_t stream) {
  const int ms = f.mlp_index(0, half, step), mt = f.mlp_index(1, half, step);
  if (n == 0) return GNF_OK;
  if (math == GNF_MATH_FP32 || f.tc_layered) {
    // fp32 kernels layer by layer; tc_layered flows under a tensor-core `math`: same skeleton, the Linears in k_gemm_tc
    // (MLP layers) and k_linear_tc (attention projections)
    const int64_t total = n * f.H;
    const unsigned blocks = (unsigned)ceil_div(total, 256);
    int rc = gnn_forward32(f, ms, true, xa, n, rowptr, csr_senders, w, w.sbuf, stream, math);
    if (rc) return rc;
    rc = gnn_forward32(f, mt, false, xa, n, rowptr, csr_senders, w, w.tbuf, stream, math);   // aggregation input is shared
    if (rc) return rc;
    const bool want_ldj = (!inverse && ldj_accum);
    k_coupling_update<<<blocks, 256, 0, stream>>>(xb, w.sbuf, w.tbuf, n, f.H, f.HP, f.HP, inverse,
                                                  want_ldj ? w.partials : nullptr);
    GNF_LAUNCH_CHECK();
    if (want_ldj) {
      k_reduce_partials<<<1, 256, 0, stream>>>(w.partials, (int)blocks, ldj_accum, 1);
      GNF_LAUNCH_CHECK();
    }
    return GNF_OK;
  }
  if (f.tc_inject) {
    // MLP input wider than the fused kernel's layer-0 tile: input assembly (aggregation or attention) and layer 0
    // (pre-activation, bias included, no activation) in the fp32 kernels, layers 1..K-1 + coupling update fused
    for (int which = 0; which < 2; ++which) {
      const int mlp = which ? mt : ms;
      int rc = GNF_OK;
      if (f.attn) rc = build_attn_input(f, mlp, xa, n, rowptr, csr_senders, w, stream, math);
      else if (which == 0) rc = fwd_agg_input(f, xa, n, rowptr, csr_senders, w.hbuf, stream);    // shared by s and t
      if (rc) return rc;
      const float* base = f.w32 + (int64_t)mlp * f.w32_per_mlp;
      rc = linear_any(f, mlp, math, 4, w.hbuf, f.in_pads[0], f.in_dim, base + f.w32_layer_off[0],
                      base + f.b32_layer_off[0], which ? w.act1 : w.act0, f.out_pads[0], f.L, n, stream);
      if (rc) return rc;
    }
    return tc_coupling_inject(f, ms, mt, math, inverse, w.act0, w.act1, xb, n, w.partials, ldj_accum, w.counter, stream);
  }
  // the fused kernel's last CTA adds the log-det partials into ldj_accum itself (no reduce launch)
  return tc_coupling_half(f, ms, mt, math, inverse, xa, xb, n, rowptr, csr_senders, w.partials, ldj_accum, w.counter,
                          stream);
}

int check_math(const Flow& f, int math, const char* who) {
  GNF_REQUIRE(math >= GNF_MATH_FP32 && math <= GNF_MATH_TC2X, GNF_EINVAL, "%s: bad math %d", who, math);
  if (math != GNF_MATH_FP32)
    GNF_REQUIRE(f.tc_ok || f.tc_inject || f.tc_layered, GNF_EUNSUPPORTED,
                "%s: no tensor-core path for this flow (its layer images would exceed the memory budget of the layered "
                "kernels; num_layers <= %d, got L=%d in=%d H=%d K=%d); use GNF_MATH_FP32",
                who, kMaxLayers, f.L, f.in_dim, f.H, f.K);
  return GNF_OK;
}

}  // namespace

int fwd_linear(const float* A, const float* W, const float* b, float* C, int64_t M, int N, int K, int act,
               cudaStream_t stream) {
  return run_linear(A, W, b, C, M, N, K, act, stream);
}

// One bias-free projection / layer: on the tensor cores (linear_tc.cu) when `math` asks for them and the matrix has an
// image (slot of Flow::lin_off), else the fp32 FFMA kernel.
static int linear_any(const Flow& f, int mlp, int math, int slot, const float* A, int lda, int k, const float* w32,
                      const float* bias, float* C, int ldc, int n, int64_t M, cudaStream_t stream) {
  if (math != GNF_MATH_FP32 && f.lin_off[slot] >= 0) {
    const size_t off = (size_t)mlp * f.wlin_per_mlp + f.lin_off[slot];
    return tc_linear(f, math, A, lda, lda, f.wlin[0] + off, f.wlin[1] + off, k, n, bias, 2, C, ldc, ldc, M, stream);
  }
  return run_linear(A, w32, bias, C, M, ldc, lda, 2, stream);
}

int fwd_attn_input(const Flow& f, int mlp, const float* xa, int64_t n, const int32_t* rowptr,
                   const int32_t* csr_senders, const AttnBufs& w, float* hbuf, cudaStream_t stream, int math) {
  const float* wa = f.wattn + (int64_t)mlp * f.wattn_per_mlp;
  k_pad_rows<<<(unsigned)ceil_div(n * f.hp8, 256), 256, 0, stream>>>(xa, n, f.HP < f.hp8 ? f.HP : f.hp8, f.hp8, w.xq);
  GNF_LAUNCH_CHECK();
  // xa rows are [HP] wide with zero padding, so reading min(HP, hp8) columns and zero-filling is exact
  const int qk = f.heads * f.kq;
  int rc = linear_any(f, mlp, math, 0, w.xq, f.hp8, f.H, wa + f.wq_off, f.zeros, w.qbuf, f.qk_pad, qk, n, stream);   // project_q  gnn.py:509-512
  if (rc) return rc;
  rc = linear_any(f, mlp, math, 1, w.xq, f.hp8, f.H, wa + f.wk_off, f.zeros, w.kbuf, f.qk_pad, qk, n, stream);       // project_k  gnn.py:513-516
  if (rc) return rc;
  rc = linear_any(f, mlp, math, 2, w.xq, f.hp8, f.H, wa + f.wv_off, f.zeros, w.vbuf, f.v_pad, f.vd, n, stream);      // project_v  gnn.py:525-528
  if (rc) return rc;
  const float inv_scale = (f.attn_flags & GNF_ATTN_KQ_DIV) ? 1.f / sqrtf((float)f.kq) : 1.f;
  // the output projection reads hv_pad columns per row (zero weight rows under the pad): the pad columns, which no
  // attention kernel writes, must hold zeros, not whatever the workspace held before (0 * NaN = NaN)
  if (f.hv_pad != f.heads * f.vd) GNF_CUDA(cudaMemsetAsync(w.att, 0, (size_t)n * f.hv_pad * 4, stream));
  // the specialised instantiation (run_grevnet.py defaults); GNF_ATTN_SPECIAL=0 keeps the generic one (A/B knob)
  const char* spec = getenv("GNF_ATTN_SPECIAL");
  const bool vec = f.heads == 8 && f.kq == 10 && !(spec && spec[0] == '0');
  const size_t attn_smem = attn_block_bytes(f.heads, f.kq, f.vd, vec);
  const bool staged = w.fallback && f.vd <= 10 && attn_smem <= 200 * 1024;
  const int32_t* only = nullptr;
  if (staged) {
    static bool attn_configured[kMaxDevices] = {};
    auto kern = k_dm_attn_block<0, 0, 3>;
    if (vec) {
      // 3 CTAs per SM at 80 registers with a few spilled values, or 2 at 96 without: GNF_ATTN_MINB picks (A/B knob)
      const char* mb = getenv("GNF_ATTN_MINB");
      kern = (mb && mb[0] == '2') ? k_dm_attn_block<10, 8, 2> : k_dm_attn_block<10, 8, 3>;
    }
    if (first_use_on_device(attn_configured)) {
      GNF_CUDA(cudaFuncSetAttribute(k_dm_attn_block<0, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GNF_CUDA(cudaFuncSetAttribute(k_dm_attn_block<10, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      GNF_CUDA(cudaFuncSetAttribute(k_dm_attn_block<10, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    const unsigned nblk = (unsigned)ceil_div(n, kAttnRecv);
    GNF_CUDA(cudaMemsetAsync(w.fallback, 0, (size_t)nblk * 4, stream));
    kern<<<nblk, kAttnWarps * 32, attn_smem, stream>>>(w.qbuf, w.kbuf, w.vbuf, f.qk_pad, f.v_pad, f.hv_pad, f.heads, f.kq,
                                                       f.vd, inv_scale, rowptr, csr_senders, n, w.att, w.stats, w.fallback);
    GNF_LAUNCH_CHECK();
    only = w.fallback;                  // receiver groups whose senders are not one compact row range
  }
  if (!staged && (f.vd >= 16 || f.kq >= 32) && f.vd <= 64 && f.kq <= 64)
    k_dm_attn_warp<<<(unsigned)ceil_div(n * f.heads, 8), 256, 0, stream>>>(w.qbuf, w.kbuf, w.vbuf, f.qk_pad, f.v_pad,
                                                                          f.hv_pad, f.heads, f.kq, f.vd, inv_scale, rowptr,
                                                                          csr_senders, n, w.att, w.stats, only, 5);
  else if (f.vd <= 32)
    k_dm_attn<32><<<(unsigned)ceil_div(n * f.heads, 128), 128, 0, stream>>>(w.qbuf, w.kbuf, w.vbuf, f.qk_pad, f.v_pad,
                                                                            f.hv_pad, f.heads, f.kq, f.vd, inv_scale,
                                                                            rowptr, csr_senders, n, w.att, w.stats, only, 5);
  else
    k_dm_attn<64><<<(unsigned)ceil_div(n * f.heads, 128), 128, 0, stream>>>(w.qbuf, w.kbuf, w.vbuf, f.qk_pad, f.v_pad,
                                                                            f.hv_pad, f.heads, f.kq, f.vd, inv_scale,
                                                                            rowptr, csr_senders, n, w.att, w.stats, only, 5);
  GNF_LAUNCH_CHECK();
  rc = linear_any(f, mlp, math, 3, w.att, f.hv_pad, f.heads * f.vd, wa + f.wo_off, f.zeros, w.proj, f.cho_pad, f.cho, n,
                  stream);                                                                  // new_node_proj gnn.py:543-545
  if (rc) return rc;
  const int concat = (f.attn_flags & GNF_ATTN_CONCAT) ? 1 : 0;
  const int width = concat ? f.H + f.cho : f.cho;
  k_attn_input<<<(unsigned)ceil_div(n * width, 256), 256, 0, stream>>>(xa, f.H, f.HP, w.proj, f.cho, f.cho_pad, concat,
                                                                       f.in_pad, n, hbuf);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

int fwd_layer_norm(const Flow& f, int mlp, float* x, int64_t n, cudaStream_t stream) {
  k_layer_norm<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(x, n, f.H, f.HP, f.wln + (int64_t)mlp * 2 * f.HP);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

int fwd_agg_input(const Flow& f, const float* xa, int64_t n, const int32_t* rowptr, const int32_t* csr_senders,
                  float* hbuf, cudaStream_t stream) {
  k_agg_input<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(
      xa, f.H, f.HP, rowptr, csr_senders, n, f.d.agg == GNF_AGG_MEAN, f.d.block == GNF_BLOCK_CONCAT, f.d.eps,
      f.in_pad, hbuf);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace gnf

using namespace gnf;

extern "C" int gnf_abi_version(void) { return GNF_ABI_VERSION; }
extern "C" const char* gnf_last_error(void) { return g_err; }
extern "C" int64_t gnf_launch_count(int reset) {
  int64_t c = launch_counter();
  if (reset) launch_counter() = 0;
  return c;
}
extern "C" int gnf_debug_set_trace(void* device_buf) {
  tc_set_trace(device_buf);
  return GNF_OK;
}
extern "C" int gnf_debug_kernel_timing(int32_t enable) { return tc_kernel_timing(enable); }
extern "C" int gnf_debug_kernel_time(double* total_ms, int64_t* launches) { return tc_kernel_time(total_ms, launches); }
extern "C" int32_t gnf_padded_half(int32_t h) { return pad_to(h, 4); }

static int validate_desc(const gnf_flow_desc* d) {
  GNF_REQUIRE(d, GNF_EINVAL, "null flow desc");
  GNF_REQUIRE(d->num_timesteps >= 1, GNF_EINVAL, "num_timesteps must be >= 1");
  GNF_REQUIRE(d->node_embedding_dim >= 2 && d->node_embedding_dim % 2 == 0, GNF_EINVAL,
              "node_embedding_dim must be even (tf.split at gnn.py:306), got %d", d->node_embedding_dim);
  GNF_REQUIRE(d->latent_dim >= 1, GNF_EINVAL, "latent_dim must be >= 1");
  GNF_REQUIRE(d->num_layers >= 2 && d->num_layers <= kMaxLayers, GNF_EINVAL,
              "num_layers must be in [2, %d]", kMaxLayers);
  GNF_REQUIRE(d->agg == GNF_AGG_SUM || d->agg == GNF_AGG_MEAN, GNF_EINVAL, "bad agg");
  GNF_REQUIRE(d->block == GNF_BLOCK_CONCAT || d->block == GNF_BLOCK_AGG_THEN || d->block == GNF_BLOCK_DM_ATTN,
              GNF_EINVAL, "bad block");
  if (d->block == GNF_BLOCK_DM_ATTN)
    GNF_REQUIRE(d->attn_num_heads >= 1 && d->attn_num_heads <= 64 && d->attn_kq_dim >= 1 && d->attn_kq_dim <= 64 &&
                    d->attn_v_dim >= 1 && d->attn_v_dim <= 64 && d->attn_out_dim >= 1 && d->attn_out_dim <= 4096,
                GNF_EINVAL, "dm_self_attn: need 1 <= heads <= 64, 1 <= kq_dim <= 64, 1 <= v_dim <= 64, 1 <= out_dim <= 4096");
  GNF_REQUIRE(d->node_embedding_dim <= 8192 && d->latent_dim <= 8192, GNF_EINVAL,
              "node_embedding_dim and latent_dim must be <= 8192");
  GNF_REQUIRE(d->act == GNF_ACT_LEAKY_RELU || d->act == GNF_ACT_RELU, GNF_EINVAL, "bad act");
  return GNF_OK;
}

static int64_t params_per_mlp(const gnf_flow_desc* d) {
  const int64_t H = d->node_embedding_dim / 2;
  int64_t in = d->block == GNF_BLOCK_CONCAT ? 2 * H : H;
  int64_t attn = 0;
  if (d->block == GNF_BLOCK_DM_ATTN) {
    const int64_t qk = (int64_t)d->attn_num_heads * d->attn_kq_dim, hv = (int64_t)d->attn_num_heads * d->attn_v_dim;
    attn = 2 * H * qk + H * d->attn_v_dim + hv * d->attn_out_dim;
    in = (d->attn_flags & GNF_ATTN_CONCAT) ? H + d->attn_out_dim : d->attn_out_dim;
  }
  const int64_t L = d->latent_dim, K = d->num_layers;
  const int64_t ln = (d->block == GNF_BLOCK_DM_ATTN && (d->attn_flags & GNF_ATTN_LAYER_NORM)) ? 2 * H : 0;
  return attn + in * L + L + (K - 2) * (L * L + L) + L * H + H + ln;
}

extern "C" int64_t gnf_flow_param_count(const gnf_flow_desc* d) {
  if (validate_desc(d)) return -1;
  const int64_t n_mlps = 4ll * (d->weight_sharing ? 1 : d->num_timesteps);
  return n_mlps * params_per_mlp(d);
}

extern "C" int gnf_flow_create(gnf_flow** out, const gnf_flow_desc* d) {
  GNF_REQUIRE(out, GNF_EINVAL, "gnf_flow_create: null out");
  int rc = validate_desc(d);
  if (rc) return rc;
  gnf_flow* h = new gnf_flow();
  Flow& f = h->f;
  f.d = *d;
  f.H = d->node_embedding_dim / 2;
  f.HP = gnf_padded_half(f.H);
  f.in_dim = d->block == GNF_BLOCK_CONCAT ? 2 * f.H : f.H;
  if (d->block == GNF_BLOCK_DM_ATTN) {
    f.attn = 1;
    f.heads = d->attn_num_heads; f.kq = d->attn_kq_dim; f.vd = d->attn_v_dim; f.cho = d->attn_out_dim;
    f.attn_flags = d->attn_flags;
    f.hp8 = pad_to(f.H, 8); f.qk_pad = pad_to(f.heads * f.kq, 8); f.v_pad = pad_to(f.vd, 8);
    f.hv_pad = pad_to(f.heads * f.vd, 8); f.cho_pad = pad_to(f.cho, 8);
    f.in_dim = (f.attn_flags & GNF_ATTN_CONCAT) ? f.H + f.cho : f.cho;
    f.wq_off = 0;
    f.wk_off = f.wq_off + (int64_t)f.hp8 * f.qk_pad;
    f.wv_off = f.wk_off + (int64_t)f.hp8 * f.qk_pad;
    f.wo_off = f.wv_off + (int64_t)f.hp8 * f.v_pad;
    f.wattn_per_mlp = f.wo_off + (int64_t)f.hv_pad * f.cho_pad;
    f.wqT_off = 0;
    f.wkT_off = f.wqT_off + (int64_t)f.qk_pad * f.hp8;
    f.wvT_off = f.wkT_off + (int64_t)f.qk_pad * f.hp8;
    f.woT_off = f.wvT_off + (int64_t)f.v_pad * f.hp8;
    f.wattnT_per_mlp = f.woT_off + (int64_t)f.cho_pad * f.hv_pad;
    f.mlp_off = 2ll * f.H * f.heads * f.kq + (int64_t)f.H * f.vd + (int64_t)f.heads * f.vd * f.cho;
  }
  f.in_pad = pad_to(f.in_dim, 8);
  if (f.attn && (f.attn_flags & GNF_ATTN_LAYER_NORM)) f.ln_off = params_per_mlp(d) - 2 * f.H;
  f.L = d->latent_dim;
  f.K = d->num_layers;
  f.n_mlps = 4 * (d->weight_sharing ? 1 : d->num_timesteps);
  f.params_per_mlp = params_per_mlp(d);
  const int lp = pad_to(f.L, 8);
  int64_t off = 0;
  for (int l = 0; l < f.K; ++l) {
    f.ins[l] = l == 0 ? f.in_dim : f.L;
    f.outs[l] = l == f.K - 1 ? f.H : f.L;
    f.in_pads[l] = l == 0 ? f.in_pad : lp;
    f.out_pads[l] = l == f.K - 1 ? f.HP : lp;
    f.w32_layer_off[l] = off;
    off += (int64_t)f.in_pads[l] * f.out_pads[l];
    f.b32_layer_off[l] = off;
    off += f.out_pads[l];
    off = (off + 3) / 4 * 4;  // keep float4 alignment of the next W
  }
  f.w32_per_mlp = off;
  int64_t offT = 0, offF = 0;
  for (int l = 0; l < f.K; ++l) {
    f.out_pad8[l] = pad_to(f.out_pads[l], 8);
    f.w32T_layer_off[l] = offT;
    offT += (int64_t)f.out_pad8[l] * f.in_pads[l];
    f.flat_w_off[l] = offF;
    offF += (int64_t)f.ins[l] * f.outs[l];
    f.flat_b_off[l] = offF;
    offF += f.outs[l];
  }
  f.w32T_per_mlp = offT;
  for (int l = 0; l < f.K; ++l) {   // flat offsets are relative to the GNN's parameter block
    f.flat_w_off[l] += f.mlp_off;
    f.flat_b_off[l] += f.mlp_off;
  }
  cudaError_t e = cudaMalloc(&f.w32, (size_t)f.n_mlps * f.w32_per_mlp * 4);
  if (e == cudaSuccess) e = cudaMalloc(&f.w32T, (size_t)f.n_mlps * f.w32T_per_mlp * 4);
  size_t n_zeros = 8192;                  // zero bias of the bias-free projections: as wide as the widest of them
  for (int v : {f.qk_pad, f.v_pad, f.hv_pad, f.cho_pad}) n_zeros = (size_t)v > n_zeros ? (size_t)v : n_zeros;
  if (e == cudaSuccess) e = cudaMalloc(&f.zeros, n_zeros * 4);
  if (e == cudaSuccess) e = cudaMemset(f.zeros, 0, n_zeros * 4);
  if (e == cudaSuccess) e = cudaMalloc(&f.range_flag, 256);
  if (e == cudaSuccess) e = cudaMemset(f.range_flag, 0, 256);
  if (e == cudaSuccess && f.attn) e = cudaMalloc(&f.wattn, (size_t)f.n_mlps * f.wattn_per_mlp * 4);
  if (e == cudaSuccess && f.attn) e = cudaMalloc(&f.wattnT, (size_t)f.n_mlps * f.wattnT_per_mlp * 4);
  if (e == cudaSuccess && f.attn) e = cudaMalloc(&f.wln, (size_t)f.n_mlps * 2 * f.HP * 4);
  if (e != cudaSuccess) {
    delete h;
    set_error("gnf_flow_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    return GNF_ECUDA;
  }
  f.tc_ok = tc_shape_supported(f);
  f.tc_inject = !f.tc_ok && tc_inject_supported(f);
  if (f.tc_ok || f.tc_inject) {
    f.wtc_per_mlp = (int64_t)tc_bytes_per_mlp(f.L, f.K);
    cudaError_t e0 = cudaMalloc(&f.wtc[0], (size_t)f.n_mlps * f.wtc_per_mlp);
    cudaError_t e1 = cudaMalloc(&f.wtc[1], (size_t)f.n_mlps * f.wtc_per_mlp);
    cudaError_t e2 = cudaMalloc(&f.btc, (size_t)f.n_mlps * f.K * 256 * 4);
    // backward chains (inject flows too: their layer-0 / last-position images are packed but never read)
    if (e2 == cudaSuccess) e2 = cudaMalloc(&f.wtcT, (size_t)f.n_mlps * f.wtc_per_mlp);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&f.wtcB[0], (size_t)f.n_mlps * f.wtc_per_mlp);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&f.wtcB[1], (size_t)f.n_mlps * f.wtc_per_mlp);
    if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess) {
      gnf_flow_destroy(h);
      set_error("gnf_flow_create: cudaMalloc (tc weights) failed");
      return GNF_ECUDA;
    }
  }
  if (!f.tc_ok && !f.tc_inject) {
    // every other shape: the MLP layers one by one in k_gemm_tc (images of all K layers, fp16 and bf16 element type);
    // flows whose images would not fit a sane share of HBM stay on the fp32 kernels
    int64_t goff = 0;
    for (int l = 0; l < f.K; ++l) {
      f.gemm_off[l] = goff;
      goff += (int64_t)align_up(tc_gemm_image_bytes(f.ins[l], f.outs[l]), 256);
    }
    f.wgemm_per_mlp = goff;
    int64_t goffT = 0;
    for (int l = 0; l < f.K; ++l) {
      f.gemmT_off[l] = goffT;
      goffT += (int64_t)align_up(tc_gemm_image_bytes(f.outs[l], f.ins[l]), 256);
    }
    f.wgemmT_per_mlp = goffT;
    if (((double)goff * 2 + (double)goffT) * f.n_mlps <= 24e9) {
      cudaError_t e0 = cudaMalloc(&f.wgemm[0], (size_t)f.n_mlps * goff);
      cudaError_t e1 = cudaMalloc(&f.wgemm[1], (size_t)f.n_mlps * goff);
      if (e1 == cudaSuccess) e1 = cudaMalloc(&f.wgemmT, (size_t)f.n_mlps * goffT);
      if (e0 != cudaSuccess || e1 != cudaSuccess) {
        gnf_flow_destroy(h);
        set_error("gnf_flow_create: cudaMalloc (gemm images) failed");
        return GNF_ECUDA;
      }
      f.tc_layered = true;
    }
  }
  if (f.tc_inject || f.tc_layered) {
    // stand-alone tensor-core linears: layer 0 (inject flows) and, for dm_self_attn, the four projections
    int64_t off = 0;
    auto want = [&](int slot, int k, int n) {
      if (tc_linear_shape_ok(k, n)) {
        f.lin_off[slot] = off;
        off += (int64_t)align_up(tc_linear_image_bytes(k, n), 256);
      }
    };
    if (f.attn) {
      want(0, f.H, f.heads * f.kq);
      want(1, f.H, f.heads * f.kq);
      want(2, f.H, f.vd);
      want(3, f.heads * f.vd, f.cho);
    }
    if (f.tc_inject) {
      want(4, f.in_dim, f.L);
      want(5, f.L, f.in_dim);                        // backward: g_h = delta_0 W0^T
      if (f.attn) want(6, f.cho, f.heads * f.vd);    // backward: g_att = g_proj Wo^T
    }
    f.wlin_per_mlp = off;
    if (off > 0) {
      cudaError_t e0 = cudaMalloc(&f.wlin[0], (size_t)f.n_mlps * off);
      cudaError_t e1 = cudaMalloc(&f.wlin[1], (size_t)f.n_mlps * off);
      if (e0 != cudaSuccess || e1 != cudaSuccess) {
        gnf_flow_destroy(h);
        set_error("gnf_flow_create: cudaMalloc (linear images) failed");
        return GNF_ECUDA;
      }
    }
  }
  if (pack_build_jobs(f) != GNF_OK || tc_build_half_tables(f) != GNF_OK) {
    gnf_flow_destroy(h);
    return GNF_ECUDA;
  }
  *out = h;
  return GNF_OK;
}

extern "C" int gnf_flow_range_flag(const gnf_flow* h, int32_t* host_flag, int32_t reset, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(h, GNF_EINVAL, "gnf_flow_range_flag: null flow");
  if (host_flag) GNF_CUDA(cudaMemcpyAsync(host_flag, h->f.range_flag, 4, cudaMemcpyDeviceToHost, stream));
  if (reset) GNF_CUDA(cudaMemsetAsync(h->f.range_flag, 0, 4, stream));
  return GNF_OK;
}

extern "C" int gnf_flow_destroy(gnf_flow* h) {
  if (!h) return GNF_OK;
  cudaFree(h->f.w32);
  cudaFree(h->f.w32T);
  cudaFree(h->f.zeros);
  cudaFree(h->f.wattn);
  cudaFree(h->f.wattnT);
  cudaFree(h->f.wln);
  cudaFree(h->f.wtc[0]);
  cudaFree(h->f.wtc[1]);
  cudaFree(h->f.btc);
  cudaFree(h->f.wtcT);
  cudaFree(h->f.wtcB[0]);
  cudaFree(h->f.wtcB[1]);
  cudaFree(h->f.pack_jobs);
  cudaFree(h->f.wlin[0]);
  cudaFree(h->f.wlin[1]);
  cudaFree(h->f.wgemm[0]);
  cudaFree(h->f.wgemm[1]);
  cudaFree(h->f.wgemmT);
  cudaFree(h->f.half_tables);
  cudaFree(h->f.range_flag);
  delete h;
  return GNF_OK;
}

extern "C" int gnf_flow_supports(const gnf_flow* h, int32_t math) {
  if (!h) return 0;
  if (math == GNF_MATH_FP32) return 1;
  return (math >= GNF_MATH_TC3X && math <= GNF_MATH_TC2X && (h->f.tc_ok || h->f.tc_inject || h->f.tc_layered)) ? 1 : 0;
}

extern "C" int gnf_flow_supports_backward(const gnf_flow* h, int32_t math) {
  if (!h) return 0;
  if (math == GNF_MATH_FP32) return 1;
  // layered flows: forward recompute on the tensor cores, dX / dW on the fp32 kernels (backward.cu::bwd_half_fp32)
  return (math >= GNF_MATH_TC3X && math <= GNF_MATH_TC2X &&
          (tc_bwd_supported(h->f) || tc_bwd_inject_supported(h->f) || h->f.tc_layered)) ? 1 : 0;
}

extern "C" int gnf_flow_set_params(gnf_flow* h, const float* params, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(h && params, GNF_EINVAL, "gnf_flow_set_params: null pointer");
  return pack_all(h->f, params, stream);
}

extern "C" size_t gnf_grevnet_workspace(const gnf_flow* h, int64_t n_nodes, int32_t math) {
  if (!h || n_nodes < 0) return 0;
  return carve(h->f, n_nodes, math, nullptr).bytes;
}

static int check_common(const gnf_flow* h, int64_t n, int64_t e, const int32_t* rowptr,
                        const int32_t* csr, void* ws, size_t ws_bytes, int math, const char* who) {
  GNF_REQUIRE(h, GNF_EINVAL, "%s: null flow", who);
  GNF_REQUIRE(n >= 0 && e >= 0, GNF_EINVAL, "%s: negative size", who);
  int rc = check_math(h->f, math, who);
  if (rc) return rc;
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(rowptr && (e == 0 || csr), GNF_EINVAL, "%s: null CSR pointer", who);
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0, GNF_EINVAL, "%s: workspace must be 256-byte aligned", who);
  GNF_REQUIRE(ws_bytes >= carve(h->f, n, math, nullptr).bytes, GNF_EWORKSPACE, "%s: workspace too small", who);
  return GNF_OK;
}

extern "C" int gnf_grevnet_forward(const gnf_flow* h, const float* x, int64_t n, int64_t e,
                                   const int32_t* rowptr, const int32_t* csr, float* z, double* ldj,
                                   int32_t math, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_grevnet_forward");
  if (rc) return rc;
  GNF_REQUIRE(ldj, GNF_EINVAL, "gnf_grevnet_forward: null ldj");
  GNF_CUDA(cudaMemsetAsync(ldj, 0, 8, stream));
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && z, GNF_EINVAL, "gnf_grevnet_forward: null x/z");
  const Flow& f = h->f;
  Workspace w = carve(f, n, math, ws);
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 256, stream));
  const int D = f.d.node_embedding_dim;
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));   // pad columns must be 0, not NaN
  k_split<<<(unsigned)ceil_div(n * f.HP, 256), 256, 0, stream>>>(x, n, D, f.H, f.HP, w.x0, w.x1);
  GNF_LAUNCH_CHECK();
  const bool persistent = math != GNF_MATH_FP32 && f.tc_ok && tc_persistent_wanted(n);
  if (persistent) {   // all 2T half steps in ONE cooperative launch (grid barrier where the launch boundaries were)
    rc = tc_flow_persistent(f, math, 0, w.x0, w.x1, n, rowptr, csr, w.partials, ldj, w.counter, stream);
    if (rc) return rc;
  }
  for (int i = 0; i < f.d.num_timesteps && !persistent; ++i) {   // gnn.py:309
    rc = coupling_half(f, 0, i, 0, w.x0, w.x1, n, rowptr, csr, ldj, math, w, stream);   // gnn.py:320-323
    if (rc) return rc;
    rc = coupling_half(f, 1, i, 0, w.x1, w.x0, n, rowptr, csr, ldj, math, w, stream);   // gnn.py:335-338
    if (rc) return rc;
  }
  k_merge<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, f.H, f.HP, z);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_grevnet_inverse(const gnf_flow* h, const float* z, int64_t n, int64_t e,
                                   const int32_t* rowptr, const int32_t* csr, float* x, int32_t math,
                                   void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_grevnet_inverse");
  if (rc) return rc;
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && z, GNF_EINVAL, "gnf_grevnet_inverse: null x/z");
  const Flow& f = h->f;
  Workspace w = carve(f, n, math, ws);
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 256, stream));
  const int D = f.d.node_embedding_dim;
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));   // pad columns must be 0, not NaN
  k_split<<<(unsigned)ceil_div(n * f.HP, 256), 256, 0, stream>>>(z, n, D, f.H, f.HP, w.x0, w.x1);
  GNF_LAUNCH_CHECK();
  const bool persistent = math != GNF_MATH_FP32 && f.tc_ok && tc_persistent_wanted(n);
  if (persistent) {
    rc = tc_flow_persistent(f, math, 1, w.x0, w.x1, n, rowptr, csr, w.partials, nullptr, w.counter, stream);
    if (rc) return rc;
  }
  for (int i = f.d.num_timesteps - 1; i >= 0 && !persistent; --i) {   // gnn.py:347
    rc = coupling_half(f, 1, i, 1, w.x1, w.x0, n, rowptr, csr, nullptr, math, w, stream);  // gnn.py:353-359
    if (rc) return rc;
    rc = coupling_half(f, 0, i, 1, w.x0, w.x1, n, rowptr, csr, nullptr, math, w, stream);  // gnn.py:366-372
    if (rc) return rc;
  }
  k_merge<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, f.H, f.HP, x);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_coupling_step(const gnf_flow* h, int32_t step, int32_t inverse, float* x0, float* x1,
                                 int64_t n, int64_t e, const int32_t* rowptr, const int32_t* csr,
                                 double* ldj_accum, int32_t math, void* ws, size_t ws_bytes,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_coupling_step");
  if (rc) return rc;
  GNF_REQUIRE(step >= 0 && step < h->f.d.num_timesteps, GNF_EINVAL, "gnf_coupling_step: bad step %d", step);
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x0 && x1, GNF_EINVAL, "gnf_coupling_step: null x0/x1");
  const Flow& f = h->f;
  Workspace w = carve(f, n, math, ws);
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 4, stream));
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));   // pad columns must be 0, not NaN
  if (!inverse) {
    rc = coupling_half(f, 0, step, 0, x0, x1, n, rowptr, csr, ldj_accum, math, w, stream);
    if (rc) return rc;
    return coupling_half(f, 1, step, 0, x1, x0, n, rowptr, csr, ldj_accum, math, w, stream);
  }
  rc = coupling_half(f, 1, step, 1, x1, x0, n, rowptr, csr, nullptr, math, w, stream);
  if (rc) return rc;
  return coupling_half(f, 0, step, 1, x0, x1, n, rowptr, csr, nullptr, math, w, stream);
}

extern "C" int gnf_gnn_forward(const gnf_flow* h, int32_t which, int32_t half, int32_t step, const float* x,
                               int64_t n, int64_t e, const int32_t* rowptr, const int32_t* csr, float* out,
                               void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, GNF_MATH_FP32, "gnf_gnn_forward");
  if (rc) return rc;
  const Flow& f = h->f;
  GNF_REQUIRE((which == 0 || which == 1) && (half == 0 || half == 1) && step >= 0 &&
                  step < f.d.num_timesteps, GNF_EINVAL, "gnf_gnn_forward: bad GNN selector");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && out, GNF_EINVAL, "gnf_gnn_forward: null x/out");
  Workspace w = carve(f, n, GNF_MATH_FP32, ws);
  GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));
  k_pad_rows<<<(unsigned)ceil_div(n * f.HP, 256), 256, 0, stream>>>(x, n, f.H, f.HP, w.x0);
  GNF_LAUNCH_CHECK();
  rc = gnn_forward32(f, f.mlp_index(which, half, step), true, w.x0, n, rowptr, csr, w, w.sbuf, stream);
  if (rc) return rc;
  k_unpad_rows<<<(unsigned)ceil_div(n * f.H, 256), 256, 0, stream>>>(w.sbuf, n, f.H, f.HP, out);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_coupling_half(const gnf_flow* h, int32_t half, int32_t step, int32_t inverse, const float* xa,
                                 float* xb, int64_t n, int64_t e, const int32_t* rowptr, const int32_t* csr,
                                 double* ldj_accum, int32_t math, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_coupling_half");
  if (rc) return rc;
  GNF_REQUIRE((half == 0 || half == 1) && step >= 0 && step < h->f.d.num_timesteps, GNF_EINVAL,
              "gnf_coupling_half: bad (half, step)");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(xa && xb && xa != xb, GNF_EINVAL, "gnf_coupling_half: xa/xb must be distinct non-null buffers");
  const Flow& f = h->f;
  Workspace w = carve(f, n, math, ws);
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 4, stream));
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));   // pad columns must be 0, not NaN
  return coupling_half(f, half, step, inverse, xa, xb, n, rowptr, csr, inverse ? nullptr : ldj_accum, math, w,
                       stream);
}

extern "C" int gnf_split_halves(const float* x, int64_t n, int32_t d, float* x0, float* x1, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(d >= 2 && d % 2 == 0 && n >= 0, GNF_EINVAL, "gnf_split_halves: bad shape");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && x0 && x1, GNF_EINVAL, "gnf_split_halves: null pointer");
  const int hh = d / 2, hp = gnf_padded_half(hh);
  k_split<<<(unsigned)ceil_div(n * hp, 256), 256, 0, stream>>>(x, n, d, hh, hp, x0, x1);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_merge_halves(const float* x0, const float* x1, int64_t n, int32_t d, float* x, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(d >= 2 && d % 2 == 0 && n >= 0, GNF_EINVAL, "gnf_merge_halves: bad shape");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && x0 && x1, GNF_EINVAL, "gnf_merge_halves: null pointer");
  const int hh = d / 2, hp = gnf_padded_half(hh);
  k_merge<<<(unsigned)ceil_div(n * d, 256), 256, 0, stream>>>(x0, x1, n, d, hh, hp, x);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

constexpr int kBnBlocks = 296;
extern "C" size_t gnf_bn_moments_workspace(int32_t h) { return (size_t)kBnBlocks * 2 * gnf_padded_half(h) * 8; }

extern "C" int gnf_bn_moments(const float* x, int64_t n, int32_t hh, double* sums, void* ws, size_t ws_bytes,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int hp = gnf_padded_half(hh);
  GNF_REQUIRE(hh >= 1 && hp <= 256 && n >= 0 && sums, GNF_EINVAL, "gnf_bn_moments: bad argument (H <= 256)");
  GNF_REQUIRE(ws && ws_bytes >= gnf_bn_moments_workspace(hh), GNF_EWORKSPACE, "gnf_bn_moments: workspace too small");
  GNF_REQUIRE(n == 0 || x, GNF_EINVAL, "gnf_bn_moments: null x");
  const int tpb = (256 / hp) * hp;
  const int64_t total = n * hp;
  int blocks = (int)(ceil_div(total, tpb) < kBnBlocks ? ceil_div(total, tpb) : kBnBlocks);
  if (blocks < 1) blocks = 1;
  k_bn_partial<<<blocks, tpb, 2 * tpb * sizeof(double), stream>>>(x, total, hp, (double*)ws);
  GNF_LAUNCH_CHECK();
  k_bn_final<<<1, 256, 0, stream>>>((const double*)ws, blocks, hh, hp, sums, (double)n);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_bn_finalize(const double* sums, int32_t hh, const float* gamma, const float* beta, double eps,
                               double n_local, double* ldj_accum, float* scale_shift, double* stats,
                               float* moving_mean, float* moving_var, float momentum, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(sums && gamma && beta && scale_shift && stats && hh >= 1 && hh <= 256, GNF_EINVAL,
              "gnf_bn_finalize: bad argument (H <= 256)");
  GNF_REQUIRE(momentum < 0.f || (moving_mean && moving_var), GNF_EINVAL, "gnf_bn_finalize: null moving statistics");
  k_bn_finalize<<<1, 256, 0, stream>>>(sums, hh, gamma, beta, eps, n_local, ldj_accum, scale_shift, stats, moving_mean,
                                       moving_var, momentum);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_bn_backward_coef(const double* sums, const double* stats, int32_t hh, const float* gamma,
                                    const float* beta, double eps, double loss_scale, float* coef, float* inv_gamma,
                                    double* g_gamma, double* g_beta, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(gamma && beta && hh >= 1 && hh <= 256, GNF_EINVAL, "gnf_bn_backward_coef: bad argument (H <= 256)");
  GNF_REQUIRE(sums ? (stats && coef && g_gamma && g_beta) : (inv_gamma != nullptr), GNF_EINVAL,
              "gnf_bn_backward_coef: null pointer");
  k_bn_bwd_coef<<<1, 256, 0, stream>>>(sums, stats, hh, gamma, beta, eps, loss_scale, coef, inv_gamma, g_gamma, g_beta);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_bn_backward_sums(const float* y, const float* g, int64_t n, int32_t hh, const float* beta,
                                    const float* inv_gamma, double* sums, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int hp = gnf_padded_half(hh);
  GNF_REQUIRE(hh >= 1 && hp <= 256 && n >= 0 && sums && beta && inv_gamma, GNF_EINVAL,
              "gnf_bn_backward_sums: bad argument (H <= 256)");
  GNF_REQUIRE(ws && ws_bytes >= gnf_bn_moments_workspace(hh), GNF_EWORKSPACE, "gnf_bn_backward_sums: workspace too small");
  GNF_REQUIRE(n == 0 || (y && g), GNF_EINVAL, "gnf_bn_backward_sums: null y/g");
  const int tpb = (256 / hp) * hp;
  const int64_t total = n * hp;
  int blocks = (int)(ceil_div(total, tpb) < kBnBlocks ? ceil_div(total, tpb) : kBnBlocks);
  if (blocks < 1) blocks = 1;
  k_bn_bwd_partial<<<blocks, tpb, 2 * tpb * sizeof(double), stream>>>(y, g, beta, inv_gamma, total, hh, hp, (double*)ws);
  GNF_LAUNCH_CHECK();
  k_bn_final<<<1, 256, 0, stream>>>((const double*)ws, blocks, hh, hp, sums, -1.0);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_bn_backward_apply(float* y, float* g, int64_t n, int32_t hh, const float* coef, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(hh >= 1 && n >= 0 && coef, GNF_EINVAL, "gnf_bn_backward_apply: bad argument");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(y && g, GNF_EINVAL, "gnf_bn_backward_apply: null y/g");
  const int hp = gnf_padded_half(hh);
  k_bn_bwd_apply<<<(unsigned)ceil_div(n * hp, 256), 256, 0, stream>>>(y, g, n, hh, hp, coef);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_affine_rows(float* x, int64_t n, int32_t hh, const float* scale, const float* shift,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(hh >= 1 && n >= 0, GNF_EINVAL, "gnf_affine_rows: bad shape");
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && scale && shift, GNF_EINVAL, "gnf_affine_rows: null pointer");
  const int hp = gnf_padded_half(hh);
  k_affine_rows<<<(unsigned)ceil_div(n * hp, 256), 256, 0, stream>>>(x, n, hh, hp, scale, shift);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// ---- a9, chained: the whole batch-norm flow in one call (single-rank batches; sharded runs all-reduce the moment sums
// between gnf_bn_moments and gnf_bn_finalize and therefore stay on the per-half-step entries) -------------------------
static size_t bn_chain_sums_off(int32_t h) { return align_up(gnf_bn_moments_workspace(h), 256); }
static size_t bn_chain_ss_off(int32_t h) { return bn_chain_sums_off(h) + align_up((size_t)(2 * h + 1) * 8, 256); }
extern "C" size_t gnf_grevnet_bn_workspace(const gnf_flow* h) {
  if (!h) return 0;
  return bn_chain_ss_off(h->f.H) + align_up((size_t)2 * h->f.H * 4, 256);
}

static int check_bn_chain(const gnf_flow* h, const float* gamma, const float* beta, const float* mm, const float* mv,
                          void* bn_ws, size_t bn_ws_bytes, const char* who) {
  GNF_REQUIRE(gamma && beta && mm && mv, GNF_EINVAL, "%s: null batch-norm parameter", who);
  GNF_REQUIRE(gnf_padded_half(h->f.H) <= 256, GNF_EINVAL, "%s: batch norm needs D/2 <= 256", who);
  GNF_REQUIRE(bn_ws && ((uintptr_t)bn_ws % 256) == 0 && bn_ws_bytes >= gnf_grevnet_bn_workspace(h), GNF_EWORKSPACE,
              "%s: batch-norm workspace too small or misaligned", who);
  return GNF_OK;
}

extern "C" int gnf_grevnet_forward_bn(const gnf_flow* h, const float* x, int64_t n, int64_t e, const int32_t* rowptr,
                                      const int32_t* csr, const float* gamma, const float* beta, float* moving_mean,
                                      float* moving_var, double eps, float momentum, float* z, double* ldj,
                                      double* stats, int32_t math, void* ws, size_t ws_bytes, void* bn_ws,
                                      size_t bn_ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_grevnet_forward_bn");
  if (rc) return rc;
  GNF_REQUIRE(ldj && stats, GNF_EINVAL, "gnf_grevnet_forward_bn: null ldj/stats");
  rc = check_bn_chain(h, gamma, beta, moving_mean, moving_var, bn_ws, bn_ws_bytes, "gnf_grevnet_forward_bn");
  if (rc) return rc;
  GNF_CUDA(cudaMemsetAsync(ldj, 0, 8, stream));
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && z, GNF_EINVAL, "gnf_grevnet_forward_bn: null x/z");
  const Flow& f = h->f;
  const int H = f.H, T = f.d.num_timesteps, D = f.d.node_embedding_dim;
  Workspace w = carve(f, n, math, ws);
  double* sums = (double*)((uint8_t*)bn_ws + bn_chain_sums_off(H));
  float* ss = (float*)((uint8_t*)bn_ws + bn_chain_ss_off(H));
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 256, stream));
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));
  k_split<<<(unsigned)ceil_div(n * f.HP, 256), 256, 0, stream>>>(x, n, D, H, f.HP, w.x0, w.x1);
  GNF_LAUNCH_CHECK();
  for (int i = 0; i < T; ++i)                                    // gnn.py:309-338
    for (int half = 0; half < 2; ++half) {
      float* xa = half ? w.x1 : w.x0;
      float* xb = half ? w.x0 : w.x1;
      const int64_t po = ((int64_t)half * T + i) * H;
      // bn.inverse + its log-det on the conditioning half (gnn.py:310-313,325-328), batch statistics
      rc = gnf_bn_moments(xa, n, H, sums, bn_ws, bn_chain_sums_off(H), stream_);
      if (rc) return rc;
      rc = gnf_bn_finalize(sums, H, gamma + po, beta + po, eps, (double)n, ldj, ss,
                           stats + ((int64_t)half * T + i) * (2 * H + 1), moving_mean + po, moving_var + po, momentum,
                           stream_);
      if (rc) return rc;
      rc = gnf_affine_rows(xa, n, H, ss, ss + H, stream_);
      if (rc) return rc;
      GNF_CUDA(cudaMemsetAsync(w.counter, 0, 4, stream));
      rc = coupling_half(f, half, i, 0, xa, xb, n, rowptr, csr, ldj, math, w, stream);
      if (rc) return rc;
    }
  k_merge<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, H, f.HP, z);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" int gnf_grevnet_inverse_bn(const gnf_flow* h, const float* z, int64_t n, int64_t e, const int32_t* rowptr,
                                      const int32_t* csr, const float* gamma, const float* beta,
                                      const float* moving_mean, const float* moving_var, double eps, float* x,
                                      int32_t math, void* ws, size_t ws_bytes, void* bn_ws, size_t bn_ws_bytes,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_common(h, n, e, rowptr, csr, ws, ws_bytes, math, "gnf_grevnet_inverse_bn");
  if (rc) return rc;
  rc = check_bn_chain(h, gamma, beta, moving_mean, moving_var, bn_ws, bn_ws_bytes, "gnf_grevnet_inverse_bn");
  if (rc) return rc;
  if (n == 0) return GNF_OK;
  GNF_REQUIRE(x && z, GNF_EINVAL, "gnf_grevnet_inverse_bn: null x/z");
  const Flow& f = h->f;
  const int H = f.H, T = f.d.num_timesteps, D = f.d.node_embedding_dim;
  Workspace w = carve(f, n, math, ws);
  float* ss = (float*)((uint8_t*)bn_ws + bn_chain_ss_off(H));
  GNF_CUDA(cudaMemsetAsync(w.counter, 0, 256, stream));
  if (w.hbuf) GNF_CUDA(cudaMemsetAsync(w.hbuf, 0, (size_t)n * f.in_pad * 4, stream));
  k_split<<<(unsigned)ceil_div(n * f.HP, 256), 256, 0, stream>>>(z, n, D, H, f.HP, w.x0, w.x1);
  GNF_LAUNCH_CHECK();
  for (int i = T - 1; i >= 0; --i)                               // gnn.py:347-372
    for (int half = 1; half >= 0; --half) {
      float* xa = half ? w.x1 : w.x0;
      float* xb = half ? w.x0 : w.x1;
      const int64_t po = ((int64_t)half * T + i) * H;
      GNF_CUDA(cudaMemsetAsync(w.counter, 0, 4, stream));
      rc = coupling_half(f, half, i, 1, xa, xb, n, rowptr, csr, nullptr, math, w, stream);
      if (rc) return rc;
      // bn.forward on the conditioning half: de-normalise with the moving statistics (gnn.py:356-358,369-371)
      k_bn_moving_scale_shift<<<(unsigned)ceil_div(H, 128), 128, 0, stream>>>(gamma + po, beta + po, moving_mean + po,
                                                                              moving_var + po, (float)eps, H, ss);
      GNF_LAUNCH_CHECK();
      rc = gnf_affine_rows(xa, n, H, ss, ss + H, stream_);
      if (rc) return rc;
    }
  k_merge<<<(unsigned)ceil_div(n * D, 256), 256, 0, stream>>>(w.x0, w.x1, n, D, H, f.HP, x);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

extern "C" size_t gnf_log_prob_workspace(int64_t, int32_t) { return (size_t)kLogProbBlocks * 8; }

namespace gnf {
// |z|^2 partials of the log-prob assembly (fp64, fixed order) into the workspace; shared with peer.cu
int log_prob_partials(const float* z, int64_t n, int32_t d, void* ws, size_t ws_bytes, cudaStream_t stream, int* n_blocks) {
  GNF_REQUIRE(n >= 0 && d > 0, GNF_EINVAL, "gnf_log_prob: bad argument");
  GNF_REQUIRE(ws && ws_bytes >= (size_t)kLogProbBlocks * 8, GNF_EWORKSPACE, "gnf_log_prob: workspace too small");
  GNF_REQUIRE(n == 0 || z, GNF_EINVAL, "gnf_log_prob: null z");
  const int64_t total = n * d;
  int blocks = (int)(ceil_div(total, 256) < kLogProbBlocks ? ceil_div(total, 256) : kLogProbBlocks);
  if (blocks < 1) blocks = 1;
  k_sumsq<<<blocks, 256, 0, stream>>>(z, total, (double*)ws);
  GNF_LAUNCH_CHECK();
  *n_blocks = blocks;
  return GNF_OK;
}
}  // namespace gnf

extern "C" int gnf_log_prob(const float* z, int64_t n, int32_t d, const double* ldj, double* out,
                            void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(out, GNF_EINVAL, "gnf_log_prob: null out");
  int blocks = 0;
  int rc = log_prob_partials(z, n, d, ws, ws_bytes, stream, &blocks);
  if (rc) return rc;
  k_log_prob_final<<<1, 256, 0, stream>>>((const double*)ws, blocks, ldj, (double)n, d, out);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}
