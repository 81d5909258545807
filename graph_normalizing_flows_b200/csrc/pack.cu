// Weight re-pack in ONE launch (gnf_flow_set_params runs after every optimiser step).
//
// The flat parameter vector (include/gnf_b200.h: which -> half -> step, per MLP W0 b0 W1 b1 ...) is repacked into
// every image a kernel family consumes: zero-padded fp32 W/b (k_linear), transposed fp32 (backward dX GEMMs), the
// fp16 and bf16 hi/lo UMMA chunk images of the fused tcgen05 kernels, their transposed bf16 images (backward dX
// chain), the bias rows, and the f1 attention projections / LayerNorm rows.  A job table is built once per flow
// (shapes are fixed); one kernel walks it, a block belongs to exactly one job.
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace gnf {
namespace {

enum PackKind : int { kPack32 = 0, kPack32T = 1, kPackTc = 2, kPackBias = 3, kPackMat = 4, kPackLn = 5 };

struct PackJob {
  int kind, transposed;
  int in, out;            // logical layer dims (rows of W = in, columns = out; Sonnet y = x @ W + b)
  int p0, p1, p2, p3;     // kind specific padding / chunk geometry
  int64_t src_off;        // floats into the flat parameter vector
  void *d0, *d1;          // destinations
  int first_block, n_blocks;
};

constexpr int kPackThreads = 256;

__global__ void __launch_bounds__(kPackThreads) k_pack_all(const PackJob* __restrict__ jobs, int n_jobs,
                                                           const float* __restrict__ params) {
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {                                   // last job whose first_block <= blockIdx.x (block uniform)
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob j = jobs[lo];
  const int i = ((int)blockIdx.x - j.first_block) * kPackThreads + (int)threadIdx.x;
  const float* __restrict__ src = params + j.src_off;
  switch (j.kind) {
    case kPack32: {        // p0 = in_pad, p1 = out_pad; d0 = W [in_pad, out_pad], d1 = b [out_pad]
      const int in_pad = j.p0, out_pad = j.p1;
      float* w = (float*)j.d0;
      float* b = (float*)j.d1;
      if (i < in_pad * out_pad) {
        const int r = i / out_pad, c = i - r * out_pad;
        w[i] = (r < j.in && c < j.out) ? src[r * j.out + c] : 0.f;
      }
      if (i < out_pad) b[i] = (i < j.out) ? src[j.in * j.out + i] : 0.f;
      break;
    }
    case kPack32T: {       // p0 = in_pad, p1 = out_pad8; d0 = W^T [out_pad8, in_pad]
      const int in_pad = j.p0;
      if (i < in_pad * j.p1) {
        const int r = i / in_pad, c = i - r * in_pad;
        ((float*)j.d0)[i] = (r < j.out && c < j.in) ? src[c * j.out + r] : 0.f;
      }
      break;
    }
    case kPackMat: {       // bias-free projection: p0 = in_pad, p1 = out_pad; d0 = W [in_pad, out_pad]
      const int out_pad = j.p1;
      if (i < j.p0 * out_pad) {
        const int r = i / out_pad, c = i - r * out_pad;
        ((float*)j.d0)[i] = (r < j.in && c < j.out) ? src[r * j.out + c] : 0.f;
      }
      break;
    }
    case kPackBias: {      // d0 = 256 floats
      if (i < 256) ((float*)j.d0)[i] = i < j.out ? src[i] : 0.f;
      break;
    }
    case kPackLn: {        // in = H, p0 = HP; src = gamma[H] beta[H]; d0 = [2][HP]
      const int hp = j.p0;
      if (i < 2 * hp) {
        const int which = i / hp, f = i - which * hp;
        ((float*)j.d0)[i] = f < j.in ? src[which * j.in + f] : 0.f;
      }
      break;
    }
    case kPackTc: {
      // One logical layer W[in, out] -> chunked UMMA B images (B[n][k] = W[k][n]), K-major no-swizzle core matrices
      // (coupling_tc.cu header).  p0 = kpad, p1 = npad, p2 = nhc (N per accumulator half), p3 = kcc (K per chunk).
      // transposed: the logical layer is W^T, i.e. src is stored [out, in] row-major.
      const int kpad = j.p0, npad = j.p1, nhc = j.p2, kcc = j.p3;
      if (i >= kpad * npad) break;
      const int n = i / kpad, k = i - n * kpad;
      const float w = (k < j.in && n < j.out) ? (j.transposed ? src[n * j.in + k] : src[k * j.out + n]) : 0.f;
      const int ph = n / nhc, nl = n - ph * nhc, kc = k / kcc, kl = k - kc * kcc;
      const int mat_bytes = nhc * kcc * 2;
      const size_t chunk = (size_t)(ph * (kpad / kcc) + kc) * (2 * mat_bytes);
      const size_t off = chunk + (size_t)(kl >> 3) * (nhc * 16) + (nl >> 3) * 128 + (nl & 7) * 16 + (kl & 7) * 2;
      if (j.d0) {
        uint8_t* img = (uint8_t*)j.d0;
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        *reinterpret_cast<__half*>(img + off) = h;
        *reinterpret_cast<__half*>(img + off + mat_bytes) = l;
      }
      {
        uint8_t* img = (uint8_t*)j.d1;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
        *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(img + off + mat_bytes) = l;
      }
      break;
    }
  }
}

// NS = accumulator column splits of the consuming kernel: kFwdNS (k_coupling_tc) or kNS (k_bwd_chain)
template <int LAT, int NS>
void tc_layer_geometry_t(int pos, int K, int& kpad, int& npad, int& nhc, int& kcc, size_t& bytes) {
  using G = tcx::Geo<LAT, NS>;
  if (pos == 0) { kpad = tcx::kK0; npad = LAT; nhc = LAT; kcc = tcx::kK0; bytes = G::L0_BYTES; }
  else if (pos == K - 1) { kpad = LAT; npad = tcx::kNOut; nhc = tcx::kNOut; kcc = LAT; bytes = G::LAST_BYTES; }
  else { kpad = LAT; npad = LAT; nhc = G::NH; kcc = G::KC; bytes = (size_t)NS * G::NKC * G::CHUNK_BYTES; }
}
void tc_layer_geometry(int L, bool forward_kernel, int pos, int K, int& kpad, int& npad, int& nhc, int& kcc, size_t& bytes) {
  if (L == 256) {
    if (forward_kernel) tc_layer_geometry_t<256, tcx::kFwdNS>(pos, K, kpad, npad, nhc, kcc, bytes);
    else tc_layer_geometry_t<256, tcx::kNS>(pos, K, kpad, npad, nhc, kcc, bytes);
  } else {
    if (forward_kernel) tc_layer_geometry_t<128, tcx::kFwdNS>(pos, K, kpad, npad, nhc, kcc, bytes);
    else tc_layer_geometry_t<128, tcx::kNS>(pos, K, kpad, npad, nhc, kcc, bytes);
  }
}

}  // namespace

// Builds the job table of a freshly created flow (all device buffers already allocated) and uploads it.
int pack_build_jobs(Flow& f) {
  std::vector<PackJob> jobs;
  int blocks = 0;
  auto add = [&](PackJob j, int64_t elems) {
    j.first_block = blocks;
    j.n_blocks = (int)ceil_div(elems, kPackThreads);
    blocks += j.n_blocks;
    jobs.push_back(j);
  };
  for (int m = 0; m < f.n_mlps; ++m) {
    const int64_t base = (int64_t)m * f.params_per_mlp;
    if (f.attn) {
      const int qk = f.heads * f.kq, hv = f.heads * f.vd;
      float* wa = f.wattn + (int64_t)m * f.wattn_per_mlp;
      float* wt = f.wattnT + (int64_t)m * f.wattnT_per_mlp;
      struct Proj { int64_t src; int in, out, in_pad, out_pad; int64_t off, offT; };
      const Proj pr[4] = {
          {0, f.H, qk, f.hp8, f.qk_pad, f.wq_off, f.wqT_off},
          {(int64_t)f.H * qk, f.H, qk, f.hp8, f.qk_pad, f.wk_off, f.wkT_off},
          {2ll * f.H * qk, f.H, f.vd, f.hp8, f.v_pad, f.wv_off, f.wvT_off},
          {2ll * f.H * qk + (int64_t)f.H * f.vd, hv, f.cho, f.hv_pad, f.cho_pad, f.wo_off, f.woT_off}};
      for (const Proj& q : pr) {
        PackJob j{};
        j.kind = kPackMat; j.in = q.in; j.out = q.out; j.p0 = q.in_pad; j.p1 = q.out_pad;
        j.src_off = base + q.src; j.d0 = wa + q.off;
        add(j, (int64_t)q.in_pad * q.out_pad);
        PackJob t{};
        t.kind = kPack32T; t.in = q.in; t.out = q.out; t.p0 = q.in_pad; t.p1 = q.out_pad;
        t.src_off = base + q.src; t.d0 = wt + q.offT;
        add(t, (int64_t)q.in_pad * q.out_pad);
      }
      if (f.attn_flags & GNF_ATTN_LAYER_NORM) {
        PackJob j{};
        j.kind = kPackLn; j.in = f.H; j.p0 = f.HP; j.src_off = base + f.ln_off;
        j.d0 = f.wln + (int64_t)m * 2 * f.HP;
        add(j, 2 * f.HP);
      }
    }
    if (f.wlin[0]) {      // images of the stand-alone tensor-core linears (linear_tc.cu): chunk geometry kcc = 16, one N block
      const int qk = f.heads * f.kq, hv = f.heads * f.vd;
      // (k, n) = logical GEMM dims; transposed: the parameter block is stored [n, k] row-major (the backward's W^T)
      struct Lin { int slot; int64_t src; int k, n, transposed; };
      const Lin lins[7] = {{0, 0, f.H, qk, 0},
                           {1, (int64_t)f.H * qk, f.H, qk, 0},
                           {2, 2ll * f.H * qk, f.H, f.vd, 0},
                           {3, 2ll * f.H * qk + (int64_t)f.H * f.vd, hv, f.cho, 0},
                           {4, f.flat_w_off[0], f.in_dim, f.L, 0},
                           {5, f.flat_w_off[0], f.L, f.in_dim, 1},
                           {6, 2ll * f.H * qk + (int64_t)f.H * f.vd, f.cho, hv, 1}};
      for (const Lin& q : lins) {
        if (f.lin_off[q.slot] < 0) continue;
        PackJob j{};
        j.kind = kPackTc; j.transposed = q.transposed; j.in = q.k; j.out = q.n;
        j.p0 = pad16(q.k); j.p1 = pad16(q.n); j.p2 = pad16(q.n); j.p3 = 16;
        j.src_off = base + q.src;
        j.d0 = f.wlin[0] + (size_t)m * f.wlin_per_mlp + f.lin_off[q.slot];
        j.d1 = f.wlin[1] + (size_t)m * f.wlin_per_mlp + f.lin_off[q.slot];
        add(j, (int64_t)j.p0 * j.p1);
      }
    }
    if (f.wgemm[0])       // layered path (gemm_tc.cu): every MLP layer, column blocks of <= 256, k16 slabs
      for (int l = 0; l < f.K; ++l) {
        int kpad, npad, nb;
        tc_gemm_geometry(f.ins[l], f.outs[l], kpad, npad, nb);
        PackJob j{};
        j.kind = kPackTc; j.transposed = 0; j.in = f.ins[l]; j.out = f.outs[l];
        j.p0 = kpad; j.p1 = npad; j.p2 = nb; j.p3 = 16;
        j.src_off = base + f.flat_w_off[l];
        j.d0 = f.wgemm[0] + (size_t)m * f.wgemm_per_mlp + f.gemm_off[l];
        j.d1 = f.wgemm[1] + (size_t)m * f.wgemm_per_mlp + f.gemm_off[l];
        add(j, (int64_t)kpad * npad);
        if (f.wgemmT) {     // W_l^T for the backward's dX: logical [outs, ins], parameter block stored [ins, outs]; bf16 only
          tc_gemm_geometry(f.outs[l], f.ins[l], kpad, npad, nb);
          PackJob t{};
          t.kind = kPackTc; t.transposed = 1; t.in = f.outs[l]; t.out = f.ins[l];
          t.p0 = kpad; t.p1 = npad; t.p2 = nb; t.p3 = 16;
          t.src_off = base + f.flat_w_off[l];
          t.d0 = nullptr;
          t.d1 = f.wgemmT + (size_t)m * f.wgemmT_per_mlp + f.gemmT_off[l];
          add(t, (int64_t)kpad * npad);
        }
      }
    size_t tc_off[kMaxLayers + 1] = {0};
    if (f.tc_ok || f.tc_inject)
      for (int pos = 0; pos < f.K; ++pos) {
        int a, b, c, d; size_t bytes;
        tc_layer_geometry(f.L, true, pos, f.K, a, b, c, d, bytes);       // (byte counts do not depend on the split)
        tc_off[pos + 1] = tc_off[pos] + bytes;
      }
    for (int l = 0; l < f.K; ++l) {
      const int64_t wsrc = base + f.flat_w_off[l];
      {
        PackJob j{};
        j.kind = kPack32; j.in = f.ins[l]; j.out = f.outs[l]; j.p0 = f.in_pads[l]; j.p1 = f.out_pads[l];
        j.src_off = wsrc;
        j.d0 = f.w32 + (int64_t)m * f.w32_per_mlp + f.w32_layer_off[l];
        j.d1 = f.w32 + (int64_t)m * f.w32_per_mlp + f.b32_layer_off[l];
        add(j, (int64_t)f.in_pads[l] * f.out_pads[l] > f.out_pads[l] ? (int64_t)f.in_pads[l] * f.out_pads[l] : f.out_pads[l]);
      }
      {
        PackJob j{};
        j.kind = kPack32T; j.in = f.ins[l]; j.out = f.outs[l]; j.p0 = f.in_pads[l]; j.p1 = f.out_pad8[l];
        j.src_off = wsrc;
        j.d0 = f.w32T + (int64_t)m * f.w32T_per_mlp + f.w32T_layer_off[l];
        add(j, (int64_t)f.in_pads[l] * f.out_pad8[l]);
      }
      if (!f.tc_ok && !f.tc_inject) continue;
      {   // forward chain position l (the layer-0 image is not used by inject flows: their layer 0 is wider than kK0)
        int kpad, npad, nhc, kcc; size_t bytes;
        tc_layer_geometry(f.L, true, l, f.K, kpad, npad, nhc, kcc, bytes);
        PackJob j{};
        j.kind = kPackTc; j.transposed = 0; j.in = f.ins[l]; j.out = f.outs[l];
        j.p0 = kpad; j.p1 = npad; j.p2 = nhc; j.p3 = kcc; j.src_off = wsrc;
        j.d0 = f.wtc[0] + (size_t)m * f.wtc_per_mlp + tc_off[l];
        j.d1 = f.wtc[1] + (size_t)m * f.wtc_per_mlp + tc_off[l];
        if (!(l == 0 && f.tc_inject)) add(j, (int64_t)kpad * npad);
        if (f.wtcB[0]) {   // the same layer in the backward kernel's chunk geometry (its forward recompute chains)
          tc_layer_geometry(f.L, false, l, f.K, kpad, npad, nhc, kcc, bytes);
          PackJob jb = j;
          jb.p0 = kpad; jb.p1 = npad; jb.p2 = nhc; jb.p3 = kcc;
          jb.d0 = f.wtcB[0] + (size_t)m * f.wtc_per_mlp + tc_off[l];
          jb.d1 = f.wtcB[1] + (size_t)m * f.wtc_per_mlp + tc_off[l];
          add(jb, (int64_t)kpad * npad);
        }
        PackJob b{};
        b.kind = kPackBias; b.out = f.outs[l]; b.src_off = base + f.flat_b_off[l];
        b.d0 = f.btc + ((size_t)m * f.K + l) * 256;
        add(b, 256);
      }
      if (f.wtcT) {   // backward dX chain: position K-1-l applies W_l^T (no bias), same chunk geometry
        const int pos = f.K - 1 - l;
        int kpad, npad, nhc, kcc; size_t bytes;
        tc_layer_geometry(f.L, false, pos, f.K, kpad, npad, nhc, kcc, bytes);
        PackJob j{};
        j.kind = kPackTc; j.transposed = 1; j.in = f.outs[l]; j.out = f.ins[l];
        j.p0 = kpad; j.p1 = npad; j.p2 = nhc; j.p3 = kcc; j.src_off = wsrc;
        j.d0 = nullptr;
        j.d1 = f.wtcT + (size_t)m * f.wtc_per_mlp + tc_off[pos];
        add(j, (int64_t)kpad * npad);
      }
    }
  }
  f.n_pack_jobs = (int)jobs.size();
  f.pack_blocks = blocks;
  GNF_CUDA(cudaMalloc(&f.pack_jobs, jobs.size() * sizeof(PackJob)));
  GNF_CUDA(cudaMemcpy(f.pack_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  return GNF_OK;
}

// one kPackTc image of a row-major W [k, n] on its own (debug entry of gemm_tc.cu); job_scratch: device, >= 256 bytes
int pack_tc_image(const float* W, int k, int n, int kpad, int npad, int nhc, int kcc, uint8_t* img_f16, uint8_t* img_bf16,
                  void* job_scratch, cudaStream_t stream) {
  static_assert(sizeof(PackJob) <= 256, "");
  PackJob j{};
  j.kind = kPackTc; j.transposed = 0; j.in = k; j.out = n;
  j.p0 = kpad; j.p1 = npad; j.p2 = nhc; j.p3 = kcc;
  j.src_off = 0; j.d0 = img_f16; j.d1 = img_bf16;
  j.first_block = 0;
  j.n_blocks = (int)ceil_div((int64_t)kpad * npad, kPackThreads);
  GNF_CUDA(cudaMemcpyAsync(job_scratch, &j, sizeof(j), cudaMemcpyHostToDevice, stream));
  GNF_CUDA(cudaStreamSynchronize(stream));       // `j` lives on this stack frame
  k_pack_all<<<(unsigned)j.n_blocks, kPackThreads, 0, stream>>>((const PackJob*)job_scratch, 1, W);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

int pack_all(const Flow& f, const float* params, cudaStream_t stream) {
  if (f.pack_blocks == 0) return GNF_OK;
  k_pack_all<<<(unsigned)f.pack_blocks, kPackThreads, 0, stream>>>((const PackJob*)f.pack_jobs, f.n_pack_jobs, params);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace gnf
