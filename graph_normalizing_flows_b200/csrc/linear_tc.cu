// One Sonnet Linear (+ bias, + activation) on the 5th-gen tensor cores, fp32 in / fp32 out (SURVEY 8 row a6 for the
// shapes the fused coupling kernel does not take: MLP inputs wider than 16 -- layer 0 of the dm_self_attn GNN and of
// wide message-passing blocks -- and the attention block's q / k / v / output projections, gnn.py:509-545).
//
//   C[M, :N] (+)= act(A[M, :K] @ W[K, N] + b)       K <= 256, N <= 256 (both padded to 16), image + A half <= 220 KB
//
// Arithmetic as the fused kernel's: both operands split into 16-bit hi/lo parts, three tcgen05.mma per product
// (A_hi B_hi + A_lo B_hi + A_hi B_lo), fp32 accumulation in TMEM -- fp32-class results (2^-22 operand error).
//
// One persistent CTA per SM (256 threads).  The whole weight image (<= 128 KB: it IS the K-major no-swizzle UMMA
// shared-memory layout, written by pack.cu) is pulled into shared memory once per CTA with a single cp.async.bulk.
// Per 128-row tile: all threads read the fp32 rows (full 32-byte sectors per thread), split them and write the K-major A
// tile (128 columns of K at a time: K = 256 -- the backward's dX = delta W^T with K = latent_dim -- takes two passes
// into the same accumulator); one elected thread issues the MMAs; all eight warps drain the accumulator (tcgen05.ld,
// + bias, activation, optionally + the previous C) and write fp32 rows.  No overlap between the three stages: the op is
// small next to the fused kernel it feeds, and even so it replaces an FFMA GEMM that took 10-20x longer.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace gnf {
namespace {

using namespace tcx;

constexpr int kLinThreads = 256;
constexpr int kLinKPass = 128;          // K columns staged per pass
constexpr int kLinMaxK = 256;
constexpr int kLinMaxN = 256;
constexpr size_t kLinMaxSmem = 220 * 1024;

struct LinParams {
  const float* A;
  float* C;
  const uint8_t* wimg;      // [Kp/16][hi: N x 16 | lo: N x 16] 16-bit, K-major core matrices
  const float* bias;        // [>= N] (zeros for the bias-free projections)
  int64_t M;
  int lda, kvalid;          // row stride of A in floats; floats of a row that may be read (multiple of 4)
  int ldc, nvalid;          // row stride of C; floats of a row that may be written (multiple of 4)
  int Kp, N;                // padded to 16
  int act;                  // GNF_ACT_* or 2 = none
  int accumulate;           // 1: C += result (after bias and activation)
  int n_tiles;
  int* range_flag;
};

__device__ __forceinline__ float lin_act(float v, int act) {
  if (act == GNF_ACT_LEAKY_RELU) return fmaxf(v, 0.2f * v);
  if (act == GNF_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

template <int NPROD, bool BF16>
__global__ void __launch_bounds__(kLinThreads, 1) k_linear_tc(const LinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int Kp = p.Kp, N = p.N;
  const uint32_t w_bytes = (uint32_t)(Kp / 16) * 2u * (uint32_t)N * 32u;
  uint8_t* wsm = smem;                                   // weight image
  const int Kb = Kp < kLinKPass ? Kp : kLinKPass;        // K columns per pass
  uint8_t* a_hi = wsm + ((w_bytes + 1023) & ~1023u);     // [Kb/8][128 rows][16 B]
  uint8_t* a_lo = a_hi + 128 * Kb * 2;
  uint64_t* bars = (uint64_t*)(a_lo + 128 * Kb * 2);     // [0] weights landed, [1] MMAs of a pass done
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(smem_u32(&bars[0]), w_bytes);
    bulk_g2s(smem_u32(wsm), p.wimg, w_bytes, smem_u32(&bars[0]));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  mbar_wait(smem_u32(&bars[0]), 0);

  const uint32_t idesc = make_idesc(N, BF16);
  const int row = tid & 127, khalf = tid >> 7;           // A staging: a thread owns one row and every other 8-k group
  const int q = warp & 3, chalf = warp >> 2;             // epilogue: TMEM lane quarter and 16-column chunk parity
  float amax = 0.f;
  uint32_t acc_par = 0;
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int64_t r = (int64_t)tile * kTileM + row;
    const float* ar = p.A + r * p.lda;
    for (int k0 = 0; k0 < Kp; k0 += kLinKPass) {
      const int kw = (Kp - k0) < kLinKPass ? (Kp - k0) : kLinKPass;
      // ---- A tile: fp32 rows -> hi/lo 16-bit, K-major core matrices (8 rows x 16 B), LBO = 2048, SBO = 128 --------
      for (int g = khalf; g < kw / 8; g += 2) {
        float v[8];
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {
          const int k = k0 + g * 8 + h4 * 4;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < p.M && k + 4 <= p.kvalid) x = *reinterpret_cast<const float4*>(ar + k);
          v[h4 * 4] = x.x; v[h4 * 4 + 1] = x.y; v[h4 * 4 + 2] = x.z; v[h4 * 4 + 3] = x.w;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (!BF16) amax = fmaxf(amax, fmaxf(fabsf(v[2 * j]), fabsf(v[2 * j + 1])));
          split_pair<BF16>(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        }
        *reinterpret_cast<uint4*>(a_hi + g * 2048 + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (NPROD == 3) *reinterpret_cast<uint4*>(a_lo + g * 2048 + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      __syncthreads();
      // ---- MMAs: one elected thread ------------------------------------------------------------------------------
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          for (int ks = 0; ks < kw / 16; ++ks) {
            const uint64_t ah = smem_desc(smem_u32(a_hi) + ks * 4096, 2048, 128);
            const uint64_t al = smem_desc(smem_u32(a_lo) + ks * 4096, 2048, 128);
            const uint32_t wb = smem_u32(wsm) + (k0 / 16 + ks) * (2 * N * 32);
            const uint64_t bh = smem_desc(wb, N * 16, 128);
            const uint64_t bl = smem_desc(wb + N * 32, N * 16, 128);
            mma_ss(tmem_base, ah, bh, idesc, (k0 || ks) ? 1u : 0u);
            if (NPROD == 3) mma_ss(tmem_base, al, bh, idesc, 1);
            if (NPROD >= 2) mma_ss(tmem_base, ah, bl, idesc, 1);
          }
          tc_commit(smem_u32(&bars[1]));
        }
        __syncwarp();
      }
      // the pass's MMAs have read the A tile (and, after the last pass, the accumulator is complete)
      mbar_wait(smem_u32(&bars[1]), acc_par);
      acc_par ^= 1u;
      tc_fence_after();
    }
    // ---- epilogue: 16-column chunks, chunk c by the warps with chalf == (c & 1) -------------------------------------
    const int64_t orow = (int64_t)tile * kTileM + q * 32 + lane;
    for (int c = chalf; c < N / 16; c += 2) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c * 16, v);
      tmem_wait_ld();
      if (orow < p.M) {
        float* out = p.C + orow * p.ldc + c * 16;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int col = c * 16 + j4 * 4;
          if (col + 4 <= p.nvalid) {
            float4 o;
            o.x = lin_act(__uint_as_float(v[j4 * 4]) + p.bias[col], p.act);
            o.y = lin_act(__uint_as_float(v[j4 * 4 + 1]) + p.bias[col + 1], p.act);
            o.z = lin_act(__uint_as_float(v[j4 * 4 + 2]) + p.bias[col + 2], p.act);
            o.w = lin_act(__uint_as_float(v[j4 * 4 + 3]) + p.bias[col + 3], p.act);
            if (p.accumulate) {
              const float4 c0 = *reinterpret_cast<const float4*>(out + j4 * 4);
              o.x += c0.x; o.y += c0.y; o.z += c0.z; o.w += c0.w;
            }
            *reinterpret_cast<float4*>(out + j4 * 4) = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();           // accumulator drained and the A tile consumed before the next tile overwrites them
  }
  if (!BF16 && amax > 65504.f && p.range_flag) *p.range_flag = 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

size_t lin_smem_bytes(int Kp, int N) {
  const uint32_t w_bytes = (uint32_t)(Kp / 16) * 2u * (uint32_t)N * 32u;
  const int kb = Kp < kLinKPass ? Kp : kLinKPass;
  return 1024 + ((w_bytes + 1023) & ~1023u) + 2 * 128 * (size_t)kb * 2 + 64;
}

template <int NPROD, bool BF16>
int launch_linear(const LinParams& p, cudaStream_t stream) {
  auto kern = k_linear_tc<NPROD, BF16>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    GNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLinMaxSmem));
  const size_t smem = lin_smem_bytes(p.Kp, p.N);
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kLinThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = getenv("GNF_NO_PDL") == nullptr ? 1 : 0;
  GNF_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace

bool tc_linear_shape_ok(int k, int n) {
  return k >= 1 && n >= 1 && pad16(k) <= kLinMaxK && pad16(n) <= kLinMaxN && lin_smem_bytes(pad16(k), pad16(n)) <= kLinMaxSmem;
}
size_t tc_linear_image_bytes(int k, int n) { return (size_t)(pad16(k) / 16) * 2 * pad16(n) * 32; }

// A [M, lda] (kvalid readable floats per row), image of W [k, n] (pack.cu, kPackTc geometry kcc = 16), C [M, ldc]
int tc_linear(const Flow& f, int math, const float* A, int lda, int kvalid, const uint8_t* img_f16, const uint8_t* img_bf16,
              int k, int n, const float* bias, int act, float* C, int ldc, int nvalid, int64_t M, cudaStream_t stream,
              int accumulate) {
  if (M == 0) return GNF_OK;
  LinParams p;
  p.A = A;
  p.C = C;
  const bool bf = !(math == GNF_MATH_TC3X || math == GNF_MATH_TC2X);
  p.wimg = bf ? img_bf16 : img_f16;
  p.bias = bias;
  p.M = M;
  p.lda = lda;
  p.kvalid = kvalid;
  p.ldc = ldc;
  p.nvalid = nvalid;
  p.Kp = pad16(k);
  p.N = pad16(n);
  p.act = act;
  p.accumulate = accumulate;
  p.n_tiles = (int)ceil_div(M, kTileM);
  p.range_flag = f.range_flag;
  // tc2x keeps fp16 activations unsplit in the fused kernel; a stand-alone layer keeps all three products (its inputs
  // are raw features / attention outputs, not bounded activations)
  if (math == GNF_MATH_BF16) return launch_linear<1, true>(p, stream);
  if (bf) return launch_linear<3, true>(p, stream);
  return launch_linear<3, false>(p, stream);
}

}  // namespace gnf
