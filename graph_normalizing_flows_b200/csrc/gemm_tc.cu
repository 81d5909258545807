// General Sonnet Linear (+ bias, + activation) on the 5th-gen tensor cores for the shapes neither the fused coupling
// kernel (latent_dim in {128, 256}) nor the single-image k_linear_tc (K, N <= 256) takes: the embedding-flow defaults of
// train_grevnet_with_data.py:41-47,104-117 (node_embedding_dim 200, gnn_latent_dim 2048, 3 layers) and any other wide MLP
// (gnn.py:143-180 make_mlp_model).
//
//   C[M, :N] = act(A[M, :K] @ W[K, N] + b)      fp32 in / fp32 out, any K and N (padded to 16 in the image)
//
// Arithmetic as everywhere else on this path: both operands split into 16-bit hi/lo parts, three tcgen05.mma per product
// (A_hi B_hi + A_lo B_hi + A_hi B_lo), fp32 accumulation in TMEM.
//
// Work item = 256 rows (two 128-row UMMA tiles, one accumulator each: all 512 TMEM columns) x one block of <= 256 output
// columns; items are dealt round-robin to one persistent CTA per SM with the column block running fastest, so the CTAs
// that share an A row block run at the same time and read it from L2.  Warp roles (320 threads):
//   warp 0   weight producer: one cp.async.bulk per k16 slab of the column block (the image -- pack.cu, kPackTc geometry
//            nhc = block width, kcc = 16 -- IS the K-major no-swizzle UMMA layout), 6-stage ring, mbarrier expect_tx
//   warp 1   MMA issuer: per slab 2 row tiles x 3 products of M128 x N<=256 x K16; tcgen05.commit frees the slab / A stage
//   warps 2-9 converters, then epilogue: take the fp32 A rows of a 32-k stage, split them and write the K-major A tile
//            (a lane owns 8 consecutive k of a row); after the last slab they drain both accumulators (tcgen05.ld,
//            + bias, activation) into fp32 rows of C.
// How the fp32 A rows reach the converters (template flag TMA):
//   TMA = true (default): the producer warp issues ONE cp.async.bulk.tensor.2d per stage through a tensor map of A
//            (box = 32 floats x 256 rows, SWIZZLE_128B, zero fill past M / past the readable row width) into a raw
//            shared-memory ring; the converters read it back with conflict-free 16-byte loads (chunk index XOR row).
//            Measured: the per-lane global loads of the other path kept the L1 request path 58 % busy and were the
//            kernel's limiter (profiles/r2_ncu_full_k_gemm_tc_ldg.csv, GNF_GEMM_VARIANT experiments).
//   TMA = false (GNF_GEMM_TMA=0): per-lane LDG.128, three stages of loads in flight in registers.
// The A tile of a row block is converted once per column block (re-read from L2, not from HBM); sharing each weight slab
// between two row tiles halves the weight stream per FLOP (the fused kernel's figure is ~43 B/cycle/SM).
#include <stdlib.h>
#include <cuda.h>              // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no -lcuda)
#include "common.cuh"
#include "tc_common.cuh"

namespace gnf {
namespace {

using namespace tcx;

constexpr int kGemmThreads = 320;
constexpr int kGemmConv = 256;           // converter / epilogue threads (warps 2..9)
constexpr int kGemmRows = 256;           // rows per work item (two UMMA tiles)
constexpr int kGemmNB = 256;             // widest column block
constexpr int kGemmWStage = kGemmNB * 64;   // one k16 slab: hi + lo, 16 k x 256 columns x 2 B
constexpr int kGemmAK = 32;              // k per A stage
constexpr int kGemmAStage = kGemmRows * kGemmAK * 2 * 2;   // converted stage, hi + lo: 32 KB
constexpr int kGemmRawStage = kGemmRows * kGemmAK * 4;     // raw fp32 stage (TMA path): 32 KB
constexpr int kGemmMaxRing = 6;
// ring depths: weight slabs, converted A stages, raw A stages
template <bool TMA>
struct GemmCfg {
  static constexpr int WS = TMA ? 5 : 6, AS = TMA ? 2 : 3, RS = TMA ? 2 : 0;
  static constexpr size_t kSmem = 1024 + (size_t)WS * kGemmWStage + (size_t)AS * kGemmAStage + (size_t)RS * kGemmRawStage + 256;
};

struct GemmParams {
  const float* A;
  float* C;
  const uint8_t* wimg;      // [n_blocks][Kp/16][hi: nb x 16 | lo: nb x 16] 16-bit, K-major core matrices
  const float* bias;        // [>= nvalid]
  int64_t M;
  int lda, kvalid;          // row stride of A in floats; floats of a row that may be read (multiple of 4)
  int ldc, nvalid;          // row stride of C; floats of a row that may be written (multiple of 4)
  int Kp;                   // K padded to 16
  int nb, n_blocks;         // column block width (multiple of 16, <= 256) and count
  int rows;                 // rows per work item: 256 (two row tiles per weight slab), or 128 when that leaves SMs idle
  int act;                  // GNF_ACT_* or 2 = none
  int n_items;              // row blocks x column blocks
  int variant;              // GNF_GEMM_VARIANT (timing experiments only; results are wrong for bits 1, 2)
  int* range_flag;
};

struct GemmBars {
  uint64_t w_full[kGemmMaxRing], w_empty[kGemmMaxRing];
  uint64_t a_full[kGemmMaxRing], a_empty[kGemmMaxRing];
  uint64_t raw_full[kGemmMaxRing], raw_empty[kGemmMaxRing];
  uint64_t acc_full, acc_empty;
};

// one box of the tensor map into shared memory, completion on an mbarrier (coordinates: innermost first)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__device__ __forceinline__ float gemm_act(float v, int act) {
  if (act == GNF_ACT_LEAKY_RELU) return fmaxf(v, 0.2f * v);
  if (act == GNF_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// The CTAs of a wave walk K in step; started at the same k they would all pull the same few weight slabs (and the same
// k columns of A) out of the same L2 lines at the same moment.  Every item starts its K loop at its own 32-k stage and
// wraps around (the sum order of an output element depends on the item only, not on the SM that ran it).
__device__ __forceinline__ int stage_offset(const GemmParams& p, int rb, int cb, int n_astages) {
  if (p.variant & 32) return 0;
  return (int)(((unsigned)rb * 7u + (unsigned)cb * 3u) % (unsigned)n_astages);
}

template <int NPROD, bool BF16, bool TMA>
__global__ void __launch_bounds__(kGemmThreads, 1) k_gemm_tc(const GemmParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int kGemmWS = GemmCfg<TMA>::WS, kGemmAS = GemmCfg<TMA>::AS, kGemmRS = GemmCfg<TMA>::RS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* rawring = smem;                                  // (TMA path; 1024-byte aligned stages: SWIZZLE_128B atoms)
  uint8_t* wring = rawring + kGemmRS * kGemmRawStage;
  uint8_t* aring = wring + kGemmWS * kGemmWStage;
  GemmBars* bars = (GemmBars*)(aring + kGemmAS * kGemmAStage);
  uint32_t* tmem_slot = (uint32_t*)(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = p.Kp, nb = p.nb;
  const int n_slabs = Kp / 16;
  const int n_astages = (Kp + kGemmAK - 1) / kGemmAK;
  const uint32_t slab_bytes = (uint32_t)nb * 64u;

  if (tid == 0) {
    for (int i = 0; i < kGemmWS; ++i) {
      mbar_init(smem_u32(&bars->w_full[i]), 1);
      mbar_init(smem_u32(&bars->w_empty[i]), 1);
    }
    for (int i = 0; i < kGemmAS; ++i) {
      mbar_init(smem_u32(&bars->a_full[i]), kGemmConv / 32);
      mbar_init(smem_u32(&bars->a_empty[i]), 1);
    }
    for (int i = 0; i < kGemmRS; ++i) {
      mbar_init(smem_u32(&bars->raw_full[i]), 1);
      mbar_init(smem_u32(&bars->raw_empty[i]), kGemmConv / 32);
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_empty), kGemmConv / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== weight producer ===================================================================================
    if (lane == 0) {
      int stage = 0, rstage = 0;
      uint32_t phase = 0, rphase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int rb = item / p.n_blocks, cb = item - rb * p.n_blocks;
        const uint8_t* src = p.wimg + (size_t)cb * n_slabs * slab_bytes;
        const int off = stage_offset(p, rb, cb, n_astages);
        for (int as = 0; as < n_astages; ++as) {
          const int asr = as + off < n_astages ? as + off : as + off - n_astages;
          const int ksteps = (Kp - asr * kGemmAK) >= kGemmAK ? 2 : 1;
          if constexpr (TMA) {      // the stage's fp32 rows: one box (32 floats x p.rows rows), zero filled out of bounds
            mbar_wait(smem_u32(&bars->raw_empty[rstage]), rphase ^ 1);
            const uint32_t fb = smem_u32(&bars->raw_full[rstage]);
            mbar_expect_tx(fb, (uint32_t)p.rows * (kGemmAK * 4));
            tma_load_2d(smem_u32(rawring + rstage * kGemmRawStage), &tmap, asr * kGemmAK, rb * p.rows, fb);
            if (++rstage == kGemmRS) { rstage = 0; rphase ^= 1; }
          }
          for (int ks = 0; ks < ksteps; ++ks) {
            const int s = asr * 2 + ks;
            mbar_wait(smem_u32(&bars->w_empty[stage]), phase ^ 1);
            const uint32_t fb = smem_u32(&bars->w_full[stage]);
            mbar_expect_tx(fb, slab_bytes);
            bulk_g2s(smem_u32(wring + stage * kGemmWStage), src + (size_t)s * slab_bytes, slab_bytes, fb);
            if (++stage == kGemmWS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule, one elected lane issues =============================
    const uint32_t idesc = make_idesc(nb, BF16);
    const uint32_t wring_u = smem_u32(wring), aring_u = smem_u32(aring);
    uint32_t wstage = 0, wphase = 0, astage = 0, aphase = 0, it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int rb = item / p.n_blocks, cb = item - rb * p.n_blocks;
      const int off = stage_offset(p, rb, cb, n_astages);
      const int n_rt = (p.rows > kTileM && (int64_t)rb * p.rows + kTileM < p.M) ? 2 : 1;   // second row tile past M: skipped
      mbar_wait(smem_u32(&bars->acc_empty), (it & 1u) ^ 1u);                 // previous item's accumulators drained
      tc_fence_after();
      for (int as = 0; as < n_astages; ++as) {
        mbar_wait(smem_u32(&bars->a_full[astage]), aphase);
        const int asr = as + off < n_astages ? as + off : as + off - n_astages;
        const int ksteps = (Kp - asr * kGemmAK) >= kGemmAK ? 2 : 1;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(smem_u32(&bars->w_full[wstage]), wphase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t wb = wring_u + wstage * kGemmWStage;
            const uint64_t bh = smem_desc(wb, nb * 16, 128);
            const uint64_t bl = smem_desc(wb + nb * 32, nb * 16, 128);
            const uint32_t acc = (as || ks) ? 1u : 0u;
            if (p.variant & 1) {          // experiment: alternate the two accumulators
              const uint32_t ab0 = aring_u + astage * kGemmAStage + ks * 4096, ab1 = ab0 + 16384;
              mma_ss(tmem_base, smem_desc(ab0, 2048, 128), bh, idesc, acc);
              mma_ss(tmem_base + kGemmNB, smem_desc(ab1, 2048, 128), bh, idesc, acc);
              mma_ss(tmem_base, smem_desc(ab0 + 8192, 2048, 128), bh, idesc, 1);
              mma_ss(tmem_base + kGemmNB, smem_desc(ab1 + 8192, 2048, 128), bh, idesc, 1);
              mma_ss(tmem_base, smem_desc(ab0, 2048, 128), bl, idesc, 1);
              mma_ss(tmem_base + kGemmNB, smem_desc(ab1, 2048, 128), bl, idesc, 1);
            } else
            for (int rt = 0; rt < ((p.variant & 4) ? 1 : n_rt); ++rt) {
              const uint32_t ab = aring_u + astage * kGemmAStage + rt * 16384 + ks * 4096;
              const uint64_t ah = smem_desc(ab, 2048, 128);
              const uint64_t al = smem_desc(ab + 8192, 2048, 128);
              const uint32_t d = tmem_base + rt * kGemmNB;
              mma_ss(d, ah, bh, idesc, acc);
              if (NPROD == 3) mma_ss(d, al, bh, idesc, 1);
              if (NPROD >= 2) mma_ss(d, ah, bl, idesc, 1);
            }
            tc_commit(smem_u32(&bars->w_empty[wstage]));
            if (ks == ksteps - 1) tc_commit(smem_u32(&bars->a_empty[astage]));
            if (as == n_astages - 1 && ks == ksteps - 1) tc_commit(smem_u32(&bars->acc_full));
          }
          __syncwarp();
          if (++wstage == kGemmWS) { wstage = 0; wphase ^= 1; }
        }
        if (++astage == kGemmAS) { astage = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===== converters, then epilogue =========================================================================
    const int ct = tid - 64, cw = ct >> 5;               // 0..255, converter warp 0..7
    const int r8 = lane & 7, g = lane >> 3;              // row inside an 8-row core matrix, k group (8 k) inside a stage
    const int q = warp & 3, chalf = cw >> 2;             // epilogue: TMEM lane quarter (warp id mod 4), column chunk parity
    float amax = 0.f;
    uint32_t astage = 0, aphase = 0, it = 0;
    [[maybe_unused]] uint32_t rstage = 0, rphase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int rb = item / p.n_blocks, cb = item - rb * p.n_blocks;
      const int64_t row0 = (int64_t)rb * p.rows;
      const int off = stage_offset(p, rb, cb, n_astages);
      auto real_stage = [&](int as) { return as + off < n_astages ? as + off : as + off - n_astages; };
      // a converter thread serves rows (j * 8 + cw) * 8 + r8, j = 0..3, k group g of every stage
      float4 b0[8], b1[8], b2[8];         // three stages of loads in flight: the conversion never waits a full latency
      auto load_stage = [&](int as, float4 (&v)[8]) {
        const int k = real_stage(as) * kGemmAK + g * 8;
        if (p.variant & 16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = make_float4(1.f, 2.f, 3.f, 4.f);
          return;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int rl = (j * 8 + cw) * 8 + r8;
          const int64_t r = rl < p.rows ? row0 + rl : p.M;          // 128-row items: the upper half of the tile is zeros
          const float* ar = p.A + r * p.lda + k;
          v[2 * j] = (r < p.M && k + 4 <= p.kvalid) ? __ldg(reinterpret_cast<const float4*>(ar)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[2 * j + 1] = (r < p.M && k + 8 <= p.kvalid) ? __ldg(reinterpret_cast<const float4*>(ar + 4))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto convert_stage = [&](int as, const float4 (&cur)[8]) {
        mbar_wait(smem_u32(&bars->a_empty[astage]), aphase ^ 1);
        if (real_stage(as) * kGemmAK + g * 8 < Kp) {
          uint8_t* st = aring + astage * kGemmAStage + g * 2048;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = (j * 8 + cw) * 8 + r8;                 // 0..255
            const float v[8] = {cur[2 * j].x, cur[2 * j].y, cur[2 * j].z, cur[2 * j].w,
                                cur[2 * j + 1].x, cur[2 * j + 1].y, cur[2 * j + 1].z, cur[2 * j + 1].w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (!BF16) amax = fmaxf(amax, fmaxf(fabsf(v[2 * e]), fabsf(v[2 * e + 1])));
              split_pair<BF16>(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
            }
            uint8_t* dst = st + (rl >> 7) * 16384 + (rl & 127) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (NPROD == 3) *reinterpret_cast<uint4*>(dst + 8192) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (!(p.variant & 8)) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[astage]));
        if (++astage == kGemmAS) { astage = 0; aphase ^= 1; }
      };
      if constexpr (TMA) {
        // raw stage [row][128 B] with the 16-byte chunks of a row XOR-swizzled by (row & 7): the 8 lanes of a quarter
        // warp (8 rows, same k group) read 8 different chunk positions -- no bank conflicts
        for (int as = 0; as < n_astages; ++as) {
          float4 v[8];
          mbar_wait(smem_u32(&bars->raw_full[rstage]), rphase);
          const uint8_t* raw = rawring + rstage * kGemmRawStage;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = (j * 8 + cw) * 8 + r8;
            if (rl < p.rows) {
              const uint8_t* rr = raw + rl * 128;
              v[2 * j] = *reinterpret_cast<const float4*>(rr + (((2 * g) ^ r8) << 4));
              v[2 * j + 1] = *reinterpret_cast<const float4*>(rr + (((2 * g + 1) ^ r8) << 4));
            } else {
              v[2 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
              v[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->raw_empty[rstage]));     // values are in registers: refill the slot
          if (++rstage == kGemmRS) { rstage = 0; rphase ^= 1; }
          convert_stage(as, v);
        }
      } else if (p.variant & 2) {     // experiment: no loads, no conversion -- the converters only keep the ring going
        for (int as = 0; as < n_astages; ++as) {
          mbar_wait(smem_u32(&bars->a_empty[astage]), aphase ^ 1);
          if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[astage]));
          if (++astage == kGemmAS) { astage = 0; aphase ^= 1; }
        }
      } else {
      load_stage(0, b0);
      if (1 < n_astages) load_stage(1, b1);
      for (int as = 0; as < n_astages; as += 3) {
        if (as + 2 < n_astages) load_stage(as + 2, b2);
        convert_stage(as, b0);
        if (as + 1 < n_astages) {
          if (as + 3 < n_astages) load_stage(as + 3, b0);
          convert_stage(as + 1, b1);
        }
        if (as + 2 < n_astages) {
          if (as + 4 < n_astages) load_stage(as + 4, b1);
          convert_stage(as + 2, b2);
        }
      }
      }
      // ---- epilogue: both accumulators, 16-column chunks of parity chalf --------------------------------------
      mbar_wait(smem_u32(&bars->acc_full), it & 1u);
      tc_fence_after();
      const int n_rt = (p.rows > kTileM && row0 + kTileM < p.M) ? 2 : 1;
      const int col_base = cb * nb;
      for (int rt = 0; rt < n_rt; ++rt) {
        const int64_t orow = row0 + rt * kTileM + q * 32 + lane;
        for (int c = chalf; c < nb / 16; c += 2) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + rt * kGemmNB + c * 16, v);
          tmem_wait_ld();
          if (orow < p.M) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int col = col_base + c * 16 + j4 * 4;
              if (col + 4 <= p.nvalid) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                float4 o;
                o.x = gemm_act(__uint_as_float(v[j4 * 4]) + b.x, p.act);
                o.y = gemm_act(__uint_as_float(v[j4 * 4 + 1]) + b.y, p.act);
                o.z = gemm_act(__uint_as_float(v[j4 * 4 + 2]) + b.z, p.act);
                o.w = gemm_act(__uint_as_float(v[j4 * 4 + 3]) + b.w, p.act);
                *reinterpret_cast<float4*>(p.C + orow * p.ldc + col) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty));
    }
    if (!BF16 && amax > 65504.f && p.range_flag) *p.range_flag = 1;
  }
  __syncwarp();                // (warp 0: lane 0 walked the schedule alone) reconverge before the aligned barrier
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// tensor map of A [M, lda] fp32: dim 0 = the kvalid readable floats of a row, dim 1 = rows; box = 32 floats x `rows`
// rows; SWIZZLE_128B (a box row is exactly one 128-byte swizzle span); elements out of bounds read as zero
int encode_a_map(CUtensorMap* map, const float* A, int64_t M, int lda, int kvalid, int rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GNF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    GNF_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, GNF_ECUDA, "tc_gemm: cuTensorMapEncodeTiled not available");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)kvalid, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)lda * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kGemmAK, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)A, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GNF_REQUIRE(r == CUDA_SUCCESS, GNF_ECUDA, "tc_gemm: cuTensorMapEncodeTiled failed (%d) for M=%lld lda=%d kvalid=%d", (int)r,
              (long long)M, lda, kvalid);
  return GNF_OK;
}

template <int NPROD, bool BF16, bool TMA>
int launch_gemm_t(const GemmParams& p, cudaStream_t stream) {
  auto kern = k_gemm_tc<NPROD, BF16, TMA>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    GNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<TMA>::kSmem));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (TMA) {
    int rc = encode_a_map(&map, p.A, p.M, p.lda, p.kvalid, p.rows);
    if (rc) return rc;
  }
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  kern<<<grid, kGemmThreads, GemmCfg<TMA>::kSmem, stream>>>(p, map);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

template <int NPROD, bool BF16>
int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  const char* t = getenv("GNF_GEMM_TMA");
  if (t && t[0] == '0') return launch_gemm_t<NPROD, BF16, false>(p, stream);
  return launch_gemm_t<NPROD, BF16, true>(p, stream);
}

}  // namespace

// image geometry of W [k, n] for k_gemm_tc (pack.cu kPackTc: kpad, npad, nhc = block width, kcc = 16)
void tc_gemm_geometry(int k, int n, int& kpad, int& npad, int& nb) {
  kpad = pad16(k);
  const int n16 = pad16(n);
  nb = n16 < kGemmNB ? n16 : kGemmNB;
  npad = (n16 + nb - 1) / nb * nb;
}
size_t tc_gemm_image_bytes(int k, int n) {
  int kpad, npad, nb;
  tc_gemm_geometry(k, n, kpad, npad, nb);
  return (size_t)kpad * npad * 4;      // hi + lo, 2 bytes each
}

// A [M, lda] (kvalid readable floats per row), image of W [k, n] in the geometry above, C [M, ldc] (nvalid writable
// floats per row; bias has at least nvalid entries)
int tc_gemm(const Flow& f, int math, const float* A, int lda, int kvalid, const uint8_t* img_f16, const uint8_t* img_bf16,
            int k, int n, const float* bias, int act, float* C, int ldc, int nvalid, int64_t M, cudaStream_t stream) {
  if (M == 0) return GNF_OK;
  GNF_REQUIRE(lda % 4 == 0 && ldc % 4 == 0 && kvalid % 4 == 0 && nvalid % 4 == 0 && ((uintptr_t)A % 16) == 0 &&
                  ((uintptr_t)C % 16) == 0 && ((uintptr_t)bias % 16) == 0,
              GNF_EINVAL, "tc_gemm: operands must be float4 aligned");
  GemmParams p;
  int kpad, npad, nb;
  tc_gemm_geometry(k, n, kpad, npad, nb);
  const bool bf = !(math == GNF_MATH_TC3X || math == GNF_MATH_TC2X);
  p.A = A;
  p.C = C;
  p.wimg = bf ? img_bf16 : img_f16;
  p.bias = bias;
  p.M = M;
  p.lda = lda;
  p.kvalid = kvalid;
  p.ldc = ldc;
  p.nvalid = nvalid;
  p.Kp = kpad;
  p.nb = nb;
  p.n_blocks = npad / nb;
  p.act = act;
  // two row tiles share every weight slab (half the weight stream per FLOP) unless that leaves SMs without an item
  p.rows = (ceil_div(M, kGemmRows) * p.n_blocks < num_sms()) ? kTileM : kGemmRows;
  p.n_items = (int)ceil_div(M, p.rows) * p.n_blocks;
  p.range_flag = f.range_flag;
  const char* var = getenv("GNF_GEMM_VARIANT");
  p.variant = var ? atoi(var) : 0;
  if (math == GNF_MATH_BF16) return launch_gemm<1, true>(p, stream);
  if (bf) return launch_gemm<3, true>(p, stream);
  return launch_gemm<3, false>(p, stream);
}

}  // namespace gnf

using namespace gnf;

extern "C" size_t gnf_debug_linear_tc_workspace(int32_t k, int32_t n) {
  if (k < 1 || n < 1) return 0;
  return 2 * align_up(tc_gemm_image_bytes(k, n), 256) + 256;
}

extern "C" int gnf_debug_linear_tc(const float* a, const float* w, const float* bias, int64_t m, int32_t k, int32_t n,
                                   int32_t act, int32_t math, float* c, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(a && w && bias && c && m >= 0 && k >= 4 && n >= 4 && k % 4 == 0 && n % 4 == 0, GNF_EINVAL,
              "gnf_debug_linear_tc: null pointer or k, n not positive multiples of 4");
  GNF_REQUIRE(math == GNF_MATH_TC3X || math == GNF_MATH_TC3X_BF16 || math == GNF_MATH_BF16 || math == GNF_MATH_TC2X,
              GNF_EINVAL, "gnf_debug_linear_tc: bad math %d", math);
  GNF_REQUIRE(act >= 0 && act <= 2, GNF_EINVAL, "gnf_debug_linear_tc: bad act %d", act);
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= gnf_debug_linear_tc_workspace(k, n), GNF_EWORKSPACE,
              "gnf_debug_linear_tc: workspace too small or misaligned");
  int kpad, npad, nb;
  tc_gemm_geometry(k, n, kpad, npad, nb);
  const size_t img = align_up(tc_gemm_image_bytes(k, n), 256);
  uint8_t* i16 = (uint8_t*)ws;
  uint8_t* ibf = i16 + img;
  int rc = pack_tc_image(w, k, n, kpad, npad, nb, 16, i16, ibf, ibf + img, stream);
  if (rc) return rc;
  static Flow none{};          // only Flow::range_flag is read (null: no fp16 range report from the debug entry)
  return tc_gemm(none, math, a, k, k, i16, ibf, k, n, bias, act, c, n, n, m, stream);
}

