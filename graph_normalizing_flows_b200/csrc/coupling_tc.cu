// Fused half coupling step on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// One persistent CTA per SM walks 128-node tiles.  Per tile and per MLP (s, then t):
//
//   gather warps   : stable-CSR in-order aggregation of sender rows (a3+a4), assemble the MLP
//                    input h = concat[x_a, agg] | eps*x_a + agg (a5), write it to shared memory
//                    as the K-major UMMA A operand (16-bit hi [+ lo] split)           -> h_full
//   producer warp  : streams the pre-packed weight chunks (already in the UMMA canonical
//                    no-swizzle K-major shared-memory image) global/L2 -> smem ring with
//                    cp.async.bulk + mbarrier complete_tx                              -> full[s]
//   MMA warp       : one thread issues tcgen05.mma.  Layer 0 reads A from smem (SS), every other
//                    layer reads A straight from TMEM (TS).  fp32 accumulators live in TMEM.
//                    3 MMAs per product in the split modes (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo).
//   epilogue warps : tcgen05.ld the accumulator, + bias, activation, split to hi/lo 16-bit,
//                    tcgen05.st it back IN PLACE as the next layer's A operand         -> a_ready
//                    last layer: s / t in registers, then x_b <- x_b*exp(s)+t (or the inverse)
//                    and the log-det partial (fp64, fixed order).
//
// TMEM map (512 columns): two regions R0 = [0, LAT), R1 = [LAT, 2 LAT).  Layer l accumulates
// into R[l&1]; its epilogue rewrites each 32-column fp32 chunk as 16 columns of packed hi and
// 16 columns of packed lo, which is exactly where the next layer's TS MMAs read A from.  The
// accumulator of a layer is produced in NS = 2 column halves with separate commits, so the
// epilogue of half 0 overlaps the MMAs of half 1, and the next layer starts as soon as the
// K-range it needs has been converted.
//
// Reference semantics: NodeBlockGNN (gnn.py:143-156), ConcatThenMLPBlock / AggThenMLPBlock
// (gnn.py:100-126), make_mlp_model (gnn.py:159-180), coupling update (gnn.py:320-323,335-338,
// 353-359,366-372).
#include "common.cuh"

namespace gnf {
namespace {

constexpr int kTileM = 128;
constexpr int kK0 = 16;      // padded MLP input width  (in_dim <= 16)
constexpr int kNOut = 16;    // padded MLP output width (H <= 16)
constexpr int kStages = 5;
constexpr int kStageBytes = 32768;
constexpr int kThreads = 320;  // warp0 producer, warp1 MMA, warps2-5 epilogue, warps6-9 gather
constexpr int kEpiThreads = 128;
constexpr int kGatherThreads = 128;
constexpr int kNS = 2;

template <int LAT>
struct Geo {
  static constexpr int NH = LAT / kNS;                                  // acc half width
  static constexpr int KC = (16384 / (2 * NH)) < LAT ? (16384 / (2 * NH)) : LAT;  // K per chunk
  static constexpr int NKC = LAT / KC;
  static constexpr int MAT_BYTES = NH * KC * 2;                          // one 16-bit matrix
  static constexpr int CHUNK_BYTES = 2 * MAT_BYTES;                      // hi + lo
  static constexpr int L0_MAT_BYTES = LAT * kK0 * 2;
  static constexpr int L0_BYTES = 2 * L0_MAT_BYTES;
  static constexpr int LAST_MAT_BYTES = kNOut * LAT * 2;
  static constexpr int LAST_BYTES = 2 * LAST_MAT_BYTES;
  static_assert(CHUNK_BYTES <= kStageBytes && L0_BYTES <= kStageBytes && LAST_BYTES <= kStageBytes, "");
};

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d),
      "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 version field = 1):
// 8-row x 16-byte core matrices, SBO between 8-row groups, LBO between the two K halves.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int n, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16-bit hi/lo split of two fp32 values; element 0 in the low half-word (K order in TMEM / smem)
template <bool BF16>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float2 hf = __bfloat1622float2(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
  }
}

__device__ __forceinline__ float act_fn(float v, int act) {
  return act == GNF_ACT_LEAKY_RELU ? fmaxf(v, 0.2f * v) : fmaxf(v, 0.f);
}

struct TcParams {
  const float* xa;
  float* xb;
  const int32_t* rowptr;
  const int32_t* csr;
  int64_t n_nodes;
  int n_tiles;
  const uint8_t* w[2];     // weight images of the s and t MLP
  const float* bias[2];    // [K][256]
  int K, H, HP, concat, mean, act, inverse;
  float eps;
  double* partials;
};

struct __align__(8) Barriers {
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full[kNS], a_ready[kNS];
  uint64_t h_full[2], h_empty[2];
};

template <int LAT>
constexpr size_t smem_bytes() {
  return 1024 /*align slack*/ + (size_t)kStages * kStageBytes + 2 * 2 * 4096 /*h tiles*/ +
         2 * kMaxLayers * LAT * 4 /*bias*/ + kGatherThreads * 17 * 4 /*h staging*/ + sizeof(Barriers) + 64;
}

template <int LAT, int NPROD, bool BF16>
__global__ void __launch_bounds__(kThreads, 1) k_coupling_tc(const TcParams p) {
  using G = Geo<LAT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* hbuf = ring + kStages * kStageBytes;                 // [buf][hi|lo][4096]
  float* bias_s = (float*)(hbuf + 2 * 2 * 4096);                // [2][kMaxLayers][LAT]
  float* hstage = bias_s + 2 * kMaxLayers * LAT;                // [128][17]
  Barriers* bars = (Barriers*)(hstage + kGatherThreads * 17);
  uint32_t* tmem_slot = (uint32_t*)(bars + 1);
  double* ldj_red = (double*)(tmem_slot + 2);                   // [4]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = p.K;

  // ---- one-time setup -----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < kNS; ++i) {
      mbar_init(smem_u32(&bars->acc_full[i]), 1);
      mbar_init(smem_u32(&bars->a_ready[i]), kEpiThreads);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->h_full[i]), kGatherThreads);
      mbar_init(smem_u32(&bars->h_empty[i]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  for (int i = tid; i < 2 * K * LAT; i += kThreads) {
    int m = i / (K * LAT), r = i - m * (K * LAT);
    int l = r / LAT, c = r - l * LAT;
    bias_s[(m * kMaxLayers + l) * LAT + c] = p.bias[m][l * 256 + c];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== weight producer =====================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto issue = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&bars->full[stage]);
        mbar_expect_tx(fb, bytes);
        bulk_g2s(smem_u32(ring + stage * kStageBytes), src, bytes, fb);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int m = 0; m < 2; ++m) {
          const uint8_t* src = p.w[m];
          issue(src, NPROD == 3 ? G::L0_BYTES : G::L0_MAT_BYTES);
          src += G::L0_BYTES;
          for (int l = 1; l < K - 1; ++l)
            for (int c = 0; c < kNS * G::NKC; ++c) {
              issue(src, NPROD == 3 ? G::CHUNK_BYTES : G::MAT_BYTES);
              src += G::CHUNK_BYTES;
            }
          issue(src, NPROD == 3 ? G::LAST_BYTES : G::LAST_MAT_BYTES);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ==========================================================================
    if (lane == 0) {
      constexpr uint32_t idesc_l0 = make_idesc(LAT, BF16);
      constexpr uint32_t idesc_h = make_idesc(G::NH, BF16);
      constexpr uint32_t idesc_last = make_idesc(kNOut, BF16);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t aready_par[kNS] = {0, 0};
      int region = 0;
      int it = 0;
      // column of the packed hi half of feature k inside a region (lo is 16 columns further)
      auto a_col = [](int k) { return (k >> 5) * 32 + ((k & 31) >> 4) * 8; };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&bars->h_full[buf]), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t h_hi = smem_u32(hbuf + buf * 8192);
        const uint32_t h_lo = h_hi + 4096;
        for (int m = 0; m < 2; ++m) {
          // ---- layer 0: A = h (smem), B = W0 chunk, N = LAT, K = 16 ---------------------------
          {
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(ring + stage * kStageBytes);
            const uint64_t a_hi = smem_desc(h_hi, 2048, 128), a_lo = smem_desc(h_lo, 2048, 128);
            const uint64_t b_hi = smem_desc(sb, LAT * 16, 128);
            const uint64_t b_lo = smem_desc(sb + G::L0_MAT_BYTES, LAT * 16, 128);
            const uint32_t d = tmem_base + region * LAT;
            mma_ss(d, a_hi, b_hi, idesc_l0, 0);
            if (NPROD == 3) {
              mma_ss(d, a_lo, b_hi, idesc_l0, 1);
              mma_ss(d, a_hi, b_lo, idesc_l0, 1);
            }
            tc_commit(smem_u32(&bars->empty[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            tc_commit(smem_u32(&bars->acc_full[0]));
            tc_commit(smem_u32(&bars->acc_full[1]));
            if (m == 1) tc_commit(smem_u32(&bars->h_empty[buf]));
            region ^= 1;
          }
          // ---- hidden layers: A from TMEM (previous region), N = NH per half -----------------
          for (int l = 1; l < K - 1; ++l) {
            const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
            const uint32_t out_col = tmem_base + region * LAT;
            bool waited[kNS] = {false, false};
            for (int ph = 0; ph < kNS; ++ph) {
              for (int kc = 0; kc < G::NKC; ++kc) {
                const int q_lo = (kc * G::KC) / G::NH, q_hi = (kc * G::KC + G::KC - 1) / G::NH;
                for (int q = q_lo; q <= q_hi; ++q)
                  if (!waited[q]) {
                    mbar_wait(smem_u32(&bars->a_ready[q]), aready_par[q]);
                    aready_par[q] ^= 1;
                    waited[q] = true;
                  }
                mbar_wait(smem_u32(&bars->full[stage]), phase);
                tc_fence_after();
                const uint32_t sb = smem_u32(ring + stage * kStageBytes);
#pragma unroll
                for (int ks = 0; ks < G::KC / 16; ++ks) {
                  const int kg = kc * G::KC + ks * 16;
                  const uint32_t a_hi = in_col + a_col(kg);
                  const uint64_t b_hi = smem_desc(sb + ks * 2 * (G::NH * 16), G::NH * 16, 128);
                  const uint32_t acc = (kc | ks) ? 1u : 0u;
                  const uint32_t d = out_col + ph * G::NH;
                  mma_ts(d, a_hi, b_hi, idesc_h, acc);
                  if (NPROD == 3) {
                    const uint64_t b_lo =
                        smem_desc(sb + G::MAT_BYTES + ks * 2 * (G::NH * 16), G::NH * 16, 128);
                    mma_ts(d, a_hi + 16, b_hi, idesc_h, 1);
                    mma_ts(d, a_hi, b_lo, idesc_h, 1);
                  }
                }
                tc_commit(smem_u32(&bars->empty[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
              tc_commit(smem_u32(&bars->acc_full[ph]));
            }
            region ^= 1;
          }
          // ---- last layer: N = 16, K = LAT, one chunk -----------------------------------------
          {
            const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
            const uint32_t d = tmem_base + region * LAT;
            for (int q = 0; q < kNS; ++q) {
              mbar_wait(smem_u32(&bars->a_ready[q]), aready_par[q]);
              aready_par[q] ^= 1;
            }
            mbar_wait(smem_u32(&bars->full[stage]), phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(ring + stage * kStageBytes);
#pragma unroll 4
            for (int ks = 0; ks < LAT / 16; ++ks) {
              const uint32_t a_hi = in_col + a_col(ks * 16);
              const uint64_t b_hi = smem_desc(sb + ks * 2 * (kNOut * 16), kNOut * 16, 128);
              mma_ts(d, a_hi, b_hi, idesc_last, ks ? 1u : 0u);
              if (NPROD == 3) {
                const uint64_t b_lo =
                    smem_desc(sb + G::LAST_MAT_BYTES + ks * 2 * (kNOut * 16), kNOut * 16, 128);
                mma_ts(d, a_hi + 16, b_hi, idesc_last, 1);
                mma_ts(d, a_hi, b_lo, idesc_last, 1);
              }
            }
            tc_commit(smem_u32(&bars->empty[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            tc_commit(smem_u32(&bars->acc_full[0]));
            region ^= 1;
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue warps ======================================================================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par[kNS] = {0, 0};
    int region = 0;
    double ldj_local = 0.0;
    const int hp4 = p.HP >> 2;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      float st[2][kNOut];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        for (int l = 0; l < K - 1; ++l) {
          const float* bl = bias_s + (m * kMaxLayers + l) * LAT;
#pragma unroll
          for (int ph = 0; ph < kNS; ++ph) {
            mbar_wait(smem_u32(&bars->acc_full[ph]), acc_par[ph]);
            acc_par[ph] ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < G::NH / 32; ++c) {
              const int col = ph * G::NH + c * 32;
              const uint32_t taddr = lane_base + region * LAT + col;
              uint32_t v[32];
              tmem_ld32(taddr, v);
              tmem_wait_ld();
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 bb = *reinterpret_cast<const float2*>(bl + col + 2 * j);
                float a = act_fn(__uint_as_float(v[2 * j]) + bb.x, p.act);
                float b = act_fn(__uint_as_float(v[2 * j + 1]) + bb.y, p.act);
                split_pair<BF16>(a, b, hi[j], lo[j]);
              }
              tmem_st16(taddr, hi);
              if (NPROD == 3) tmem_st16(taddr + 16, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->a_ready[ph]));
          }
          region ^= 1;
        }
        // last layer: s (m == 0) or t (m == 1)
        {
          const float* bl = bias_s + (m * kMaxLayers + (K - 1)) * LAT;
          mbar_wait(smem_u32(&bars->acc_full[0]), acc_par[0]);
          acc_par[0] ^= 1;
          tc_fence_after();
          uint32_t v[16];
          tmem_ld16(lane_base + region * LAT, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < kNOut; ++j) st[m][j] = __uint_as_float(v[j]) + bl[j];
          region ^= 1;
        }
      }
      // ---- affine coupling update + log-det partial (gnn.py:322-323 / :359) ------------------
      const int64_t node = (int64_t)tile * kTileM + row;
      if (node < p.n_nodes) {
        float* xrow = p.xb + node * p.HP;
#pragma unroll
        for (int g4 = 0; g4 < kNOut / 4; ++g4) {
          if (g4 < hp4) {
            float4 x = *reinterpret_cast<const float4*>(xrow + g4 * 4);
            float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int f = g4 * 4 + j;
              const float s = st[0][f], t = st[1][f];
              if (f < p.H) {
                if (!p.inverse) {
                  xv[j] = __fadd_rn(__fmul_rn(xv[j], expf(s)), t);
                  ldj_local += (double)s;
                } else {
                  xv[j] = __fmul_rn(__fsub_rn(xv[j], t), expf(-s));
                }
              }
            }
            *reinterpret_cast<float4*>(xrow + g4 * 4) = make_float4(xv[0], xv[1], xv[2], xv[3]);
          }
        }
      }
    }
    // fixed-order reduction of the log-det partial: lanes, then the 4 epilogue warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ldj_local += __shfl_down_sync(0xffffffffu, ldj_local, o);
    if (lane == 0) ldj_red[warp - 2] = ldj_local;
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
    if (warp == 2 && lane == 0 && p.partials)
      p.partials[blockIdx.x] = ((ldj_red[0] + ldj_red[1]) + ldj_red[2]) + ldj_red[3];
  } else {
    // ===== gather warps: a3 + a4 + a5 =========================================================
    const int row = tid - (kThreads - kGatherThreads);
    float* my = hstage + row * 17;
    const int hp4 = p.HP >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int64_t node = (int64_t)tile * kTileM + row;
      float self[kNOut], agg[kNOut];
#pragma unroll
      for (int j = 0; j < kNOut; ++j) { self[j] = 0.f; agg[j] = 0.f; }
      if (node < p.n_nodes) {
        const float* xr = p.xa + node * p.HP;
#pragma unroll
        for (int g4 = 0; g4 < kNOut / 4; ++g4)
          if (g4 < hp4) {
            float4 x = *reinterpret_cast<const float4*>(xr + g4 * 4);
            self[g4 * 4] = x.x; self[g4 * 4 + 1] = x.y; self[g4 * 4 + 2] = x.z; self[g4 * 4 + 3] = x.w;
          }
        int32_t e = p.rowptr[node];
        const int32_t end = p.rowptr[node + 1];
        const int32_t cnt = end - e;
        for (; e < end; ++e) {       // ascending edge index inside the segment: TF-CPU order
          const float* sr = p.xa + (int64_t)p.csr[e] * p.HP;
#pragma unroll
          for (int g4 = 0; g4 < kNOut / 4; ++g4)
            if (g4 < hp4) {
              float4 x = *reinterpret_cast<const float4*>(sr + g4 * 4);
              agg[g4 * 4] = __fadd_rn(agg[g4 * 4], x.x);
              agg[g4 * 4 + 1] = __fadd_rn(agg[g4 * 4 + 1], x.y);
              agg[g4 * 4 + 2] = __fadd_rn(agg[g4 * 4 + 2], x.z);
              agg[g4 * 4 + 3] = __fadd_rn(agg[g4 * 4 + 3], x.w);
            }
        }
        if (p.mean) {
          const float dv = fmaxf((float)cnt, 1.f);
#pragma unroll
          for (int j = 0; j < kNOut; ++j) agg[j] = __fdiv_rn(agg[j], dv);
        }
      }
      // assemble h[16] through a private smem row (runtime H needs dynamic placement)
#pragma unroll
      for (int j = 0; j < kK0; ++j) my[j] = 0.f;
      if (p.concat) {
#pragma unroll
        for (int j = 0; j < kNOut; ++j)
          if (j < p.H) { my[j] = self[j]; my[p.H + j] = agg[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < kNOut; ++j)
          if (j < p.H) my[j] = __fadd_rn(__fmul_rn(p.eps, self[j]), agg[j]);
      }
      __syncwarp();
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_pair<BF16>(my[2 * j], my[2 * j + 1], hi[j], lo[j]);
      mbar_wait(smem_u32(&bars->h_empty[buf]), ((it >> 1) & 1) ^ 1);
      uint8_t* hb = hbuf + buf * 8192;
      *reinterpret_cast<uint4*>(hb + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(hb + 2048 + row * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (NPROD == 3) {
        *reinterpret_cast<uint4*>(hb + 4096 + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(hb + 4096 + 2048 + row * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->h_full[buf]));
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- weight image packing ------------------------------------------------------------------
// One logical layer W[in,out] -> chunked UMMA B images (B[n][k] = W[k][n]); see header comment.
__global__ void k_pack_tc(const float* __restrict__ W, int in, int out, int kpad, int npad, int nhc,
                          int kcc, uint8_t* __restrict__ img_f16, uint8_t* __restrict__ img_bf16) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kpad * npad) return;
  const int n = i / kpad, k = i - n * kpad;
  const float w = (k < in && n < out) ? W[k * out + n] : 0.f;
  const int ph = n / nhc, nl = n - ph * nhc, kc = k / kcc, kl = k - kc * kcc;
  const int mat_bytes = nhc * kcc * 2;
  const size_t chunk = (size_t)(ph * (kpad / kcc) + kc) * (2 * mat_bytes);
  const size_t off = chunk + (size_t)(kl >> 3) * (nhc * 16) + (nl >> 3) * 128 + (nl & 7) * 16 + (kl & 7) * 2;
  {
    __half h = __float2half_rn(w);
    __half l = __float2half_rn(w - __half2float(h));
    *reinterpret_cast<__half*>(img_f16 + off) = h;
    *reinterpret_cast<__half*>(img_f16 + off + mat_bytes) = l;
  }
  {
    __nv_bfloat16 h = __float2bfloat16_rn(w);
    __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(img_bf16 + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img_bf16 + off + mat_bytes) = l;
  }
}

__global__ void k_pack_bias(const float* __restrict__ b, int out, float* __restrict__ dst) {
  int i = threadIdx.x;
  if (i < 256) dst[i] = i < out ? b[i] : 0.f;
}

template <int LAT>
size_t bytes_per_mlp_t(int K) {
  using G = Geo<LAT>;
  return (size_t)G::L0_BYTES + (size_t)(K - 2) * kNS * G::NKC * G::CHUNK_BYTES + G::LAST_BYTES;
}

template <int LAT>
int pack_mlp_t(const Flow& f, int mlp, const float* params, cudaStream_t stream) {
  using G = Geo<LAT>;
  uint8_t* i16 = f.wtc[0] + (size_t)mlp * f.wtc_per_mlp;
  uint8_t* ibf = f.wtc[1] + (size_t)mlp * f.wtc_per_mlp;
  float* bias = f.btc + (size_t)mlp * f.K * 256;
  const float* src = params;
  size_t off = 0;
  for (int l = 0; l < f.K; ++l) {
    const int in = f.ins[l], out = f.outs[l];
    int kpad, npad, nhc, kcc;
    size_t bytes;
    if (l == 0) { kpad = kK0; npad = LAT; nhc = LAT; kcc = kK0; bytes = G::L0_BYTES; }
    else if (l == f.K - 1) { kpad = LAT; npad = kNOut; nhc = kNOut; kcc = LAT; bytes = G::LAST_BYTES; }
    else { kpad = LAT; npad = LAT; nhc = G::NH; kcc = G::KC; bytes = (size_t)kNS * G::NKC * G::CHUNK_BYTES; }
    k_pack_tc<<<(unsigned)ceil_div((int64_t)kpad * npad, 256), 256, 0, stream>>>(src, in, out, kpad, npad, nhc,
                                                                                 kcc, i16 + off, ibf + off);
    GNF_LAUNCH_CHECK();
    k_pack_bias<<<1, 256, 0, stream>>>(src + (size_t)in * out, out, bias + l * 256);
    GNF_LAUNCH_CHECK();
    off += bytes;
    src += (size_t)in * out + out;
  }
  return GNF_OK;
}

template <int LAT, int NPROD, bool BF16>
int launch_tc(const TcParams& p, int grid, cudaStream_t stream) {
  auto kern = k_coupling_tc<LAT, NPROD, BF16>;
  static bool configured = false;
  const size_t smem = smem_bytes<LAT>();
  if (!configured) {
    GNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  kern<<<grid, kThreads, smem, stream>>>(p);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace

bool tc_shape_supported(const Flow& f) {
  return (f.L == 128 || f.L == 256) && f.in_dim <= kK0 && f.H <= kNOut && f.K >= 2 && f.K <= kMaxLayers;
}

size_t tc_bytes_per_mlp(int L, int K) { return L == 256 ? bytes_per_mlp_t<256>(K) : bytes_per_mlp_t<128>(K); }

int tc_pack_mlp(const Flow& f, int mlp, const float* params, void* stream) {
  return f.L == 256 ? pack_mlp_t<256>(f, mlp, params, (cudaStream_t)stream)
                    : pack_mlp_t<128>(f, mlp, params, (cudaStream_t)stream);
}

int tc_coupling_half(const Flow& f, int mlp_s, int mlp_t, int math, int inverse, const float* xa,
                     float* xb, int64_t n_nodes, const int32_t* rowptr, const int32_t* csr_senders,
                     double* ldj_partials, int* n_partials, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcParams p;
  p.xa = xa;
  p.xb = xb;
  p.rowptr = rowptr;
  p.csr = csr_senders;
  p.n_nodes = n_nodes;
  p.n_tiles = (int)ceil_div(n_nodes, kTileM);
  const int img = (math == GNF_MATH_TC3X) ? 0 : 1;
  p.w[0] = f.wtc[img] + (size_t)mlp_s * f.wtc_per_mlp;
  p.w[1] = f.wtc[img] + (size_t)mlp_t * f.wtc_per_mlp;
  p.bias[0] = f.btc + (size_t)mlp_s * f.K * 256;
  p.bias[1] = f.btc + (size_t)mlp_t * f.K * 256;
  p.K = f.K;
  p.H = f.H;
  p.HP = f.HP;
  p.concat = f.d.block == GNF_BLOCK_CONCAT;
  p.mean = f.d.agg == GNF_AGG_MEAN;
  p.act = f.d.act;
  p.inverse = inverse;
  p.eps = f.d.eps;
  p.partials = ldj_partials;
  int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  *n_partials = grid;
  if (f.L == 256) {
    if (math == GNF_MATH_TC3X) return launch_tc<256, 3, false>(p, grid, stream);
    if (math == GNF_MATH_TC3X_BF16) return launch_tc<256, 3, true>(p, grid, stream);
    return launch_tc<256, 1, true>(p, grid, stream);
  }
  if (math == GNF_MATH_TC3X) return launch_tc<128, 3, false>(p, grid, stream);
  if (math == GNF_MATH_TC3X_BF16) return launch_tc<128, 3, true>(p, grid, stream);
  return launch_tc<128, 1, true>(p, grid, stream);
}

}  // namespace gnf
