// Reversible backward of one half coupling step on the 5th-gen tensor cores.
//
// Two kernels per half step, both tcgen05 + TMEM, hi/lo split 16-bit operands (3 MMAs per product), fp32 accumulate:
//
//   k_bwd_chain : persistent, one CTA per SM, 128-node tiles.  Per tile FOUR MLP chains run through
//                 the same TMEM ping-pong as the forward kernel (coupling_tc.cu):
//                   F_s, F_t : recompute s = S(h), t = T(h) from xa (gather + segment-sum as in the
//                              forward; fp16 hi/lo in the tc3x/tc2x modes), every hidden activation a_l also
//                              written to HBM as a bf16 hi/lo image for the weight-gradient GEMM, its sign
//                              bits kept in shared memory for act';
//                   coupling : xb <- (xb' - t) * exp(-s)  (gnn.py:359,372),  g_s = g*xb*exp(s) - scale,
//                              g_t = g,  g_xb <- g * exp(s);
//                   B_s, B_t : delta_{l-1} = (delta_l W_l^T) (.) act'(a_{l-1}) with the TRANSPOSED weight
//                              images (bf16 hi/lo); delta_l images to HBM, g_h = delta_0 W_0^T out in fp32.
//   k_dw_tc     : dW_l = a_{l-1}^T delta_l over all nodes: M = input features, N = output features,
//                 K = nodes.  Both operands are MN-major (node index strided), read straight from
//                 the images with cp.async.bulk; split over node ranges, fixed-order reduction after.
//                 Its epilogue warps, idle during the main loop, sum the columns of the delta operand
//                 (read from the images through L2, in step with the ring): the bias gradients.
//
// Inject flows (template flag BINJ: MLP input wider than 16 columns -- dm_self_attn, message passing with D > 16): the
// forward chains start from the layer-0 PRE-activations computed by k_linear_tc (read from global memory by the epilogue
// warps, as k_coupling_tc MODE inject does), the backward chains stop at delta_0, which also leaves in fp32 row-major:
// g_h = delta_0 W_0^T is a k_linear_tc call (K = latent_dim), dW_0 = h^T delta_0 a pair of the weight-gradient GEMM whose
// B operand is an image of the MLP input built by k_make_images.  The gather warps idle.
//
// Image layout (one per 128-node tile, per part hi|lo): the UMMA no-swizzle MN-major canonical form,
//   elem(f, n) = (n>>3)*(F*8) + (f>>3)*64 + (n&7)*8 + (f&7)        F = feature count (LAT or 16)
// i.e. 8 features x 8 nodes core matrices of 128 B; SBO (feature groups) = 128 B, LBO (node groups) = F*16 B.
// An epilogue thread (one node row, 32 consecutive features) writes it with four 16-byte stores.
#include "common.cuh"
#include "tc_common.cuh"

namespace gnf {
namespace {

using namespace tcx;

constexpr bool kBF = true;      // element type of every GRADIENT operand: bf16 (gradients underflow fp16)
// The recomputed forward chains (template flag F16F) use the fp16 hi/lo split when the math mode is
// tc3x / tc2x: 2^-22 operands keep the recomputed pre-activations at fp32 accuracy, so the sign masks of
// act' flip no more often than in fp32 arithmetic (bf16 hi/lo, 2^-17, flips ~40x more near the kink).

struct BwdParams {
  const float* xa;
  float* xb;
  float* gxb;
  const int32_t* rowptr;
  const int32_t* csr;
  int64_t n_nodes;
  int n_tiles;
  const uint8_t* wf[2];      // forward weight images (fp16 or bf16 hi/lo, template flag F16F) of the s and t MLP
  const uint8_t* wb[2];      // transposed-chain weight images
  const float* bias[2];
  int K, H, HP, concat, mean;
  float eps, scale;
  uint16_t* act_img;         // [2][K-1][tiles][2][LAT*128]
  uint16_t* dlt_img;         // [2][K-1][tiles][2][LAT*128]
  uint16_t* h_img;           // [tiles][2][16*128]
  uint16_t* g_img;           // [2][tiles][2][16*128]
  float* gh;                 // [tiles*128][16]
  int write_lo;              // 0: the weight-gradient GEMM reads only the hi parts, skip the lo images
  const float* pre0[2];      // BINJ: layer-0 pre-activations (bias included) of the s and t MLP, [n, LAT] fp32
  float* d0[2];              // BINJ: delta_0 = dL/d(pre0) of the s and t MLP, [n, LAT] fp32
};

constexpr int kBStages = 3;          // weight ring stages (the forward kernel has 4; 32 KB go to the mask bits)
constexpr int kMaskLayers = 4;       // hidden activations per MLP whose act' sign bits fit the smem mask store
constexpr int kMaskWords = 2 * kMaskLayers * kNS * 2;   // per epilogue thread: [mlp][layer][acc half][chunk]

struct __align__(8) BwdBarriers {
  uint64_t full[kBStages], empty[kBStages];
  uint64_t acc_full[kNS], a_ready[kMaxNA];
  uint64_t h_full[2], h_empty[2];
  uint64_t g_full;
  uint64_t acc_last;   // last chain layer: its own barrier, so the next chain's layer 0 (which commits acc_full
                       // without waiting for the epilogue) can never advance a barrier two phases past a waiter
};

template <int LAT>
constexpr size_t bwd_smem_bytes() {
  return 1024 + (size_t)kBStages * kStageBytes + (size_t)kMaskWords * kEpiThreads * 4 /*mask bits*/ +
         2 * 8192 /*h tiles*/ + 2 * 8192 /*g tiles*/ +
         (kMaxLayers - 1) * 4096 + 2 * LAT * 32 + 2 * kNOut * 4 + kGatherThreads * 17 * 4 + sizeof(BwdBarriers) + 128;
}

template <int ACT>
__device__ __forceinline__ float act_f(float v) {
  return ACT == GNF_ACT_LEAKY_RELU ? fmaxf(v, 0.2f * v) : fmaxf(v, 0.f);
}
template <int ACT>
__device__ __forceinline__ float dact_f(bool positive) {
  return positive ? 1.f : (ACT == GNF_ACT_LEAKY_RELU ? 0.2f : 0.f);
}

// element offset of (feature group fg, node n) inside an image with F features
__device__ __forceinline__ size_t img_off(int F, int fg, int n) {
  return (size_t)(n >> 3) * (F * 8) + (size_t)fg * 64 + (n & 7) * 8;
}

template <int LAT, int ACT, bool F16F, bool BINJ>
__global__ void __launch_bounds__(kThreads, 1) k_bwd_chain(const BwdParams p) {
  constexpr bool kFB = !F16F;   // forward chains: bf16 split?
  using G = Geo<LAT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint32_t* mask_s = (uint32_t*)(ring + kBStages * kStageBytes);     // [kMaskWords][kEpiThreads] act' sign bits
  uint8_t* hbuf = (uint8_t*)(mask_s + kMaskWords * kEpiThreads);  // [buf][hi|lo][4096]
  uint8_t* gbuf = hbuf + 2 * 8192;                              // [m][hi|lo][4096]
  uint8_t* sel = gbuf + 2 * 8192;
  uint8_t* btile = sel + (kMaxLayers - 1) * 4096;
  float* blast = (float*)(btile + 2 * LAT * 32);
  float* hstage = blast + 2 * kNOut;
  BwdBarriers* bars = (BwdBarriers*)(hstage + kGatherThreads * 17);
  uint32_t* tmem_slot = (uint32_t*)(bars + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  const size_t img_elems = (size_t)LAT * 128;                   // one part of one tile
  const size_t layer_stride = (size_t)p.n_tiles * 2 * img_elems;

  if (tid == 0) {
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < kNS; ++i) mbar_init(smem_u32(&bars->acc_full[i]), 1);
    for (int i = 0; i < kMaxNA; ++i) mbar_init(smem_u32(&bars->a_ready[i]), kEpiThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->h_full[i]), kGatherThreads);
      mbar_init(smem_u32(&bars->h_empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->g_full), 128);
    mbar_init(smem_u32(&bars->acc_last), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {   // bias selector / bias tiles: as in the forward kernel
    const uint16_t one = F16F ? 0x3C00 : 0x3F80;
    for (int i = tid; i < (K - 1) * 128 * 16; i += kThreads) {
      const int l = i / 2048, r = i - l * 2048, m = r >> 4, k = r & 15;
      *reinterpret_cast<uint16_t*>(sel + l * 4096 + (k >> 3) * 2048 + m * 16 + (k & 7) * 2) =
          (k == 2 * l || k == 2 * l + 1) ? one : (uint16_t)0;
    }
    for (int i = tid; i < 2 * LAT * 8; i += kThreads) {
      const int m = i / (LAT * 8), r = i - m * (LAT * 8), n = r >> 3, j = r & 7;
      const float b = (j < K - 1) ? p.bias[m][j * 256 + n] : 0.f;
      uint32_t hi, lo;
      split_pair<kFB>(b, 0.f, hi, lo);
      const uint32_t packed = (hi & 0xFFFFu) | (lo << 16);
      const int k = 2 * j;
      *reinterpret_cast<uint32_t*>(btile + m * (LAT * 32) + (k >> 3) * (LAT * 16) + (n >> 3) * 128 + (n & 7) * 16 +
                                   (k & 7) * 2) = packed;
    }
    if (tid < 2 * kNOut) blast[tid] = p.bias[tid >> 4][(K - 1) * 256 + (tid & 15)];
    fence_proxy_async();
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== weight producer: F_s, F_t, B_s, B_t images per tile =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto issue = [&](const uint8_t* src, uint32_t bytes) {
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&bars->full[stage]);
        mbar_expect_tx(fb, bytes);
        bulk_g2s(smem_u32(ring + stage * kStageBytes), src, bytes, fb);
        if (++stage == kBStages) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < 4; ++c) {
          const uint8_t* src = c < 2 ? p.wf[c] : p.wb[c - 2];
          if (!(BINJ && c < 2)) issue(src, G::L0_BYTES);      // BINJ: the forward chains' layer 0 ran in k_linear_tc
          src += G::L0_BYTES;
          for (int l = 1; l < K - 1; ++l)
            for (int cc = 0; cc < kNS * G::NKC; ++cc) {
              issue(src, G::CHUNK_BYTES);
              src += G::CHUNK_BYTES;
            }
          if (!(BINJ && c >= 2)) issue(src, G::LAST_BYTES);   // BINJ: the backward chains stop at delta_0
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ==============================================================================
    constexpr uint32_t idesc_l0_f = make_idesc(LAT, kFB), idesc_l0_b = make_idesc(LAT, kBF);
    constexpr uint32_t idesc_h_f = make_idesc(G::NH, kFB), idesc_h_b = make_idesc(G::NH, kBF);
    constexpr uint32_t idesc_last_f = make_idesc(kNOut, kFB), idesc_last_b = make_idesc(kNOut, kBF);
    const uint32_t bar_full = smem_u32(&bars->full[0]), bar_empty = smem_u32(&bars->empty[0]);
    const uint32_t bar_acc = smem_u32(&bars->acc_full[0]), bar_ready = smem_u32(&bars->a_ready[0]);
    const uint32_t bar_hfull = smem_u32(&bars->h_full[0]), bar_hempty = smem_u32(&bars->h_empty[0]);
    const uint32_t bar_gfull = smem_u32(&bars->g_full), bar_acc_last = smem_u32(&bars->acc_last);
    const uint32_t ring_u = smem_u32(ring), sel_u = smem_u32(sel), btile_u = smem_u32(btile);
    const uint32_t hbuf_u = smem_u32(hbuf), gbuf_u = smem_u32(gbuf);
    uint32_t stage = 0, phase = 0;
    uint32_t aready_par = 0;
    uint32_t region = 0;
    int it = 0;
    auto a_col = [](int k) { return (uint32_t)((k >> 5) * 32 + ((k & 31) >> 4) * 8); };
    auto wait_groups = [&](uint32_t& waited, int q_lo, int q_hi) {
      for (int q = q_lo; q <= q_hi; ++q)
        if (!(waited >> q & 1u)) {
          mbar_wait(bar_ready + 8 * q, (aready_par >> q) & 1u);
          aready_par ^= 1u << q;
          waited |= 1u << q;
        }
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      for (int c = 0; c < 4; ++c) {
        const bool bwd = c >= 2;
        const int m = c & 1;
        if (c == 0 && !BINJ) mbar_wait(bar_hfull + 8 * buf, (it >> 1) & 1);
        if (c == 2) mbar_wait(bar_gfull, it & 1);
        const uint32_t a0_hi = bwd ? gbuf_u + m * 8192 : hbuf_u + buf * 8192;
        const uint32_t a0_lo = a0_hi + 4096;
        const uint32_t bt = btile_u + m * (LAT * 32);
        const uint32_t idesc_l0 = bwd ? idesc_l0_b : idesc_l0_f, idesc_h = bwd ? idesc_h_b : idesc_h_f;
        const uint32_t idesc_last = bwd ? idesc_last_b : idesc_last_f;
        // ---- chain layer 0: A from smem (K = 16), N = LAT --------------------------------------
        if (BINJ && !bwd) {
          region ^= 1;         // the epilogue warps write layer 0's activations into this region from global memory
        } else {
          const uint32_t d = tmem_base + region * LAT;
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sb = ring_u + stage * kStageBytes;
          if (elect_one()) {
            if (!bwd) mma_ss(d, smem_desc(sel_u, 2048, 128), smem_desc(bt, LAT * 16, 128), idesc_l0, 0);
            const uint64_t a_hi = smem_desc(a0_hi, 2048, 128), a_lo = smem_desc(a0_lo, 2048, 128);
            const uint64_t b_hi = smem_desc(sb, LAT * 16, 128);
            const uint64_t b_lo = smem_desc(sb + G::L0_MAT_BYTES, LAT * 16, 128);
            mma_ss(d, a_hi, b_hi, idesc_l0, bwd ? 0u : 1u);
            mma_ss(d, a_lo, b_hi, idesc_l0, 1);
            mma_ss(d, a_hi, b_lo, idesc_l0, 1);
            tc_commit(bar_empty + 8 * stage);
            tc_commit(bar_acc);
            tc_commit(bar_acc + 8);
            if (c == 1) tc_commit(bar_hempty + 8 * buf);
          }
          __syncwarp();
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
          region ^= 1;
        }
        // ---- hidden chain layers: A from TMEM ----------------------------------------------------
        // (two N = 128 column halves with separate commits: unlike the forward kernel, ONE N = 256 MMA per k-step was
        //  measured slower here -- 815 vs 793 us per launch -- because this epilogue also writes the tile images and
        //  needs the overlap of half 0's epilogue with half 1's MMAs more than it needs the cheaper instruction stream)
        for (int l = 1; l < K - 1; ++l) {
          const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
          const uint32_t out_col = tmem_base + region * LAT;
          const uint32_t sel_l = sel_u + l * 4096;
          uint32_t waited = 0;
#pragma unroll 1
          for (int ph = 0; ph < kNS; ++ph) {
            const uint32_t d = out_col + ph * G::NH;
#pragma unroll 1
            for (int kc = 0; kc < G::NKC; ++kc) {
              wait_groups(waited, (kc * G::KC) >> 6, (kc * G::KC + G::KC - 1) >> 6);
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              const uint32_t sb = ring_u + stage * kStageBytes;
              const uint32_t a0 = in_col + a_col(kc * G::KC);
              if (elect_one()) {
                if (kc == 0 && !bwd)
                  mma_ss(d, smem_desc(sel_l, 2048, 128), smem_desc(bt + ph * G::NH * 16, LAT * 16, 128), idesc_h, 0);
                const uint64_t b_hi0 = smem_desc(sb, G::NH * 16, 128);
                const uint64_t b_lo0 = smem_desc(sb + G::MAT_BYTES, G::NH * 16, 128);
#pragma unroll
                for (int ks = 0; ks < G::KC / 16; ++ks) {
                  const uint64_t koff = (uint64_t)((ks * 2 * (G::NH * 16)) >> 4);
                  const uint32_t a_hi = a0 + a_col(ks * 16);
                  mma_ts(d, a_hi, b_hi0 + koff, idesc_h, (bwd && kc == 0 && ks == 0) ? 0u : 1u);
                  mma_ts(d, a_hi + 16, b_hi0 + koff, idesc_h, 1);
                  mma_ts(d, a_hi, b_lo0 + koff, idesc_h, 1);
                }
                tc_commit(bar_empty + 8 * stage);
                if (kc == G::NKC - 1) tc_commit(bar_acc + 8 * ph);
              }
              __syncwarp();
              if (++stage == kBStages) { stage = 0; phase ^= 1; }
            }
          }
          region ^= 1;
        }
        // ---- last chain layer: N = 16, K = LAT ----------------------------------------------------
        // BINJ backward chains stop at delta_0: no MMA here, but the issuer still waits until the epilogue warps have
        // drained the delta_0 accumulator (they arrive on a_ready as if they had converted it).  Without that wait the
        // next chain's layer 0 -- which needs nothing from the epilogue -- could commit acc_full a second time before the
        // epilogue has seen the first commit (a barrier two phases ahead of its waiter: deadlock), and could overwrite
        // the accumulator region still being read.  The region is not toggled.
        if (BINJ && bwd) {
          uint32_t waited = 0;
          wait_groups(waited, 0, G::NA - 1);
        } else {
          const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
          const uint32_t d = tmem_base + region * LAT;
          uint32_t waited = 0;
          wait_groups(waited, 0, G::NA - 1);
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sb = ring_u + stage * kStageBytes;
          if (elect_one()) {
            const uint64_t b_hi0 = smem_desc(sb, kNOut * 16, 128);
            const uint64_t b_lo0 = smem_desc(sb + G::LAST_MAT_BYTES, kNOut * 16, 128);
#pragma unroll
            for (int ks = 0; ks < LAT / 16; ++ks) {
              const uint64_t koff = (uint64_t)((ks * 2 * (kNOut * 16)) >> 4);
              const uint32_t a_hi = in_col + a_col(ks * 16);
              mma_ts(d, a_hi, b_hi0 + koff, idesc_last, ks ? 1u : 0u);
              mma_ts(d, a_hi + 16, b_hi0 + koff, idesc_last, 1);
              mma_ts(d, a_hi, b_lo0 + koff, idesc_last, 1);
            }
            tc_commit(bar_empty + 8 * stage);
            tc_commit(bar_acc_last);
          }
          __syncwarp();
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
          region ^= 1;
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===== epilogue warps ==========================================================================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0;
    int region = 0;
    const int hp4 = p.HP >> 2;
    constexpr int NCH = G::GPH;                     // 32-column chunks per accumulator half per warp
    // act' masks: written by the forward chains' epilogue, read back by the SAME thread in the backward chains.
    // Up to kMaskLayers hidden layers they live in shared memory (one word per 32-column chunk); deeper MLPs
    // re-read the sign bits from the activation image in HBM.
    const bool smem_mask = (K - 1) <= kMaskLayers;
    uint32_t* my_mask = mask_s + (tid - 64);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      float sv[kNOut], tv[kNOut], gh[kNOut];
      const int64_t node = (int64_t)tile * kTileM + row;
      const bool valid = node < p.n_nodes;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const bool bwd = c >= 2;
        const int m = c & 1;
        for (int l = 0; l < K - 1; ++l) {
          // forward chain layer l produces a_l; backward chain layer l produces delta_{K-2-l}
          const int li = bwd ? K - 2 - l : l;
          uint16_t* out_img = (bwd ? p.dlt_img : p.act_img) + ((size_t)m * (K - 1) + li) * layer_stride +
                              (size_t)tile * 2 * img_elems;
          const uint16_t* mask_img = p.act_img + ((size_t)m * (K - 1) + li) * layer_stride + (size_t)tile * 2 * img_elems;
#pragma unroll
          for (int ph = 0; ph < kNS; ++ph) {
            uint32_t pos[NCH];                       // bit j: element 2j positive, bit 16+j: element 2j+1 positive
            if (bwd && smem_mask) {
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch) pos[ch] = my_mask[(((m * kMaskLayers + li) * kNS + ph) * 2 + ch) * kEpiThreads];
            } else if (bwd) {
#pragma unroll
              for (int ch = 0; ch < NCH; ++ch) {
                const int col0 = ph * G::NH + grp * 32 + ch * 64;
                const uint4* mp = reinterpret_cast<const uint4*>(mask_img + img_off(LAT, col0 >> 3, row));
                uint4 mk[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) mk[g] = mp[g * 8];          // 64 elements = 8 uint4 per feature group
                const uint32_t* mw = reinterpret_cast<const uint32_t*>(&mk[0]);
                uint32_t bits = 0;          // bit j: element 2j positive, bit 16+j: element 2j+1 positive
                if (ACT == GNF_ACT_LEAKY_RELU) {
                  // leaky_relu keeps the sign of the pre-activation: positive <=> sign bit clear (a == 0 is measure zero)
#pragma unroll
                  for (int j = 0; j < 16; ++j) bits |= ((mw[j] >> 15) & 0x10001u) << j;
                  bits = ~bits;
                } else {
                  // relu output is >= 0: positive <=> non-zero
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const uint32_t w = mw[j];
                    bits |= ((w & 0xFFFFu) ? 1u : 0u) << j;
                    bits |= ((w >> 16) ? 1u : 0u) << (16 + j);
                  }
                }
                pos[ch] = bits;
              }
            }
            // BINJ: the forward chains' layer 0 comes from global memory (pre-activations of k_linear_tc), the
            // backward chains end at delta_0 (kept in fp32 for g_h = delta_0 W_0^T, not converted back into TMEM)
            const bool from_global = BINJ && !bwd && l == 0;
            const bool stop_here = BINJ && bwd && l == K - 2;
            if (!from_global) {
              mbar_wait(smem_u32(&bars->acc_full[ph]), (acc_par >> ph) & 1u);
              acc_par ^= 1u << ph;
              tc_fence_after();
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              const int col0 = ph * G::NH + grp * 32 + ch * 64;
              const uint32_t t0 = lane_base + region * LAT + col0;
              uint32_t v[32];
              if (from_global) {
                const float4* src = reinterpret_cast<const float4*>(p.pre0[m] + node * LAT + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 x = valid ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                  v[4 * j] = __float_as_uint(x.x); v[4 * j + 1] = __float_as_uint(x.y);
                  v[4 * j + 2] = __float_as_uint(x.z); v[4 * j + 3] = __float_as_uint(x.w);
                }
              } else {
                tmem_ld32(t0, v);
                tmem_wait_ld();
              }
              float d[32];
              if (!bwd) {
#pragma unroll
                for (int j = 0; j < 32; ++j) d[j] = act_f<ACT>(__uint_as_float(v[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  d[2 * j] = __uint_as_float(v[2 * j]) * dact_f<ACT>((pos[ch] >> j) & 1u);
                  d[2 * j + 1] = __uint_as_float(v[2 * j + 1]) * dact_f<ACT>((pos[ch] >> (16 + j)) & 1u);
                }
              }
              uint32_t hi[16], lo[16];
              if (bwd) {
#pragma unroll
                for (int j = 0; j < 16; ++j) split_pair<kBF>(d[2 * j], d[2 * j + 1], hi[j], lo[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) split_pair<kFB>(d[2 * j], d[2 * j + 1], hi[j], lo[j]);
              }
              if (!stop_here) {
                tmem_st16(t0, hi);
                tmem_st16(t0 + 16, lo);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->a_ready[NCH == 2 ? ph * 2 + ch : ph]));
              } else {
                tc_fence_before();        // accumulator chunk read (tcgen05.wait::ld above): release it to the issuer
                mbar_arrive(smem_u32(&bars->a_ready[NCH == 2 ? ph * 2 + ch : ph]));
                if (valid) {
                  float4* dst = reinterpret_cast<float4*>(p.d0[m] + node * LAT + col0);
#pragma unroll
                  for (int j = 0; j < 8; ++j) dst[j] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
                }
              }
              if (!bwd && smem_mask) {       // sign bits of a_l for the backward chains (fp16 and bf16 share bit 15)
                uint32_t bits = 0;
                if (ACT == GNF_ACT_LEAKY_RELU) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) bits |= ((hi[j] >> 15) & 0x10001u) << j;
                  bits = ~bits;
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    bits |= ((hi[j] & 0xFFFFu) ? 1u : 0u) << j;
                    bits |= ((hi[j] >> 16) ? 1u : 0u) << (16 + j);
                  }
                }
                my_mask[(((m * kMaskLayers + li) * kNS + ph) * 2 + ch) * kEpiThreads] = bits;
              }
              // image for the weight-gradient GEMM: always bf16 (tcgen05 kind::f16 needs one operand format)
              if (F16F && !bwd) {
#pragma unroll
                for (int j = 0; j < 16; ++j) split_pair<kBF>(d[2 * j], d[2 * j + 1], hi[j], lo[j]);
              }
              uint4* oh = reinterpret_cast<uint4*>(out_img + img_off(LAT, col0 >> 3, row));
              uint4* ol = reinterpret_cast<uint4*>(out_img + img_elems + img_off(LAT, col0 >> 3, row));
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                oh[g * 8] = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
                if (p.write_lo) ol[g * 8] = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
              }
            }
          }
          region ^= 1;
        }
        // ---- last chain layer: s / t (forward chains) or g_h (backward chains) ----------------------
        if (!(BINJ && bwd)) {
          mbar_wait(smem_u32(&bars->acc_last), (acc_par >> 8) & 1u);
          acc_par ^= 1u << 8;
          tc_fence_after();
          if (grp == 0) {
            uint32_t v[16];
            tmem_ld16(lane_base + region * LAT, v);
            tmem_wait_ld();
            if (c == 0) {
#pragma unroll
              for (int j = 0; j < kNOut; ++j) sv[j] = __uint_as_float(v[j]) + blast[j];
            } else if (c == 1) {
#pragma unroll
              for (int j = 0; j < kNOut; ++j) tv[j] = __uint_as_float(v[j]) + blast[kNOut + j];
            } else if (c == 2) {
#pragma unroll
              for (int j = 0; j < kNOut; ++j) gh[j] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < kNOut; ++j) gh[j] += __uint_as_float(v[j]);
            }
          }
          region ^= 1;
        }
        if (c == 1 && grp == 0) {
          // ---- undo the affine update, top gradients (gnn.py:359 / :372; run_grevnet.py:292-296) -----
          float gs[kNOut], gt[kNOut];
#pragma unroll
          for (int j = 0; j < kNOut; ++j) { gs[j] = 0.f; gt[j] = 0.f; }
          if (valid) {
            float* xrow = p.xb + node * p.HP;
            float* grow = p.gxb + node * p.HP;
#pragma unroll
            for (int g4 = 0; g4 < kNOut / 4; ++g4) {
              if (g4 < hp4) {
                float4 x4 = *reinterpret_cast<const float4*>(xrow + g4 * 4);
                float4 g4v = *reinterpret_cast<const float4*>(grow + g4 * 4);
                float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                float gv[4] = {g4v.x, g4v.y, g4v.z, g4v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int f = g4 * 4 + j;
                  if (f < p.H) {
                    const float s = sv[f], t = tv[f];
                    const float es = expf(s);
                    const float x = __fmul_rn(__fsub_rn(xv[j], t), expf(-s));
                    gs[f] = gv[j] * x * es - p.scale;
                    gt[f] = gv[j];
                    xv[j] = x;
                    gv[j] = gv[j] * es;
                  }
                }
                *reinterpret_cast<float4*>(xrow + g4 * 4) = make_float4(xv[0], xv[1], xv[2], xv[3]);
                *reinterpret_cast<float4*>(grow + g4 * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
              }
            }
          }
          uint32_t hi[2][8], lo[2][8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            split_pair<kBF>(gs[2 * j], gs[2 * j + 1], hi[0][j], lo[0][j]);
            split_pair<kBF>(gt[2 * j], gt[2 * j + 1], hi[1][j], lo[1][j]);
          }
#pragma unroll
          for (int mm = 0; mm < 2; ++mm) {
            uint8_t* gb = gbuf + mm * 8192;          // K-major A tile of the backward chain's layer 0
            *reinterpret_cast<uint4*>(gb + row * 16) = make_uint4(hi[mm][0], hi[mm][1], hi[mm][2], hi[mm][3]);
            *reinterpret_cast<uint4*>(gb + 2048 + row * 16) = make_uint4(hi[mm][4], hi[mm][5], hi[mm][6], hi[mm][7]);
            *reinterpret_cast<uint4*>(gb + 4096 + row * 16) = make_uint4(lo[mm][0], lo[mm][1], lo[mm][2], lo[mm][3]);
            *reinterpret_cast<uint4*>(gb + 4096 + 2048 + row * 16) = make_uint4(lo[mm][4], lo[mm][5], lo[mm][6], lo[mm][7]);
            uint16_t* gi = p.g_img + ((size_t)mm * p.n_tiles + tile) * 2 * (16 * 128);   // MN-major image, F = 16
            uint4* oh = reinterpret_cast<uint4*>(gi + img_off(16, 0, row));
            uint4* ol = reinterpret_cast<uint4*>(gi + 16 * 128 + img_off(16, 0, row));
            oh[0] = make_uint4(hi[mm][0], hi[mm][1], hi[mm][2], hi[mm][3]);
            oh[8] = make_uint4(hi[mm][4], hi[mm][5], hi[mm][6], hi[mm][7]);
            ol[0] = make_uint4(lo[mm][0], lo[mm][1], lo[mm][2], lo[mm][3]);
            ol[8] = make_uint4(lo[mm][4], lo[mm][5], lo[mm][6], lo[mm][7]);
          }
          fence_proxy_async();
          mbar_arrive(smem_u32(&bars->g_full));
        }
        if (!BINJ && c == 3 && grp == 0 && valid) {
          float* go = p.gh + node * kK0;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            *reinterpret_cast<float4*>(go + g4 * 4) = make_float4(gh[g4 * 4], gh[g4 * 4 + 1], gh[g4 * 4 + 2], gh[g4 * 4 + 3]);
        }
      }
    }
  } else {
    // ===== gather warps: a3 + a4 + a5, as the forward kernel, plus the h image ==================
    const int row = tid - (kThreads - kGatherThreads);
    float* my = hstage + row * 17;
    const int hp4 = p.HP >> 2;
    int it = 0;
    for (int tile = blockIdx.x; !BINJ && tile < p.n_tiles; tile += gridDim.x, ++it) {   // BINJ: the MLP input never enters
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
        for (; e < end; ++e) {
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
      for (int j = 0; j < 8; ++j) split_pair<kBF>(my[2 * j], my[2 * j + 1], hi[j], lo[j]);
      {
        uint16_t* hi_img = p.h_img + (size_t)tile * 2 * (16 * 128);
        uint4* oh = reinterpret_cast<uint4*>(hi_img + img_off(16, 0, row));
        uint4* ol = reinterpret_cast<uint4*>(hi_img + 16 * 128 + img_off(16, 0, row));
        oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        oh[8] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        ol[8] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      if (F16F) {
#pragma unroll
        for (int j = 0; j < 8; ++j) split_pair<kFB>(my[2 * j], my[2 * j + 1], hi[j], lo[j]);
      }
      mbar_wait(smem_u32(&bars->h_empty[buf]), ((it >> 1) & 1) ^ 1);
      uint8_t* hb = hbuf + buf * 8192;
      *reinterpret_cast<uint4*>(hb + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(hb + 2048 + row * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(hb + 4096 + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(hb + 4096 + 2048 + row * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->h_full[buf]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// =================================================================================================
// Weight-gradient GEMM: part[split][M = a_feats][N = b_feats] = sum over the split's tiles of
// A_img^T-view [a_feats x nodes] * B_img [nodes x b_feats]   (both MN-major, K = node index)
// =================================================================================================
constexpr int kDwStages = 3;
constexpr int kDwStageBytes = 65536;
constexpr int kDwThreads = 192;        // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int kDwMaxPairs = 2 * kMaxLayers;

struct DwPair {
  const uint16_t* A;       // image array of the M-side operand: [tiles][2][a_feats*128]
  const uint16_t* B;       // image array of the N-side operand: [tiles][2][b_feats*128]
  float* part;             // [n_splits][a_feats][b_feats]
  int a_feats, b_feats;
  int first_cta, n_splits;
  int a_bf16, b_bf16;      // element type of each operand's images (0 = fp16)
  int db_mode;             // bias gradient = column sums over the nodes of: 0 nothing, 1 the B operand, 2 the A operand
  float* db_part;          // [n_splits][feats of that operand]
};
struct DwParams {
  DwPair pair[kDwMaxPairs];
  int n_pairs, n_tiles;
};

struct __align__(8) DwBarriers {
  uint64_t full[kDwStages], empty[kDwStages], acc_full;
};

// 16-bit x 16-bit -> f32, both operands MN-major; the two operand formats are independent fields
__host__ __device__ constexpr uint32_t make_idesc_mn(int n, bool a_bf16, bool b_bf16) {
  return (1u << 4) | ((a_bf16 ? 1u : 0u) << 7) | ((b_bf16 ? 1u : 0u) << 10) | (1u << 15) | (1u << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

template <int PARTS>   // 2: hi/lo split, 3 MMAs per product; 1: single bf16 pass
__global__ void __launch_bounds__(kDwThreads, 1) k_dw_tc(const DwParams p) {
  constexpr int KS = PARTS == 2 ? 2 : 4;          // 16-node k-steps per stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  DwBarriers* bars = (DwBarriers*)(smem + kDwStages * kDwStageBytes);
  uint32_t* tmem_slot = (uint32_t*)(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int pi = 0;
  for (int i = 1; i < p.n_pairs; ++i)
    if ((int)blockIdx.x >= p.pair[i].first_cta) pi = i;
  const DwPair& pr = p.pair[pi];
  const int split = blockIdx.x - pr.first_cta;
  if (split >= pr.n_splits) return;               // whole CTA exits together (grid may be padded)
  const int t0 = (int)((int64_t)split * p.n_tiles / pr.n_splits);
  const int t1 = (int)((int64_t)(split + 1) * p.n_tiles / pr.n_splits);
  const int FA = pr.a_feats, FB = pr.b_feats;
  const int MH = FA / 128;
  const uint32_t a_step = 2u * FA * 16u, b_step = 2u * FB * 16u;      // bytes per 16-node k-step per part
  const uint32_t a_part = KS * a_step, b_part = KS * b_step;          // bytes per stage per part
  const uint32_t stage_bytes = PARTS * (a_part + b_part);

  if (tid == 0) {
    for (int i = 0; i < kDwStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), pr.db_mode ? 5 : 1);     // MMA commit (+ the 4 column-sum warps)
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t0; tile < t1; ++tile) {
        const uint8_t* a_img = (const uint8_t*)(pr.A + (size_t)tile * 2 * FA * 128);
        const uint8_t* b_img = (const uint8_t*)(pr.B + (size_t)tile * 2 * FB * 128);
        for (int s = 0; s < 8 / KS; ++s) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&bars->full[stage]);
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t dst = smem_u32(smem + stage * kDwStageBytes);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) {
            bulk_g2s(dst + part * a_part, a_img + (size_t)part * FA * 256 + (size_t)s * a_part, a_part, fb);
            bulk_g2s(dst + PARTS * a_part + part * b_part, b_img + (size_t)part * FB * 256 + (size_t)s * b_part, b_part, fb);
          }
          if (++stage == kDwStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_mn(FB, pr.a_bf16 != 0, pr.b_bf16 != 0);
    // MN-major no-swizzle: SBO = stride between 8-feature groups, LBO = stride between 8-node groups
    const uint32_t a_lbo = FA * 16u, a_sbo = 128u, b_lbo = FB * 16u, b_sbo = 128u;
    uint32_t stage = 0, phase = 0;
    bool first = true;
    for (int tile = t0; tile < t1; ++tile) {
      for (int s = 0; s < 8 / KS; ++s) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kDwStageBytes);
        const uint32_t sb = sa + PARTS * a_part;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t b_hi = smem_desc(sb + ks * b_step, b_lbo, b_sbo);
            const uint64_t b_lo = smem_desc(sb + b_part + ks * b_step, b_lbo, b_sbo);
            for (int mh = 0; mh < MH; ++mh) {
              const uint32_t d = tmem_base + mh * 256;      // (FB <= 256; a fixed stride keeps every FB aligned)
              const uint64_t a_hi = smem_desc(sa + ks * a_step + mh * 2048, a_lbo, a_sbo);
              mma_ss(d, a_hi, b_hi, idesc, (first && ks == 0) ? 0u : 1u);
              if (PARTS == 2) {
                const uint64_t a_lo = smem_desc(sa + a_part + ks * a_step + mh * 2048, a_lbo, a_sbo);
                mma_ss(d, a_lo, b_hi, idesc, 1);
                mma_ss(d, a_hi, b_lo, idesc, 1);
              }
            }
          }
          tc_commit(smem_u32(&bars->empty[stage]));
        }
        __syncwarp();
        first = false;
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
    }
    if (elect_one()) tc_commit(smem_u32(&bars->acc_full));
    __syncwarp();
  } else {
    if (pr.db_mode) {
      // ---- bias gradient: column sums of one operand over the nodes -------------------------------------------
      // The warps follow the ring (wait on full[stage], arrive on empty[stage]) only to stay in step with the
      // producer: they read the SAME bytes the stage's bulk copy has just pulled through L2, but from the image in
      // global memory (read-only path), not from shared memory -- no generic-proxy reads of async-proxy writes.
      // MN-major image: per 8-node group, feature group fg is 8 nodes x 16 B; lane i reads node i&7 of feature
      // group fg0 + (i>>3): a warp-wide 16-byte load covers 4 feature groups x 8 nodes = 512 contiguous bytes.
      const int cw = warp - 2;
      const int F = pr.db_mode == 1 ? FB : FA;
      const uint32_t part_bytes = pr.db_mode == 1 ? b_part : a_part;
      const uint16_t* op_img = pr.db_mode == 1 ? pr.B : pr.A;
      const int fg_per_warp = F >= 128 ? F / 32 : 2;
      const int nslots = fg_per_warp > 4 ? fg_per_warp / 4 : 1;
      const bool active = F >= 128 ? true : (cw == 0 && lane < 16);
      float acc[2][8];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
      uint32_t stage = 0, phase = 0;
      for (int tile = t0; tile < t1; ++tile) {
        for (int s = 0; s < 8 / KS; ++s) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          if (active) {
            const uint8_t* base = (const uint8_t*)(op_img + (size_t)tile * 2 * F * 128) + (size_t)s * part_bytes;
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              for (int ng = 0; ng < KS * 2; ++ng) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  if (j < nslots) {
                    const int fg = (F >= 128 ? cw * fg_per_warp + 4 * j : 0) + (lane >> 3);
                    const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)part * F * 256 + (size_t)ng * F * 16 +
                                                                          fg * 128 + (lane & 7) * 16));
                    const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      acc[j][2 * e] += __uint_as_float(ww[e] << 16);
                      acc[j][2 * e + 1] += __uint_as_float(ww[e] & 0xFFFF0000u);
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->empty[stage]));
          if (++stage == kDwStages) { stage = 0; phase ^= 1; }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = acc[j][e];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          acc[j][e] = v;
        }
      if (active && (lane & 7) == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (j < nslots) {
            const int fg = (F >= 128 ? cw * fg_per_warp + 4 * j : 0) + (lane >> 3);
#pragma unroll
            for (int e = 0; e < 8; ++e) pr.db_part[(size_t)split * F + fg * 8 + e] = acc[j][e];
          }
      }
    }
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    float* out = pr.part + (size_t)split * FA * FB;
    for (int mh = 0; mh < MH; ++mh) {
      float* orow = out + (size_t)(mh * 128 + row) * FB;
      if (FB >= 32) {
        for (int c = 0; c < FB; c += 32) {
          uint32_t v[32];
          tmem_ld32(lane_base + mh * 256 + c, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(orow + c + 4 * j) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
        }
      } else {
        uint32_t v[16];
        tmem_ld16(lane_base + mh * 256, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(orow + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// grads += fixed-order sum of the split partials; transposed = partial is [out][in]
struct DwReducePair {
  const float* part;
  float* grad;            // W block of this layer inside the flat gradient ([in][out] row-major)
  int n_splits, fa, fb, in_dim, out_dim, transposed;
  const float* db_part;   // [n_splits][db_feats] column sums (bias gradient), or null
  float* grad_b;
  int db_feats;
};
struct DwReduceParams {
  DwReducePair pair[kDwMaxPairs];
};
__global__ void k_dw_reduce(const DwReduceParams p) {
  const DwReducePair& r = p.pair[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && r.db_part && (int)threadIdx.x < r.out_dim) {     // bias gradient (out_dim <= 256 = blockDim)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int z0 = 0; z0 < r.n_splits; z0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (z0 + u < r.n_splits) acc[u] += r.db_part[(size_t)(z0 + u) * r.db_feats + threadIdx.x];
    }
    r.grad_b[threadIdx.x] += (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  if (i >= r.fa * r.fb) return;
  const int a = i / r.fb, b = i - a * r.fb;
  const int in = r.transposed ? b : a, out = r.transposed ? a : b;
  if (in >= r.in_dim || out >= r.out_dim) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int z0 = 0; z0 < r.n_splits; z0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (z0 + u < r.n_splits) acc[u] += r.part[(size_t)(z0 + u) * r.fa * r.fb + i];
  }
  r.grad[(size_t)in * r.out_dim + out] += (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ---- debug / unit-test helper: fp32 [n, F] row-major -> bf16 hi/lo tile images ----------------
// (also the product path of inject flows: the MLP input h [n, ld] -> the B operand of dW_0 = h^T delta_0)
__global__ void k_make_images(const float* __restrict__ x, int64_t n, int ld, int fvalid, int F, int n_tiles,
                              uint16_t* __restrict__ img) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_tiles * 128 * F) return;
  const int64_t node = i / F;
  const int f = (int)(i - node * F);
  const float v = (node < n && f < fvalid) ? x[node * ld + f] : 0.f;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  const int tile = (int)(node >> 7), r = (int)(node & 127);
  uint16_t* base = img + (size_t)tile * 2 * F * 128;
  const size_t off = img_off(F, f >> 3, r) + (f & 7);
  base[off] = *reinterpret_cast<const uint16_t*>(&h);
  base[(size_t)F * 128 + off] = *reinterpret_cast<const uint16_t*>(&l);
}

int launch_dw(const DwParams& p, int grid, int parts, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)kDwStages * kDwStageBytes + sizeof(DwBarriers) + 64;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    GNF_CUDA(cudaFuncSetAttribute(k_dw_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GNF_CUDA(cudaFuncSetAttribute(k_dw_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (parts == 2) k_dw_tc<2><<<grid, kDwThreads, smem, stream>>>(p);
  else k_dw_tc<1><<<grid, kDwThreads, smem, stream>>>(p);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

template <int LAT, int ACT, bool F16F, bool BINJ>
int launch_chain_t(const BwdParams& p, int grid, cudaStream_t stream) {
  auto kern = k_bwd_chain<LAT, ACT, F16F, BINJ>;
  static bool configured[kMaxDevices] = {};
  const size_t smem = bwd_smem_bytes<LAT>();
  if (first_use_on_device(configured))
    GNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreads, smem, stream>>>(p);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

struct BwdTcWs {
  float *x0, *x1, *g0, *g1, *gh, *dw_part;
  uint16_t *act_img, *dlt_img, *h_img, *g_img;
  float *pre0[2], *d0[2];        // inject flows: layer-0 pre-activations / delta_0 of the s and t MLP, [n, L] fp32
  size_t dw_part_floats;
  size_t bytes;
};

// feature count of the MLP-input image (B operand of dW_0): 16 for the fully fused shapes; inject flows pad to 32
// (k_dw_tc drains its accumulator 32 columns at a time)
int dw_in0(const Flow& f) { return f.tc_ok ? kK0 : (f.in_dim + 31) / 32 * 32; }

// split plan of the weight-gradient GEMM: CTAs per pair proportional to the bytes each pair streams
struct DwPlan {
  int n_pairs;
  int splits[kDwMaxPairs], first[kDwMaxPairs];
  int grid;
  size_t part_off[kDwMaxPairs];      // floats
  size_t db_off[kDwMaxPairs];        // floats: [splits][L] column-sum partials of the pair
  size_t part_floats;
};

DwPlan plan_dw(const Flow& f, int n_tiles) {
  DwPlan pl{};
  const int K = f.K, L = f.L;
  pl.n_pairs = 2 * K;
  double cost[kDwMaxPairs];
  double total = 0;
  const int in0 = dw_in0(f);
  for (int m = 0; m < 2; ++m)
    for (int l = 0; l < K; ++l) {
      const bool hidden = l > 0 && l < K - 1;
      cost[m * K + l] = hidden ? 2.0 * L : (double)L + (l == 0 ? in0 : 16);
      total += cost[m * K + l];
    }
  const int sms = num_sms();
  int used = 0;
  for (int i = 0; i < pl.n_pairs; ++i) {
    int s = (int)(cost[i] / total * sms);
    if (s < 1) s = 1;
    if (s > n_tiles) s = n_tiles;
    if (s < 1) s = 1;
    pl.splits[i] = s;
    used += s;
  }
  // hand out what rounding left over to the most expensive pairs
  for (int i = 0; used < sms && i < 4 * pl.n_pairs; ++i) {
    const int j = i % pl.n_pairs;
    const bool hidden = (j % K) > 0 && (j % K) < K - 1;
    if ((hidden || K == 2) && pl.splits[j] < n_tiles) { pl.splits[j]++; used++; }
  }
  size_t off = 0;
  int first = 0;
  for (int i = 0; i < pl.n_pairs; ++i) {
    const int l = i % K;
    const int fb = (l > 0 && l < K - 1) ? L : (l == 0 ? in0 : 16);
    pl.first[i] = first;
    first += pl.splits[i];
    pl.part_off[i] = off;
    off += (size_t)pl.splits[i] * L * fb;
    pl.db_off[i] = off;
    off += (size_t)pl.splits[i] * L;
  }
  pl.grid = first;
  pl.part_floats = off;
  return pl;
}

BwdTcWs carve_bwd_tc(const Flow& f, int64_t n, void* base) {
  BwdTcWs w{};
  uint8_t* p = (uint8_t*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base ? (void*)(p + off) : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const size_t tiles = (size_t)ceil_div((int64_t)nn, kTileM);
  const size_t img = (size_t)f.L * 128 * 2;          // elements, both parts
  w.x0 = (float*)take(nn * f.HP * 4);
  w.x1 = (float*)take(nn * f.HP * 4);
  w.g0 = (float*)take(nn * f.HP * 4);
  w.g1 = (float*)take(nn * f.HP * 4);
  w.gh = (float*)take(tiles * 128 * kK0 * 4);
  const DwPlan pl = plan_dw(f, (int)tiles);
  w.dw_part_floats = pl.part_floats;
  w.dw_part = (float*)take(pl.part_floats * 4);
  w.act_img = (uint16_t*)take((size_t)2 * (f.K - 1) * tiles * img * 2);
  w.dlt_img = (uint16_t*)take((size_t)2 * (f.K - 1) * tiles * img * 2);
  w.h_img = (uint16_t*)take((f.tc_ok ? 1 : 2) * tiles * 2 * (size_t)dw_in0(f) * 128 * 2);   // inject: one per MLP
  w.g_img = (uint16_t*)take((size_t)2 * tiles * 2 * 16 * 128 * 2);
  if (!f.tc_ok)
    for (int m = 0; m < 2; ++m) {
      w.pre0[m] = (float*)take(nn * f.L * 4);
      w.d0[m] = (float*)take(nn * f.L * 4);
    }
  w.bytes = off;
  return w;
}

}  // namespace

bool tc_bwd_supported(const Flow& f) { return f.tc_ok && f.wtcT != nullptr && f.wtcB[0] != nullptr && f.K >= 2; }

size_t tc_bwd_workspace(const Flow& f, int64_t n) { return carve_bwd_tc(f, n, nullptr).bytes; }

// byte offsets of the per-half-step buffers inside the workspace (bring-up / tests decode the images)
void tc_bwd_layout(const Flow& f, int64_t n, int64_t* out) {
  BwdTcWs w = carve_bwd_tc(f, n, (void*)4096);
  const uint8_t* b = (const uint8_t*)4096;
  out[0] = (const uint8_t*)w.act_img - b;
  out[1] = (const uint8_t*)w.dlt_img - b;
  out[2] = (const uint8_t*)w.h_img - b;
  out[3] = (const uint8_t*)w.g_img - b;
  out[4] = (const uint8_t*)w.gh - b;
  out[5] = (const uint8_t*)w.x0 - b;
  out[6] = (const uint8_t*)w.g0 - b;
  out[7] = (int64_t)w.bytes;
}

template <bool BINJ>
static int launch_chain(const Flow& f, const BwdParams& p, int grid, bool fwd_f16, cudaStream_t stream) {
  const bool leaky = f.d.act == GNF_ACT_LEAKY_RELU;
  if (f.L == 256) {
    if (fwd_f16) return leaky ? launch_chain_t<256, GNF_ACT_LEAKY_RELU, true, BINJ>(p, grid, stream)
                              : launch_chain_t<256, GNF_ACT_RELU, true, BINJ>(p, grid, stream);
    return leaky ? launch_chain_t<256, GNF_ACT_LEAKY_RELU, false, BINJ>(p, grid, stream)
                 : launch_chain_t<256, GNF_ACT_RELU, false, BINJ>(p, grid, stream);
  }
  if (fwd_f16) return leaky ? launch_chain_t<128, GNF_ACT_LEAKY_RELU, true, BINJ>(p, grid, stream)
                            : launch_chain_t<128, GNF_ACT_RELU, true, BINJ>(p, grid, stream);
  return leaky ? launch_chain_t<128, GNF_ACT_LEAKY_RELU, false, BINJ>(p, grid, stream)
               : launch_chain_t<128, GNF_ACT_RELU, false, BINJ>(p, grid, stream);
}

static BwdParams chain_params(const Flow& f, int ms, int mt, float* xb, float* gb, int64_t n, float scale,
                              const BwdTcWs& w, int dw_parts, bool fwd_f16) {
  BwdParams p{};
  p.xb = xb; p.gxb = gb;
  p.n_nodes = n; p.n_tiles = (int)ceil_div(n, kTileM);
  p.wf[0] = f.wtcB[fwd_f16 ? 0 : 1] + (size_t)ms * f.wtc_per_mlp;
  p.wf[1] = f.wtcB[fwd_f16 ? 0 : 1] + (size_t)mt * f.wtc_per_mlp;
  p.wb[0] = f.wtcT + (size_t)ms * f.wtc_per_mlp;
  p.wb[1] = f.wtcT + (size_t)mt * f.wtc_per_mlp;
  p.bias[0] = f.btc + (size_t)ms * f.K * 256;
  p.bias[1] = f.btc + (size_t)mt * f.K * 256;
  p.K = f.K; p.H = f.H; p.HP = f.HP;
  p.concat = f.d.block == GNF_BLOCK_CONCAT;
  p.mean = f.d.agg == GNF_AGG_MEAN;
  p.eps = f.d.eps; p.scale = scale;
  p.act_img = w.act_img; p.dlt_img = w.dlt_img; p.h_img = w.h_img; p.g_img = w.g_img;
  p.gh = w.gh;
  p.write_lo = dw_parts == 2;
  return p;
}

// weight (and bias) gradients of the s and t MLP from the tile images of one half step.  h_img[m]: image of MLP m's
// input with dw_in0(f) features (the fully fused shapes share one image)
static int dw_half(const Flow& f, int ms, int mt, int n_tiles, const BwdTcWs& w, const uint16_t* const h_img[2],
                   float* grads, int dw_parts, cudaStream_t stream) {
  const int K = f.K, L = f.L;
  const int in0 = dw_in0(f);
  const DwPlan pl = plan_dw(f, n_tiles);
  DwParams dp{};
  DwReduceParams rp{};
  dp.n_pairs = pl.n_pairs;
  dp.n_tiles = n_tiles;
  const size_t img = (size_t)L * 128 * 2;
  const size_t layer_stride = (size_t)n_tiles * img;
  for (int m = 0; m < 2; ++m) {
    const int mlp = m == 0 ? ms : mt;
    float* gm = grads + (int64_t)mlp * f.params_per_mlp;
    for (int l = 0; l < K; ++l) {
      const int i = m * K + l;
      DwPair& pr = dp.pair[i];
      DwReducePair& rr = rp.pair[i];
      pr.part = w.dw_part + pl.part_off[i];
      pr.first_cta = pl.first[i];
      pr.n_splits = pl.splits[i];
      pr.a_feats = L;
      if (l == 0) {                       // dW_0^T = delta_0^T h
        pr.A = w.dlt_img + ((size_t)m * (K - 1) + 0) * layer_stride;
        pr.B = h_img[m];
        pr.b_feats = in0;
        pr.a_bf16 = 1; pr.b_bf16 = 1;
        pr.db_mode = 2;                   // db_0 = column sums of delta_0 (the A operand)
      } else if (l == K - 1) {            // dW_{K-1} = a_{K-2}^T g_top
        pr.A = w.act_img + ((size_t)m * (K - 1) + (K - 2)) * layer_stride;
        pr.B = w.g_img + (size_t)m * n_tiles * 2 * 16 * 128;
        pr.b_feats = 16;
        pr.a_bf16 = 1; pr.b_bf16 = 1;
        pr.db_mode = 1;                   // db_{K-1} = column sums of the top gradient (the B operand)
      } else {                            // dW_l = a_{l-1}^T delta_l
        pr.A = w.act_img + ((size_t)m * (K - 1) + (l - 1)) * layer_stride;
        pr.B = w.dlt_img + ((size_t)m * (K - 1) + l) * layer_stride;
        pr.b_feats = L;
        pr.a_bf16 = 1; pr.b_bf16 = 1;
        pr.db_mode = 1;                   // db_l = column sums of delta_l (the B operand)
      }
      rr.part = pr.part;
      rr.grad = gm + f.flat_w_off[l];
      rr.n_splits = pr.n_splits;
      rr.fa = pr.a_feats;
      rr.fb = pr.b_feats;
      rr.in_dim = f.ins[l];
      rr.out_dim = f.outs[l];
      rr.transposed = l == 0;
      pr.db_part = w.dw_part + pl.db_off[i];
      rr.db_part = pr.db_part;
      rr.db_feats = pr.db_mode == 1 ? pr.b_feats : pr.a_feats;
      rr.grad_b = gm + f.flat_b_off[l];
    }
  }
  int rc = launch_dw(dp, pl.grid, dw_parts, stream);
  if (rc) return rc;
  dim3 rg((unsigned)ceil_div((int64_t)L * L, 256), (unsigned)pl.n_pairs);
  k_dw_reduce<<<rg, 256, 0, stream>>>(rp);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

// one reversed half step: (xa, xb', g_xa, g_xb') -> (xb, g_xa += ..., g_xb), grads += ...
static int bwd_half_tc(const Flow& f, int ms, int mt, const float* xa, float* xb, float* ga, float* gb, int64_t n,
                       const int32_t* rowptr, const int32_t* csr_senders, const int32_t* rowptr_s,
                       const int32_t* csr_receivers, float scale, float* grads, const BwdTcWs& w, int dw_parts,
                       bool fwd_f16, cudaStream_t stream) {
  const int n_tiles = (int)ceil_div(n, kTileM);
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  BwdParams p = chain_params(f, ms, mt, xb, gb, n, scale, w, dw_parts, fwd_f16);
  p.xa = xa;
  p.rowptr = rowptr; p.csr = csr_senders;
  int rc = launch_chain<false>(f, p, grid, fwd_f16, stream);
  if (rc) return rc;
  rc = bwd_agg_transpose(f, w.gh, kK0, rowptr_s, csr_receivers, rowptr, n, ga, stream);
  if (rc) return rc;
  const uint16_t* const h_img[2] = {w.h_img, w.h_img};
  return dw_half(f, ms, mt, n_tiles, w, h_img, grads, dw_parts, stream);
}

int tc_grevnet_backward(const Flow& f, const float* z, int64_t n, const int32_t* rowptr, const int32_t* csr_senders,
                        const int32_t* rowptr_s, const int32_t* csr_receivers, double loss_scale, float* grads,
                        float* x_out, void* ws, size_t ws_bytes, int dw_parts, int fwd_f16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE(tc_bwd_supported(f), GNF_EUNSUPPORTED, "tensor-core backward: unsupported flow shape");
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= carve_bwd_tc(f, n, nullptr).bytes, GNF_EWORKSPACE,
              "gnf_grevnet_backward: workspace too small or misaligned");
  BwdTcWs w = carve_bwd_tc(f, n, ws);
  const int D = f.d.node_embedding_dim;
  const float scale = (float)loss_scale;
  int rc = bwd_split_scale(z, n, D, f.H, f.HP, scale, w.x0, w.x1, w.g0, w.g1, stream);
  if (rc) return rc;
  for (int i = f.d.num_timesteps - 1; i >= 0; --i) {
    for (int half = 1; half >= 0; --half) {
      float* xa = half == 0 ? w.x0 : w.x1;
      float* xb = half == 0 ? w.x1 : w.x0;
      float* ga = half == 0 ? w.g0 : w.g1;
      float* gb = half == 0 ? w.g1 : w.g0;
      const int ms = f.mlp_index(0, half, i), mt = f.mlp_index(1, half, i);
      rc = bwd_half_tc(f, ms, mt, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_s, csr_receivers, scale, grads, w,
                       dw_parts, fwd_f16 != 0, stream);
      if (rc) return rc;
    }
  }
  if (x_out) {
    rc = bwd_merge(w.x0, w.x1, n, D, f.H, f.HP, x_out, stream);
    if (rc) return rc;
  }
  return GNF_OK;
}

int tc_half_backward(const Flow& f, int half, int step, const float* xa, float* xb, float* ga, float* gb, int64_t n,
                     const int32_t* rowptr, const int32_t* csr_senders, const int32_t* rowptr_s,
                     const int32_t* csr_receivers, double loss_scale, float* grads, void* ws, int dw_parts, int fwd_f16,
                     void* stream_) {
  BwdTcWs w = carve_bwd_tc(f, n, ws);
  const int ms = f.mlp_index(0, half, step), mt = f.mlp_index(1, half, step);
  return bwd_half_tc(f, ms, mt, xa, xb, ga, gb, n, rowptr, csr_senders, rowptr_s, csr_receivers, (float)loss_scale,
                     grads, w, dw_parts, fwd_f16 != 0, (cudaStream_t)stream_);
}

// ---- inject flows (MLP input wider than 16 columns) ---------------------------------------------------------------
// The caller (backward.cu) owns the input assembly and everything below layer 0; this side runs the MLP chains.
bool tc_bwd_inject_supported(const Flow& f) {
  return f.tc_inject && f.wtcT != nullptr && f.wtcB[0] != nullptr && f.K >= 2 && f.lin_off[4] >= 0 && f.lin_off[5] >= 0 &&
         (!f.attn || f.lin_off[6] >= 0) && dw_in0(f) <= 128;
}

size_t tc_bwd_inject_workspace(const Flow& f, int64_t n) { return carve_bwd_tc(f, n, nullptr).bytes; }

void tc_bwd_inject_buffers(const Flow& f, int64_t n, void* ws, float* pre0[2], float* d0[2]) {
  const BwdTcWs w = carve_bwd_tc(f, n, ws);
  for (int m = 0; m < 2; ++m) { pre0[m] = w.pre0[m]; d0[m] = w.d0[m]; }
}

// MLP chains of one reversed half step.  In: the layer-0 pre-activations (workspace buffers pre0[m], bias included),
// the MLP inputs hin[m] ([n, in_ld] fp32; the same pointer twice when s and t share their input), xb', g_xb'.
// Out: xb, g_xb, the workspace buffers d0[m] = dL/d(pre0[m]); grads += dW, db of every layer (layer 0 included).
int tc_half_backward_inject(const Flow& f, int ms, int mt, const float* const hin[2], int in_ld, float* xb, float* gb,
                            int64_t n, double loss_scale, float* grads, void* ws, int dw_parts, int fwd_f16,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const BwdTcWs w = carve_bwd_tc(f, n, ws);
  const int n_tiles = (int)ceil_div(n, kTileM);
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  const int in0 = dw_in0(f);
  const size_t h_elems = (size_t)n_tiles * 2 * in0 * 128;
  uint16_t* h_img[2] = {w.h_img, hin[1] == hin[0] ? w.h_img : w.h_img + h_elems};
  for (int m = 0; m < (hin[1] == hin[0] ? 1 : 2); ++m) {
    k_make_images<<<(unsigned)ceil_div((int64_t)n_tiles * 128 * in0, 256), 256, 0, stream>>>(hin[m], n, in_ld, in_ld, in0,
                                                                                             n_tiles, h_img[m]);
    GNF_LAUNCH_CHECK();
  }
  BwdParams p = chain_params(f, ms, mt, xb, gb, n, (float)loss_scale, w, dw_parts, fwd_f16 != 0);
  for (int m = 0; m < 2; ++m) { p.pre0[m] = w.pre0[m]; p.d0[m] = w.d0[m]; }
  int rc = launch_chain<true>(f, p, grid, fwd_f16 != 0, stream);
  if (rc) return rc;
  return dw_half(f, ms, mt, n_tiles, w, h_img, grads, dw_parts, stream);
}

// unit-test entry: out[fa][fb] = A^T B for fp32 A [n, fa], B [n, fb] through the image format + k_dw_tc
int tc_dw_gemm_test(const float* A, const float* B, int64_t n, int fa, int fb, int parts, int n_splits, float* out,
                    void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GNF_REQUIRE((fa == 128 || fa == 256) && (fb == 16 || fb == 128 || fb == 256) && (fa / 128) * fb <= 512, GNF_EINVAL,
              "gnf_debug_dw_gemm: fa in {128,256}, fb in {16,128,256}");
  GNF_REQUIRE(parts == 1 || parts == 2, GNF_EINVAL, "gnf_debug_dw_gemm: parts in {1,2}");
  const int n_tiles = (int)ceil_div(n, kTileM);
  GNF_REQUIRE(n_tiles >= 1 && n_splits >= 1 && n_splits <= n_tiles, GNF_EINVAL, "gnf_debug_dw_gemm: bad split");
  const size_t a_bytes = align_up((size_t)n_tiles * 2 * fa * 128 * 2, 1024), b_bytes = align_up((size_t)n_tiles * 2 * fb * 128 * 2, 1024);
  const size_t part_bytes = (size_t)n_splits * fa * fb * 4;
  GNF_REQUIRE(ws && ((uintptr_t)ws % 256) == 0 && ws_bytes >= a_bytes + b_bytes + part_bytes, GNF_EWORKSPACE,
              "gnf_debug_dw_gemm: workspace needs %zu bytes", a_bytes + b_bytes + part_bytes);
  uint16_t* ai = (uint16_t*)ws;
  uint16_t* bi = (uint16_t*)((uint8_t*)ws + a_bytes);
  float* part = (float*)((uint8_t*)ws + a_bytes + b_bytes);
  k_make_images<<<(unsigned)ceil_div((int64_t)n_tiles * 128 * fa, 256), 256, 0, stream>>>(A, n, fa, fa, fa, n_tiles, ai);
  GNF_LAUNCH_CHECK();
  k_make_images<<<(unsigned)ceil_div((int64_t)n_tiles * 128 * fb, 256), 256, 0, stream>>>(B, n, fb, fb, fb, n_tiles, bi);
  GNF_LAUNCH_CHECK();
  DwParams dp{};
  dp.n_pairs = 1;
  dp.n_tiles = n_tiles;
  dp.pair[0] = DwPair{ai, bi, part, fa, fb, 0, n_splits, 1, 1, 0, nullptr};
  int rc = launch_dw(dp, n_splits, parts, stream);
  if (rc) return rc;
  GNF_CUDA(cudaMemsetAsync(out, 0, (size_t)fa * fb * 4, stream));
  DwReduceParams rp{};
  rp.pair[0] = DwReducePair{part, out, n_splits, fa, fb, fa, fb, 0, nullptr, nullptr, 0};
  k_dw_reduce<<<dim3((unsigned)ceil_div((int64_t)fa * fb, 256), 1), 256, 0, stream>>>(rp);
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

}  // namespace gnf
