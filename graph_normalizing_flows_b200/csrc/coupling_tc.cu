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
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"

namespace gnf {
namespace {

using namespace tcx;
// One half coupling step: which planar half is read (xa) / updated (xb) and the two MLP images.
struct HalfDesc {
  const uint8_t* w[2];     // weight images of the s and t MLP
  const float* bias[2];    // [K][256]
  int32_t swap;            // 0: xa = x[0], xb = x[1] (gnn.py:320-323);  1: xa = x[1], xb = x[0] (gnn.py:335-338)
  int32_t pad;
};

struct TcParams {
  float* x[2];                 // planar halves [N, HP]
  const int32_t* rowptr;
  const int32_t* csr;
  int64_t n_nodes;
  int n_tiles;
  HalfDesc one;                // per-half-step launch: this launch's half step
  const HalfDesc* halves;      // persistent launch: device table of the n_halves half steps, in execution order
  int n_halves;
  unsigned int* grid_bar;      // persistent launch: {arrival count, generation} of the grid barrier between half steps
  int K, H, HP, concat, mean, act, inverse;
  float eps;
  double* partials;
  double* ldj_accum;           // forward: the LAST CTA to finish adds the fixed-order sum of the partials here
  unsigned int* counter;       // arrival counter of that hand-off (zero on entry, reset by the last CTA)
  int* range_flag;             // sticky: set when an fp16-split operand exceeded the fp16 range (fp16 modes only)
  const float* pre0[2];        // INJECT: layer-0 pre-activations (bias included) of the s and t MLP, [N, LAT] fp32
  unsigned long long* trace;   // optional timeline of CTA 0 (gnf_debug_set_trace); null in production
  unsigned long long* tstamp;  // optional {min start, max end} globaltimer of this launch (gnf_debug_kernel_timing)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// timeline entry: [63:56] event, [55:48] m*16+l, [47:40] ph*4+kc, [39:0] clock
struct Tracer {
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ void init(unsigned long long* base, int role, bool on) {
    buf = on && base ? base + role * 2048 : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void ev(int event, int ml, int pk) {
    if (buf && n < 2047) {
      buf[1 + n++] = ((unsigned long long)event << 56) | ((unsigned long long)(ml & 255) << 48) |
                     ((unsigned long long)(pk & 255) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull);
      buf[0] = n;
    }
  }
};

struct __align__(8) Barriers {
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full[kNS], a_ready[kMaxNA];      // (the forward kernel uses acc_full[0 .. kFwdNS))
  uint64_t h_full[2], h_empty[2];
  uint64_t x_full[2];  // bulk-staged sender rows of the next tiles (MODE 0): one cp.async.bulk per tile and buffer
  uint64_t role_bar;   // persistent launch: CTA-level barrier of the epilogue + gather warps at a half-step boundary (an
                       // mbarrier, not bar.sync: arrivals are per thread, no warp convergence required)
  uint64_t acc_last;   // last layer's own barrier: the next MLP's layer 0 commits acc_full without waiting for the
                       // epilogue, so sharing acc_full[0] could advance it two phases past a late waiter
};

// Node-feature staging (north star: TMA / shared-memory staging of node-feature tiles): graphs are contiguous node
// blocks, so the sender rows a 128-node tile aggregates are one compact row range [lo, hi] of the planar half; its
// 32-byte rows are 16-byte aligned, i.e. ONE legal cp.async.bulk.  Two buffers (the gather warps run a tile ahead).
constexpr int kXStageBytes = 8192;

template <int LAT>
constexpr size_t smem_bytes() {
  return 1024 /*align slack*/ + (size_t)kStages * kStageBytes + 2 * 2 * 4096 /*h tiles*/ +
         (kMaxLayers - 1) * 4096 /*selector A tiles*/ + 2 * LAT * 32 /*bias B tiles*/ +
         2 * kNOut * 4 /*last-layer bias*/ + kGatherThreads * 17 * 4 /*h staging*/ + sizeof(Barriers) + 128 +
         2 * kXStageBytes /*bulk-staged node-feature rows*/ + 384 /*xred + alignment of the stage buffers*/;
}

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  return ACT == GNF_ACT_LEAKY_RELU ? fmaxf(v, 0.2f * v) : fmaxf(v, 0.f);
}

// activation + hi/lo split of one 32-column accumulator chunk, written back in place
// (fp16 modes keep a running max |a| of what they split: above 65504 the hi part is inf where the reference's fp32
// arithmetic stays finite -- reported through TcParams::range_flag, never silently)
template <int NPROD, bool BF16, int ACT>
__device__ __forceinline__ void convert_chunk(uint32_t taddr, const uint32_t (&v)[32], float& amax) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = act_t<ACT>(__uint_as_float(v[2 * j])), b = act_t<ACT>(__uint_as_float(v[2 * j + 1]));
    if constexpr (!BF16) amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
    split_pair<BF16>(a, b, hi[j], lo[j]);
  }
  tmem_st16(taddr, hi);
  if (NPROD == 3) tmem_st16(taddr + 16, lo);
}

// Grid-wide barrier of the persistent launch (one CTA per SM, all co-resident: cooperative launch).  Executed by ONE
// thread per CTA after a CTA-level barrier; generation counter, so the same two words serve every half step.  The
// spin gives up after ~2 s of %globaltimer (a lost CTA must not hang the device; results are then garbage and the
// range flag is set to 2 so the host raises).
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int n_ctas, int* fail_flag) {
  __threadfence();                                            // release: this CTA's x_b rows, before it arrives
  const unsigned int gen = *reinterpret_cast<volatile unsigned int*>(bar + 1);
  if (atomicAdd(bar, 1u) == n_ctas - 1) {
    *reinterpret_cast<volatile unsigned int*>(bar) = 0u;
    __threadfence();
    atomicAdd(bar + 1, 1u);
  } else {
    const unsigned long long t0 = globaltimer_ns();
    while (*reinterpret_cast<volatile unsigned int*>(bar + 1) == gen) {
      __nanosleep(32);
      if (globaltimer_ns() - t0 > 2000000000ull) {
        if (fail_flag) *fail_flag = 2;
        break;
      }
    }
  }
  __threadfence();                                            // acquire: the other CTAs' rows, before anyone reads them
}

// Biases ride on the tensor pipe: layer l (< K-1) adds  Sel_l[128x16] * BiasTile[N x 16]^T  where Sel_l has ones in
// K-columns 2l, 2l+1 and BiasTile holds (hi, lo) of b_l in those columns.  (Re)built per half step by `nthr` threads.
template <int LAT, bool BF16>
__device__ __forceinline__ void build_bias_tiles(uint8_t* btile, float* blast, const float* bias_s, const float* bias_t,
                                                 int K, int t, int nthr) {
  for (int i = t; i < 2 * LAT * 8; i += nthr) {                // (mlp, n, layer slot j): k = 2j, 2j+1
    const int m = i / (LAT * 8), r = i - m * (LAT * 8), n = r >> 3, j = r & 7;
    const float b = (j < K - 1) ? (m ? bias_t : bias_s)[j * 256 + n] : 0.f;
    uint32_t hi, lo;
    split_pair<BF16>(b, 0.f, hi, lo);
    const uint32_t packed = (hi & 0xFFFFu) | (lo << 16);       // k = 2j -> hi(b), k = 2j+1 -> lo(b)
    const int k = 2 * j;
    *reinterpret_cast<uint32_t*>(btile + m * (LAT * 32) + (k >> 3) * (LAT * 16) + (n >> 3) * 128 + (n & 7) * 16 +
                                 (k & 7) * 2) = packed;
  }
  if (t >= 0 && t < 2 * kNOut) blast[t] = ((t >> 4) ? bias_t : bias_s)[(K - 1) * 256 + (t & 15)];
}

// Hidden layers issue ONE N = LAT tcgen05.mma per (k-step, product): measured per half step at the bench workload,
// two sequential N = 128 halves 0.317 ms, two N = 128 halves fed alternately (an independent MMA between two that share
// an accumulator) 0.321 ms, one N = 256 instruction 0.290 ms -- the cost is per instruction (~20 of 84 cycles at N = 128),
// not a read-after-write stall on the accumulator.
//
// PERSIST = false: one half coupling step per launch (2T launches per flow, chained by programmatic dependent launch).
// PERSIST = true : the WHOLE flow in one cooperative launch -- every CTA keeps its tiles (same static round robin) for
//                  all 2T half steps, a grid barrier stands where the launch boundary was, the weight ring streams
//                  straight on into the next half step's images, x halves are read with ld.global.cg (another SM
//                  rewrote them since this SM last cached them).
// MODE 2 (INJECT): per-half-step launch whose layer 0 was computed by the layered fp32 kernels (MLP inputs wider
//                  than 16: the dm_self_attn GNN's [x, attention] rows, node widths above 8): the epilogue warps read
//                  the layer-0 pre-activations from global memory where they would read the layer-0 accumulator from
//                  TMEM; layers 1..K-1, the coupling update and the log-det run here exactly as in MODE 0.
constexpr int kModeStep = 0, kModePersist = 1, kModeInject = 2;

template <int LAT, int NPROD, bool BF16, int ACT, int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_coupling_tc(const TcParams p) {
  constexpr bool PERSIST = MODE == kModePersist;
  constexpr bool INJECT = MODE == kModeInject;
  using G = Geo<LAT, kFwdNS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* hbuf = ring + kStages * kStageBytes;                 // [buf][hi|lo][4096]
  uint8_t* sel = hbuf + 2 * 2 * 4096;                           // [kMaxLayers-1][4096] selector A tiles
  uint8_t* btile = sel + (kMaxLayers - 1) * 4096;               // [2][LAT*32] bias B tiles (K-major, K=16)
  float* blast = (float*)(btile + 2 * LAT * 32);                // [2][16] last-layer bias
  float* hstage = blast + 2 * kNOut;                            // [128][17]
  Barriers* bars = (Barriers*)(hstage + kGatherThreads * 17);
  uint32_t* tmem_slot = (uint32_t*)(bars + 1);
  double* ldj_red = (double*)(tmem_slot + 2);                   // [4]
  int32_t* xred = (int32_t*)(ldj_red + 4);                      // [2][8] sender range reduction of the gather warps
  uint8_t* xstage = (uint8_t*)(((uintptr_t)(xred + 16) + 127) & ~(uintptr_t)127);   // [2][kXStageBytes]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = p.K;

  // ---- one-time setup -----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < kNS; ++i) mbar_init(smem_u32(&bars->acc_full[i]), 1);   // all of them: harmless
    for (int i = 0; i < kMaxNA; ++i) mbar_init(smem_u32(&bars->a_ready[i]), kEpiThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->h_full[i]), kGatherThreads);
      mbar_init(smem_u32(&bars->h_empty[i]), 1);
      mbar_init(smem_u32(&bars->x_full[i]), 1);
    }
    mbar_init(smem_u32(&bars->acc_last), 1);
    mbar_init(smem_u32(&bars->role_bar), kEpiThreads + kGatherThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n_halves = PERSIST ? p.n_halves : 1;
  {
    const uint16_t one = BF16 ? 0x3F80 : 0x3C00;
    for (int i = tid; i < (K - 1) * 128 * 16; i += kThreads) {
      const int l = i / 2048, r = i - l * 2048, m = r >> 4, k = r & 15;
      *reinterpret_cast<uint16_t*>(sel + l * 4096 + (k >> 3) * 2048 + m * 16 + (k & 7) * 2) =
          (k == 2 * l || k == 2 * l + 1) ? one : (uint16_t)0;
    }
    const float* const bs0 = PERSIST ? p.halves[0].bias[0] : p.one.bias[0];
    const float* const bt0 = PERSIST ? p.halves[0].bias[1] : p.one.bias[1];
    build_bias_tiles<LAT, BF16>(btile, blast, bs0, bt0, K, tid, kThreads);
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
      for (int hs = 0; hs < n_halves; ++hs) {
      // (runs ahead into the next half step's images)
      const uint8_t* const w_s = PERSIST ? p.halves[hs].w[0] : p.one.w[0];
      const uint8_t* const w_t = PERSIST ? p.halves[hs].w[1] : p.one.w[1];
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int m = 0; m < 2; ++m) {
          const uint8_t* src = m ? w_t : w_s;
          if constexpr (!INJECT) issue(src, NPROD >= 2 ? G::L0_BYTES : G::L0_MAT_BYTES);
          src += G::L0_BYTES;
          for (int l = 1; l < K - 1; ++l)
            for (int c = 0; c < kFwdNS * G::NKC; ++c) {
              issue(src, NPROD >= 2 ? G::CHUNK_BYTES : G::MAT_BYTES);
              src += G::CHUNK_BYTES;
            }
          issue(src, NPROD >= 2 ? G::LAST_BYTES : G::LAST_MAT_BYTES);
        }
      }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule, one elected lane issues =============
    {
      constexpr uint32_t idesc_l0 = make_idesc(LAT, BF16);
      constexpr uint32_t idesc_h = make_idesc(G::NH, BF16);
      constexpr uint32_t idesc_last = make_idesc(kNOut, BF16);
      const uint32_t bar_full = smem_u32(&bars->full[0]), bar_empty = smem_u32(&bars->empty[0]);
      const uint32_t bar_acc = smem_u32(&bars->acc_full[0]), bar_ready = smem_u32(&bars->a_ready[0]);
      const uint32_t bar_hfull = smem_u32(&bars->h_full[0]), bar_hempty = smem_u32(&bars->h_empty[0]);
      const uint32_t ring_u = smem_u32(ring), sel_u = smem_u32(sel), btile_u = smem_u32(btile);
      const uint32_t hbuf_u = smem_u32(hbuf);
      uint32_t stage = 0, phase = 0;
      uint32_t aready_par = 0;          // bit q = parity to wait for on a_ready[q]
      uint32_t region = 0;
      int it = 0;
      Tracer tr;
      tr.init(p.trace, 0, blockIdx.x == 0 && lane == 0);
      // column of the packed hi half of feature k inside a region (lo is 16 columns further)
      auto a_col = [](int k) { return (uint32_t)((k >> 5) * 32 + ((k & 31) >> 4) * 8); };
      auto wait_groups = [&](uint32_t& waited, int q_lo, int q_hi) {
        for (int q = q_lo; q <= q_hi; ++q)
          if (!(waited >> q & 1u)) {
            mbar_wait(bar_ready + 8 * q, (aready_par >> q) & 1u);
            aready_par ^= 1u << q;
            waited |= 1u << q;
          }
      };
      for (int hs = 0; hs < n_halves; ++hs)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        if constexpr (!INJECT) mbar_wait(bar_hfull + 8 * buf, (it >> 1) & 1);
        tr.ev(10, 0, 0);
        const uint32_t h_hi = hbuf_u + buf * 8192;
        const uint32_t h_lo = h_hi + 4096;
        for (int m = 0; m < 2; ++m) {
          const uint32_t bt = btile_u + m * (LAT * 32);
          // ---- layer 0: bias, then A = h (smem), B = W0 chunk, N = LAT, K = 16 ----------------
          if constexpr (INJECT) {
            region ^= 1;       // the epilogue warps write layer 0's activations into this region from global memory
          } else {
            const uint32_t d = tmem_base + region * LAT;
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sb = ring_u + stage * kStageBytes;
            if (elect_one()) {
              mma_ss(d, smem_desc(sel_u, 2048, 128), smem_desc(bt, LAT * 16, 128), idesc_l0, 0);
              const uint64_t a_hi = smem_desc(h_hi, 2048, 128), a_lo = smem_desc(h_lo, 2048, 128);
              const uint64_t b_hi = smem_desc(sb, LAT * 16, 128);
              const uint64_t b_lo = smem_desc(sb + G::L0_MAT_BYTES, LAT * 16, 128);
              mma_ss(d, a_hi, b_hi, idesc_l0, 1);
              if (NPROD == 3) mma_ss(d, a_lo, b_hi, idesc_l0, 1);   // activation residual
              if (NPROD >= 2) mma_ss(d, a_hi, b_lo, idesc_l0, 1);   // weight residual
              tc_commit(bar_empty + 8 * stage);
#pragma unroll
              for (int ph = 0; ph < kFwdNS; ++ph) tc_commit(bar_acc + 8 * ph);
              if (m == 1) tc_commit(bar_hempty + 8 * buf);
            }
            __syncwarp();
            tr.ev(11, m * 16, 0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            region ^= 1;
          }
          // ---- hidden layers: A from TMEM (previous region), N = NH per half -----------------
          for (int l = 1; l < K - 1; ++l) {
            const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
            const uint32_t out_col = tmem_base + region * LAT;
            const uint32_t sel_l = sel_u + l * 4096;
            uint32_t waited = 0;
#pragma unroll 1
            for (int ph = 0; ph < kFwdNS; ++ph) {
              const uint32_t d = out_col + ph * G::NH;
#pragma unroll 1
              for (int kc = 0; kc < G::NKC; ++kc) {
                wait_groups(waited, (kc * G::KC) >> 6, (kc * G::KC + G::KC - 1) >> 6);
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sb = ring_u + stage * kStageBytes;
                const uint32_t a0 = in_col + a_col(kc * G::KC);
                tr.ev(12, m * 16 + l, ph * 4 + kc);
                if (elect_one()) {
                  if (kc == 0)   // bias first: it overwrites (accumulate = 0) the stale region contents
                    mma_ss(d, smem_desc(sel_l, 2048, 128), smem_desc(bt + ph * G::NH * 16, LAT * 16, 128), idesc_h, 0);
                  const uint64_t b_hi0 = smem_desc(sb, G::NH * 16, 128);
                  const uint64_t b_lo0 = smem_desc(sb + G::MAT_BYTES, G::NH * 16, 128);
#pragma unroll
                  for (int ks = 0; ks < G::KC / 16; ++ks) {
                    // K advances by 16 features = 2 core matrices = 2*LBO bytes (>>4 in the descriptor)
                    const uint64_t koff = (uint64_t)((ks * 2 * (G::NH * 16)) >> 4);
                    const uint32_t a_hi = a0 + a_col(ks * 16);
                    mma_ts(d, a_hi, b_hi0 + koff, idesc_h, 1);
                    if (NPROD == 3) mma_ts(d, a_hi + 16, b_hi0 + koff, idesc_h, 1);
                    if (NPROD >= 2) mma_ts(d, a_hi, b_lo0 + koff, idesc_h, 1);
                  }
                  tc_commit(bar_empty + 8 * stage);
                  if (kc == G::NKC - 1) tc_commit(bar_acc + 8 * ph);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
            }
            region ^= 1;
          }
          // ---- last layer: N = 16, K = LAT, one chunk (bias added by the epilogue) -----------
          {
            const uint32_t in_col = tmem_base + (region ^ 1) * LAT;
            const uint32_t d = tmem_base + region * LAT;
            uint32_t waited = 0;
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sb = ring_u + stage * kStageBytes;
            tr.ev(13, m * 16 + K - 1, 0);
            // one 64-feature group of the previous layer at a time: the k-steps of the groups converted while that
            // layer's second accumulator half was still being multiplied run under the epilogue of that half
#pragma unroll 1
            for (int qg = 0; qg < G::NA; ++qg) {
              wait_groups(waited, qg, qg);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t b_hi0 = smem_desc(sb, kNOut * 16, 128);
                const uint64_t b_lo0 = smem_desc(sb + G::LAST_MAT_BYTES, kNOut * 16, 128);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const int ks = qg * 4 + k4;
                  const uint64_t koff = (uint64_t)((ks * 2 * (kNOut * 16)) >> 4);
                  const uint32_t a_hi = in_col + a_col(ks * 16);
                  mma_ts(d, a_hi, b_hi0 + koff, idesc_last, ks ? 1u : 0u);
                  if (NPROD == 3) mma_ts(d, a_hi + 16, b_hi0 + koff, idesc_last, 1);
                  if (NPROD >= 2) mma_ts(d, a_hi, b_lo0 + koff, idesc_last, 1);
                }
                if (qg == G::NA - 1) {
                  tc_commit(bar_empty + 8 * stage);
                  tc_commit(smem_u32(&bars->acc_last));
                }
              }
              __syncwarp();
            }
            tr.ev(14, m * 16 + K - 1, 0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            region ^= 1;
          }
        }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===== epilogue warps ======================================================================
    asm volatile("griddepcontrol.wait;" ::: "memory");      // x_b and the log-det come from the previous launch
    // kernel duration from the device's own clock: first CTA past the dependency wait -> last CTA done (no host
    // events between launches, so programmatic dependent launch overlaps exactly as in production)
    if (p.tstamp && tid == 64) atomicMin(p.tstamp, globaltimer_ns());
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;             // which 32-column chunk of every 64-column group
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0;
    int region = 0;
    double ldj_local = 0.0;
    float amax = 0.f;
    uint32_t role_par = 0;
    const int hp4 = p.HP >> 2;
    Tracer tr;
    tr.init(p.trace, 1 + (warp - 2), blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6));
    for (int hs = 0; hs < n_halves; ++hs) {
    const int swap = PERSIST ? p.halves[hs].swap : 0;
    float* const xb_base = swap ? p.x[0] : p.x[1];
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      float st[2][kNOut];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        for (int l = 0; l < K - 1; ++l) {
          const bool from_global = INJECT && l == 0;
#pragma unroll
          for (int ph = 0; ph < kFwdNS; ++ph) {
            if (!from_global) {
              mbar_wait(smem_u32(&bars->acc_full[ph]), (acc_par >> ph) & 1u);
              acc_par ^= 1u << ph;
              tc_fence_after();
            }
            tr.ev(20, m * 16 + l, ph * 4);
            const uint32_t t0 = lane_base + region * LAT + ph * G::NH + grp * 32;
            // INJECT, layer 0: this thread's row of the pre-activation matrix, the same 32-column chunks
            const int64_t inj_node = (int64_t)tile * kTileM + row;
            const float* inj = (INJECT && inj_node < p.n_nodes)
                                   ? p.pre0[m] + inj_node * LAT + ph * G::NH + grp * 32 : nullptr;
            auto load_chunk = [&](const float* src, uint32_t (&v)[32]) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 q = src ? __ldg(reinterpret_cast<const float4*>(src) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * j] = __float_as_uint(q.x); v[4 * j + 1] = __float_as_uint(q.y);
                v[4 * j + 2] = __float_as_uint(q.z); v[4 * j + 3] = __float_as_uint(q.w);
              }
            };
            // this warp's 32-column chunk of every 64-column group, in K order (the next layer starts on group 0
            // while the later groups are still being converted); two chunks in flight: the next one is loaded before
            // the current one is converted
            uint32_t va[32], vb[32];
            auto fetch = [&](int g, uint32_t (&v)[32]) {
              if (from_global) load_chunk(inj ? inj + 64 * g : nullptr, v);
              else tmem_ld32(t0 + 64 * g, v);
            };
            fetch(0, va);
#pragma unroll
            for (int g = 0; g < G::GPH; ++g) {
              uint32_t (&cur)[32] = (g & 1) ? vb : va;
              uint32_t (&nxt)[32] = (g & 1) ? va : vb;
              if (g + 1 < G::GPH) fetch(g + 1, nxt);
              if (!from_global) tmem_wait_ld();
              convert_chunk<NPROD, BF16, ACT>(t0 + 64 * g, cur, amax);
              tmem_wait_st();
              tc_fence_before();
              mbar_arrive(smem_u32(&bars->a_ready[ph * G::GPH + g]));
            }
            tr.ev(21, m * 16 + l, ph * 4);
          }
          region ^= 1;
        }
        // last layer: s (m == 0) or t (m == 1); only the first warp of each lane quarter reads it
        {
          mbar_wait(smem_u32(&bars->acc_last), (acc_par >> 8) & 1u);
          acc_par ^= 1u << 8;
          tc_fence_after();
          tr.ev(22, m * 16 + K - 1, 0);
          if (grp == 0) {
            uint32_t v[16];
            tmem_ld16(lane_base + region * LAT, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < kNOut; ++j) st[m][j] = __uint_as_float(v[j]) + blast[m * kNOut + j];
          }
          region ^= 1;
        }
      }
      // ---- affine coupling update + log-det partial (gnn.py:322-323 / :359) ------------------
      const int64_t node = (int64_t)tile * kTileM + row;
      if (grp == 0 && node < p.n_nodes) {
        float* xrow = xb_base + node * p.HP;
#pragma unroll
        for (int g4 = 0; g4 < kNOut / 4; ++g4) {
          if (g4 < hp4) {
            float4 x = PERSIST ? __ldcg(reinterpret_cast<const float4*>(xrow + g4 * 4))
                               : *reinterpret_cast<const float4*>(xrow + g4 * 4);
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
    if (PERSIST && hs + 1 < n_halves) {
      // where the launch boundary was: every x_b row of this half step written (CTA barrier, then the grid barrier
      // with its fences), the next half step's bias tiles rebuilt; the gather warps stand at the same two CTA barriers
      mbar_arrive(smem_u32(&bars->role_bar));
      mbar_wait(smem_u32(&bars->role_bar), role_par);
      role_par ^= 1u;
      if (tid == 64) grid_barrier(p.grid_bar, gridDim.x, p.range_flag);
      build_bias_tiles<LAT, BF16>(btile, blast, p.halves[hs + 1].bias[0], p.halves[hs + 1].bias[1], K, tid - 64,
                                  kEpiThreads + kGatherThreads);
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->role_bar));
      mbar_wait(smem_u32(&bars->role_bar), role_par);
      role_par ^= 1u;
    }
    }
    tr.ev(23, 0, 0);
    if (p.tstamp && tid == 64) atomicMax(p.tstamp + 1, globaltimer_ns());
    if (!BF16 && amax > 65504.f && p.range_flag) *p.range_flag = 1;
    // fixed-order reduction of the log-det partial: lanes, then the 4 lane-quarter warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ldj_local += __shfl_down_sync(0xffffffffu, ldj_local, o);
    if (lane == 0 && grp == 0) ldj_red[q] = ldj_local;
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
    if (warp == 2 && p.partials) {
      // log-det hand-off without a second launch: the last CTA to arrive folds every CTA's partial in a FIXED order
      // (lane-strided sums, then a fixed shuffle tree: the same bits whatever the arrival order), 32 loads in flight
      // instead of one thread walking the array, and resets the counter for the next launch
      unsigned prev = 0;
      if (lane == 0) {
        p.partials[blockIdx.x] = ((ldj_red[0] + ldj_red[1]) + ldj_red[2]) + ldj_red[3];
        if (p.ldj_accum) {
          __threadfence();
          prev = atomicAdd(p.counter, 1u);
        }
      }
      prev = __shfl_sync(0xffffffffu, prev, 0);
      if (p.ldj_accum && prev == gridDim.x - 1) {
        __threadfence();
        const volatile double* part = p.partials;
        double acc = 0.0;
        for (unsigned i = lane; i < gridDim.x; i += 32) acc += part[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          *p.ldj_accum += acc;
          *p.counter = 0u;
        }
      }
    }
  } else {
    // ===== gather warps: a3 + a4 + a5 =========================================================
    asm volatile("griddepcontrol.wait;" ::: "memory");      // x_a comes from the previous launch
    const int row = tid - (kThreads - kGatherThreads);
    float* my = hstage + row * 17;
    const int hp4 = p.HP >> 2;
    int it = 0;
    float amax = 0.f;
    uint32_t role_par = 0, xpar = 0;
    Tracer tr;
    tr.init(p.trace, 9, blockIdx.x == 0 && row == 0);
    for (int hs = 0; hs < (INJECT ? 0 : n_halves); ++hs) {      // INJECT: the MLP input never enters this kernel
    const int swap = PERSIST ? p.halves[hs].swap : 0;
    const float* const xa_base = swap ? p.x[1] : p.x[0];
    // persistent launch: rows of xa were rewritten by other SMs one half step ago -> bypass this SM's L1
    auto ldx = [](const float* q) -> float4 {
      return PERSIST ? __ldcg(reinterpret_cast<const float4*>(q)) : *reinterpret_cast<const float4*>(q);
    };
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tr.ev(30, 0, 0);
      const int64_t node = (int64_t)tile * kTileM + row;
      float self[kNOut], agg[kNOut];
#pragma unroll
      for (int j = 0; j < kNOut; ++j) { self[j] = 0.f; agg[j] = 0.f; }
      // ---- stage the tile's sender rows with ONE bulk copy (MODE 0; the persistent launch keeps ld.global.cg: its
      //      rows were written through the generic proxy by this and other CTAs one half step ago) -----------------
      int32_t e_beg = 0, e_end = 0;
      if (node < p.n_nodes) { e_beg = p.rowptr[node]; e_end = p.rowptr[node + 1]; }
      bool staged = false;
      int32_t x_lo = 0;
      const int xbuf = it & 1;
      if constexpr (!PERSIST) {
        int32_t lo_t = 0x7fffffff, hi_t = -1;
        for (int32_t e = e_beg; e < e_end; ++e) {
          const int32_t sidx = p.csr[e];
          lo_t = min(lo_t, sidx);
          hi_t = max(hi_t, sidx);
        }
        lo_t = __reduce_min_sync(0xffffffffu, lo_t);
        hi_t = __reduce_max_sync(0xffffffffu, hi_t);
        int32_t* red = xred + xbuf * 8;
        if ((row & 31) == 0) { red[row >> 5] = lo_t; red[4 + (row >> 5)] = hi_t; }
        __syncwarp();
        asm volatile("bar.sync 3, %0;" ::"r"(kGatherThreads) : "memory");   // all 128 gather threads, converged here
        x_lo = min(min(red[0], red[1]), min(red[2], red[3]));
        const int32_t x_hi = max(max(red[4], red[5]), max(red[6], red[7]));
        const uint32_t bytes = (uint32_t)(x_hi - x_lo + 1) * (uint32_t)p.HP * 4u;
        staged = x_hi >= x_lo && (int64_t)(x_hi - x_lo + 1) * p.HP * 4 <= kXStageBytes;
        if (staged) {
          const uint32_t xb = smem_u32(&bars->x_full[xbuf]);
          if (row == 0) {
            mbar_expect_tx(xb, bytes);
            bulk_g2s(smem_u32(xstage + xbuf * kXStageBytes), xa_base + (int64_t)x_lo * p.HP, bytes, xb);
          }
          mbar_wait(xb, (xpar >> xbuf) & 1u);
          xpar ^= 1u << xbuf;
        }
      }
      if (node < p.n_nodes) {
        const float* xr = xa_base + node * p.HP;
#pragma unroll
        for (int g4 = 0; g4 < kNOut / 4; ++g4)
          if (g4 < hp4) {
            float4 x = ldx(xr + g4 * 4);
            self[g4 * 4] = x.x; self[g4 * 4 + 1] = x.y; self[g4 * 4 + 2] = x.z; self[g4 * 4 + 3] = x.w;
          }
        int32_t e = e_beg;
        const int32_t end = e_end;
        const int32_t cnt = end - e;
        if (staged) {
          const float* xs = reinterpret_cast<const float*>(xstage + xbuf * kXStageBytes);
          for (; e < end; ++e) {     // ascending edge index inside the segment: TF-CPU order
            const float* sr = xs + (p.csr[e] - x_lo) * p.HP;
#pragma unroll
            for (int g4 = 0; g4 < kNOut / 4; ++g4)
              if (g4 < hp4) {
                const float4 x = *reinterpret_cast<const float4*>(sr + g4 * 4);
                agg[g4 * 4] = __fadd_rn(agg[g4 * 4], x.x);
                agg[g4 * 4 + 1] = __fadd_rn(agg[g4 * 4 + 1], x.y);
                agg[g4 * 4 + 2] = __fadd_rn(agg[g4 * 4 + 2], x.z);
                agg[g4 * 4 + 3] = __fadd_rn(agg[g4 * 4 + 3], x.w);
              }
          }
        }
        for (; e < end; ++e) {       // (not staged) rows straight from global memory, same order
          const float* sr = xa_base + (int64_t)p.csr[e] * p.HP;
#pragma unroll
          for (int g4 = 0; g4 < kNOut / 4; ++g4)
            if (g4 < hp4) {
              float4 x = ldx(sr + g4 * 4);
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
      for (int j = 0; j < 8; ++j) {
        if constexpr (!BF16) amax = fmaxf(amax, fmaxf(fabsf(my[2 * j]), fabsf(my[2 * j + 1])));
        split_pair<BF16>(my[2 * j], my[2 * j + 1], hi[j], lo[j]);
      }
      tr.ev(31, 0, 0);
      mbar_wait(smem_u32(&bars->h_empty[buf]), ((it >> 1) & 1) ^ 1);
      tr.ev(32, 0, 0);
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
    if (PERSIST && hs + 1 < n_halves) {
      mbar_arrive(smem_u32(&bars->role_bar));
      mbar_wait(smem_u32(&bars->role_bar), role_par);
      role_par ^= 1u;
      build_bias_tiles<LAT, BF16>(btile, blast, p.halves[hs + 1].bias[0], p.halves[hs + 1].bias[1], K, tid - 64,
                                  kEpiThreads + kGatherThreads);
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->role_bar));
      mbar_wait(smem_u32(&bars->role_bar), role_par);
      role_par ^= 1u;
    }
    }
    if (!BF16 && amax > 65504.f && p.range_flag) *p.range_flag = 1;
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// (the weight images are written by pack.cu: one launch for every format)
template <int LAT>
size_t bytes_per_mlp_t(int K) {
  using G = Geo<LAT, kFwdNS>;
  return (size_t)G::L0_BYTES + (size_t)(K - 2) * kFwdNS * G::NKC * G::CHUNK_BYTES + G::LAST_BYTES;
}

template <int LAT, int NPROD, bool BF16, int ACT, int MODE>
int launch_tc_act(const TcParams& p, int grid, cudaStream_t stream) {
  constexpr bool PERSIST = MODE == kModePersist;
  auto kern = k_coupling_tc<LAT, NPROD, BF16, ACT, MODE>;
  static bool configured[kMaxDevices] = {};
  const size_t smem = smem_bytes<LAT>();
  if (first_use_on_device(configured))
    GNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (PERSIST) {
    // whole flow in one launch: the CTAs meet at a grid barrier between half steps, so all of them must be resident
    // (grid <= number of SMs, one CTA per SM): cooperative launch makes the runtime check exactly that
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    // Programmatic dependent launch: the CTAs of this launch may start on an SM as soon as the previous kernel's
    // CTA there has exited, run their prologue (barrier init, TMEM allocation, bias tiles, first weight chunks) and
    // block in griddepcontrol.wait before the first read of anything the previous kernel wrote (x halves, log-det).
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool pdl = getenv("GNF_NO_PDL") == nullptr;   // GNF_NO_PDL=1: plain stream-ordered launches (A/B timing)
    cfg.numAttrs = pdl ? 1 : 0;
  }
  GNF_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  GNF_LAUNCH_CHECK();
  return GNF_OK;
}

template <int LAT, int NPROD, bool BF16, int MODE>
int launch_tc(const TcParams& p, int grid, cudaStream_t stream) {
  return p.act == GNF_ACT_LEAKY_RELU ? launch_tc_act<LAT, NPROD, BF16, GNF_ACT_LEAKY_RELU, MODE>(p, grid, stream)
                                     : launch_tc_act<LAT, NPROD, BF16, GNF_ACT_RELU, MODE>(p, grid, stream);
}

template <int MODE>
int launch_tc_math(const TcParams& p, int L, int math, int grid, cudaStream_t stream) {
  if (L == 256) {
    if (math == GNF_MATH_TC2X) return launch_tc<256, 2, false, MODE>(p, grid, stream);
    if (math == GNF_MATH_TC3X) return launch_tc<256, 3, false, MODE>(p, grid, stream);
    if (math == GNF_MATH_TC3X_BF16) return launch_tc<256, 3, true, MODE>(p, grid, stream);
    return launch_tc<256, 1, true, MODE>(p, grid, stream);
  }
  if (math == GNF_MATH_TC2X) return launch_tc<128, 2, false, MODE>(p, grid, stream);
  if (math == GNF_MATH_TC3X) return launch_tc<128, 3, false, MODE>(p, grid, stream);
  if (math == GNF_MATH_TC3X_BF16) return launch_tc<128, 3, true, MODE>(p, grid, stream);
  return launch_tc<128, 1, true, MODE>(p, grid, stream);
}

}  // namespace

static unsigned long long* g_trace = nullptr;
void tc_set_trace(void* buf) { g_trace = (unsigned long long*)buf; }

// Optional per-launch device timestamps (gnf_debug_kernel_timing): every fused launch gets a {min start, max end}
// %globaltimer slot, written by the kernel itself, so bench.py reads the kernel's duration INSIDE its timed steps
// without host events between the launches (those would defeat the overlap of programmatic dependent launch).
namespace {
constexpr int kMaxTimed = 8192;
struct KernelTimer {
  bool on = false;
  int n = 0;
  unsigned long long* slots = nullptr;     // device [kMaxTimed][2]
};
KernelTimer g_timer;
}  // namespace

int tc_kernel_timing(int enable) {
  if (enable) {
    if (!g_timer.slots) GNF_CUDA(cudaMalloc(&g_timer.slots, (size_t)kMaxTimed * 16));
    static unsigned long long init[kMaxTimed * 2];
    for (int i = 0; i < kMaxTimed; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
    GNF_CUDA(cudaMemcpy(g_timer.slots, init, sizeof(init), cudaMemcpyHostToDevice));
    g_timer.n = 0;
  }
  g_timer.on = enable != 0;
  return GNF_OK;
}

int tc_kernel_time(double* total_ms, int64_t* launches) {
  double tot = 0.0;
  int64_t cnt = 0;
  if (g_timer.slots && g_timer.n > 0) {
    static unsigned long long host[kMaxTimed * 2];
    GNF_CUDA(cudaDeviceSynchronize());
    GNF_CUDA(cudaMemcpy(host, g_timer.slots, (size_t)g_timer.n * 16, cudaMemcpyDeviceToHost));
    for (int i = 0; i < g_timer.n; ++i)
      if (host[2 * i + 1] >= host[2 * i] && host[2 * i] != ~0ull) {
        tot += (double)(host[2 * i + 1] - host[2 * i]) * 1e-6;
        ++cnt;
      }
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = cnt;
  return GNF_OK;
}

bool tc_shape_supported(const Flow& f) {
  return !f.attn && (f.L == 128 || f.L == 256) && f.in_dim <= kK0 && f.H <= kNOut && f.K >= 2 && f.K <= kMaxLayers;
}

size_t tc_bytes_per_mlp(int L, int K) { return L == 256 ? bytes_per_mlp_t<256>(K) : bytes_per_mlp_t<128>(K); }

int tc_coupling_half(const Flow& f, int mlp_s, int mlp_t, int math, int inverse, const float* xa,
                     float* xb, int64_t n_nodes, const int32_t* rowptr, const int32_t* csr_senders,
                     double* ldj_partials, double* ldj_accum, unsigned int* counter, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcParams p = {};
  p.x[0] = const_cast<float*>(xa);
  p.x[1] = xb;
  p.rowptr = rowptr;
  p.csr = csr_senders;
  p.n_nodes = n_nodes;
  p.n_tiles = (int)ceil_div(n_nodes, kTileM);
  const int img = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 0 : 1;
  p.one.w[0] = f.wtc[img] + (size_t)mlp_s * f.wtc_per_mlp;
  p.one.w[1] = f.wtc[img] + (size_t)mlp_t * f.wtc_per_mlp;
  p.one.bias[0] = f.btc + (size_t)mlp_s * f.K * 256;
  p.one.bias[1] = f.btc + (size_t)mlp_t * f.K * 256;
  p.one.swap = 0;
  p.n_halves = 1;
  p.K = f.K;
  p.H = f.H;
  p.HP = f.HP;
  p.concat = f.d.block == GNF_BLOCK_CONCAT;
  p.mean = f.d.agg == GNF_AGG_MEAN;
  p.act = f.d.act;
  p.inverse = inverse;
  p.eps = f.d.eps;
  p.partials = ldj_partials;
  p.ldj_accum = (!inverse && counter) ? ldj_accum : nullptr;
  p.counter = counter;
  p.range_flag = f.range_flag;
  p.trace = g_trace;
  int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  p.tstamp = (g_timer.on && g_timer.n < kMaxTimed) ? g_timer.slots + 2 * (size_t)g_timer.n++ : nullptr;
  return launch_tc_math<kModeStep>(p, f.L, math, grid, stream);
}

// Flows whose MLP input is wider than the fused kernel's 16-column layer-0 tile (dm_self_attn: [x, attention] rows;
// message passing with node_embedding_dim > 16): layer 0 runs in the layered fp32 kernels, everything after it here.
bool tc_inject_supported(const Flow& f) {
  const bool plain_attn = f.attn && !(f.attn_flags & (GNF_ATTN_RESIDUAL | GNF_ATTN_LAYER_NORM));
  return (f.L == 128 || f.L == 256) && f.H <= kNOut && f.K >= 2 && f.K <= kMaxLayers &&
         (plain_attn || (!f.attn && f.in_dim > kK0));
}

int tc_coupling_inject(const Flow& f, int mlp_s, int mlp_t, int math, int inverse, const float* pre0_s,
                       const float* pre0_t, float* xb, int64_t n_nodes, double* ldj_partials, double* ldj_accum,
                       unsigned int* counter, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcParams p = {};
  p.x[0] = nullptr;
  p.x[1] = xb;
  p.pre0[0] = pre0_s;
  p.pre0[1] = pre0_t;
  p.n_nodes = n_nodes;
  p.n_tiles = (int)ceil_div(n_nodes, kTileM);
  const int img = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 0 : 1;
  p.one.w[0] = f.wtc[img] + (size_t)mlp_s * f.wtc_per_mlp;
  p.one.w[1] = f.wtc[img] + (size_t)mlp_t * f.wtc_per_mlp;
  p.one.bias[0] = f.btc + (size_t)mlp_s * f.K * 256;
  p.one.bias[1] = f.btc + (size_t)mlp_t * f.K * 256;
  p.one.swap = 0;
  p.n_halves = 1;
  p.K = f.K;
  p.H = f.H;
  p.HP = f.HP;
  p.act = f.d.act;
  p.inverse = inverse;
  p.partials = ldj_partials;
  p.ldj_accum = (!inverse && counter) ? ldj_accum : nullptr;
  p.counter = counter;
  p.range_flag = f.range_flag;
  p.trace = g_trace;
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  p.tstamp = (g_timer.on && g_timer.n < kMaxTimed) ? g_timer.slots + 2 * (size_t)g_timer.n++ : nullptr;
  return launch_tc_math<kModeInject>(p, f.L, math, grid, stream);
}

// Half-step tables of the persistent launch, per weight image (fp16 / bf16) and direction (f: step-major, half 0 then
// half 1, gnn.py:309-338; g: reversed steps, half 1 then half 0, gnn.py:347-372).  Built once per flow.
int tc_build_half_tables(Flow& f) {
  if (!f.tc_ok) return GNF_OK;          // (the persistent launch serves the fully fused shapes only)
  const int T = f.d.num_timesteps, n = 2 * T;
  std::vector<HalfDesc> tab((size_t)4 * n);
  for (int img = 0; img < 2; ++img)
    for (int inv = 0; inv < 2; ++inv)
      for (int k = 0; k < n; ++k) {
        int step, half;
        if (!inv) { step = k >> 1; half = k & 1; }
        else { step = T - 1 - (k >> 1); half = 1 - (k & 1); }
        const int ms = f.mlp_index(0, half, step), mt = f.mlp_index(1, half, step);
        HalfDesc& d = tab[((size_t)img * 2 + inv) * n + k];
        d.w[0] = f.wtc[img] + (size_t)ms * f.wtc_per_mlp;
        d.w[1] = f.wtc[img] + (size_t)mt * f.wtc_per_mlp;
        d.bias[0] = f.btc + (size_t)ms * f.K * 256;
        d.bias[1] = f.btc + (size_t)mt * f.K * 256;
        d.swap = half;
        d.pad = 0;
      }
  GNF_CUDA(cudaMalloc(&f.half_tables, tab.size() * sizeof(HalfDesc)));
  GNF_CUDA(cudaMemcpy(f.half_tables, tab.data(), tab.size() * sizeof(HalfDesc), cudaMemcpyHostToDevice));
  return GNF_OK;
}

bool tc_persistent_wanted(int64_t n_nodes) {
  // GNF_PERSIST=0: one launch per half step always; =1: whole flow in one cooperative launch always;
  // default: the cooperative launch up to 2 tiles per SM (GNF_PERSIST_TILES_PER_SM), i.e. small batches, where the
  // launch boundaries are a visible share of the step (protein B=256: 0.627 -> 0.574 ms); at 8.8 tiles per SM the 2T
  // launches chained by programmatic dependent launch are as fast (A/B in profiles/r2_ab_launch_schemes.jsonl)
  const char* env = getenv("GNF_PERSIST");          // read per call: tests and A/B timings flip it at run time
  if (env && env[0] == '0') return false;
  if (env && env[0] == '1') return true;
  const char* lim = getenv("GNF_PERSIST_TILES_PER_SM");
  const int64_t per_sm = lim ? atoll(lim) : 2;
  return ceil_div(n_nodes, kTileM) <= per_sm * num_sms();
}

int tc_flow_persistent(const Flow& f, int math, int inverse, float* x0, float* x1, int64_t n_nodes,
                       const int32_t* rowptr, const int32_t* csr_senders, double* ldj_partials, double* ldj_accum,
                       unsigned int* counter, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcParams p = {};
  p.x[0] = x0;
  p.x[1] = x1;
  p.rowptr = rowptr;
  p.csr = csr_senders;
  p.n_nodes = n_nodes;
  p.n_tiles = (int)ceil_div(n_nodes, kTileM);
  const int img = (math == GNF_MATH_TC3X || math == GNF_MATH_TC2X) ? 0 : 1;
  p.n_halves = 2 * f.d.num_timesteps;
  p.halves = (const HalfDesc*)f.half_tables + ((size_t)img * 2 + (inverse ? 1 : 0)) * p.n_halves;
  p.grid_bar = counter + 16;                 // counter block is 256 bytes, zeroed by the caller
  p.K = f.K;
  p.H = f.H;
  p.HP = f.HP;
  p.concat = f.d.block == GNF_BLOCK_CONCAT;
  p.mean = f.d.agg == GNF_AGG_MEAN;
  p.act = f.d.act;
  p.inverse = inverse;
  p.eps = f.d.eps;
  p.partials = ldj_partials;
  p.ldj_accum = inverse ? nullptr : ldj_accum;
  p.counter = counter;
  p.range_flag = f.range_flag;
  p.trace = g_trace;
  p.tstamp = (g_timer.on && g_timer.n < kMaxTimed) ? g_timer.slots + 2 * (size_t)g_timer.n++ : nullptr;
  const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  return launch_tc_math<kModePersist>(p, f.L, math, grid, stream);
}

}  // namespace gnf
