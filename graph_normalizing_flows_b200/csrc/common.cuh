// Shared host/device helpers for libgnf_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gnf_b200.h"

namespace gnf {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int64_t& launch_counter();

#define GNF_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      ::gnf::set_error(__VA_ARGS__);                      \
      return (code);                                      \
    }                                                     \
  } while (0)

#define GNF_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::gnf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return GNF_ECUDA;                                                             \
    }                                                                               \
  } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define GNF_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ::gnf::launch_counter() += 1;                                                   \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      ::gnf::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                       __FILE__, __LINE__);                                         \
      return GNF_ECUDA;                                                             \
    }                                                                               \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int pad16(int x) { return (x + 15) / 16 * 16; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

// Per-device one-time setup (cudaFuncSetAttribute belongs to the device's copy of a function): returns true the
// first time it is called with `done` on the current device.
constexpr int kMaxDevices = 64;
static inline bool first_use_on_device(bool (&done)[kMaxDevices]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// ---- packed flow ---------------------------------------------------------------------------
constexpr int kMaxLayers = 8;

// One MLP of the flow, in every format a kernel family consumes.
struct MlpView {
  // fp32 layered path: per layer W [in_pad, out_pad] row-major (zero padded), b [out_pad]
  const float* w32[kMaxLayers];
  const float* b32[kMaxLayers];
};

struct Flow {
  gnf_flow_desc d;
  int H, HP, in_dim, in_pad, L, K;
  int n_mlps;                  // 4*T, or 4 when weight_sharing
  int64_t params_per_mlp;
  // fp32 packed weights
  float* w32 = nullptr;        // all MLPs, all layers
  int64_t w32_per_mlp = 0;
  int64_t w32_layer_off[kMaxLayers];
  int64_t b32_layer_off[kMaxLayers];
  int in_pads[kMaxLayers], out_pads[kMaxLayers], ins[kMaxLayers], outs[kMaxLayers];
  // transposed fp32 weights for the backward dX GEMMs: per layer W^T [out_pad8, in_pad]
  float* w32T = nullptr;
  int64_t w32T_per_mlp = 0;
  int64_t w32T_layer_off[kMaxLayers];
  int out_pad8[kMaxLayers];
  int64_t flat_w_off[kMaxLayers], flat_b_off[kMaxLayers];   // offsets inside one MLP of the flat parameter vector
  // f1 attention block (GNF_BLOCK_DM_ATTN): padded fp32 projections per GNN
  int attn = 0, heads = 0, kq = 0, vd = 0, cho = 0, attn_flags = 0;
  int hp8 = 0, qk_pad = 0, v_pad = 0, hv_pad = 0, cho_pad = 0;
  int64_t mlp_off = 0;                    // offset of the MLP inside one GNN's flat parameters
  float* wattn = nullptr;
  int64_t wattn_per_mlp = 0, wq_off = 0, wk_off = 0, wv_off = 0, wo_off = 0;
  float* zeros = nullptr;                 // zero bias for the bias-free projections
  float* wln = nullptr;                   // LayerNorm gamma|beta per GNN: [n_mlps][2*HP] (GNF_ATTN_LAYER_NORM)
  int64_t ln_off = 0;                     // offset of gamma inside one GNN's flat parameters
  float* wattnT = nullptr;                // transposed projections for the backward dX GEMMs: W^T [out_pad, in_pad]
  int64_t wattnT_per_mlp = 0, wqT_off = 0, wkT_off = 0, wvT_off = 0, woT_off = 0;
  // tensor-core packed weights (two images: fp16 and bf16 element type), per MLP a stream of
  // shared-memory chunk images in consumption order (see coupling_tc.cu)
  uint8_t* wtc[2] = {nullptr, nullptr};   // [0]=fp16 hi/lo, [1]=bf16 hi/lo; forward kernel's chunk geometry (kFwdNS)
  uint8_t* wtcB[2] = {nullptr, nullptr};  // the same MLPs in the backward kernel's chunk geometry (kNS): its recompute chains
  int64_t wtc_per_mlp = 0;
  float* btc = nullptr;        // biases for the TC kernel: per MLP K*256 floats
  uint8_t* wtcT = nullptr;     // bf16 hi/lo images of the TRANSPOSED MLP chains (backward dX), same geometry
  bool tc_ok = false;          // fully fused tcgen05 kernel (MLP input <= 16)
  bool tc_inject = false;      // layer 0 in k_linear_tc, layers 1..K-1 + coupling in the tcgen05 kernel (MODE inject)
  void* pack_jobs = nullptr;   // device job table of the one-launch re-pack (pack.cu)
  int n_pack_jobs = 0, pack_blocks = 0;
  // stand-alone tensor-core linears (linear_tc.cu): fp16 / bf16 hi-lo images per GNN of Wq, Wk, Wv, Wo (attention) and
  // of layer 0 (inject flows); offset < 0: that matrix stays on the fp32 kernels
  uint8_t* wlin[2] = {nullptr, nullptr};
  int64_t wlin_per_mlp = 0;
  int64_t lin_off[7] = {-1, -1, -1, -1, -1, -1, -1};   // q, k, v, o, layer 0; backward dX: W0^T (L -> in), Wo^T (cho -> heads*v)
  // layered tensor-core path (gemm_tc.cu) for the shapes the fused kernels do not take (latent_dim not in {128, 256},
  // residual / layer_norm attention variants): fp16 / bf16 hi-lo images of every MLP layer, k_gemm_tc geometry
  uint8_t* wgemm[2] = {nullptr, nullptr};
  int64_t wgemm_per_mlp = 0;
  int64_t gemm_off[kMaxLayers];
  bool tc_layered = false;     // MLP layers in k_gemm_tc, input assembly / coupling update in the fp32 path's kernels
  uint8_t* wgemmT = nullptr;   // bf16 hi-lo images of every W_l^T (backward dX = delta W_l^T of the layered flows)
  int64_t wgemmT_per_mlp = 0;
  int64_t gemmT_off[kMaxLayers];
  void* half_tables = nullptr; // device HalfDesc[2 images][2 directions][2T] of the persistent launch (coupling_tc.cu)
  int* range_flag = nullptr;   // sticky device flag: an fp16-split operand left the fp16 range (gnf_flow_range_flag)

  int mlp_index(int which, int half, int step) const {
    int T = d.weight_sharing ? 1 : d.num_timesteps;
    int s = d.weight_sharing ? 0 : step;
    return (which * 2 + half) * T + s;
  }
};

// flow.cu (shared with backward.cu)
int fwd_linear(const float* A, const float* W, const float* b, float* C, int64_t M, int N, int K, int act,
               cudaStream_t stream);
struct AttnBufs {      // per-GNN intermediates of the f1 attention block (all [n, *_pad] fp32)
  float *xq, *qbuf, *kbuf, *vbuf, *att, *proj;
  float* stats;        // optional [n, heads, 3]: segment max, segment sum (written by k_dm_attn), dot (backward)
  int32_t* fallback;   // optional [n / 32 + 1] scratch of the staged attention kernel (null: thread-per-head kernel only)
};
int fwd_attn_input(const Flow& f, int mlp, const float* xa, int64_t n, const int32_t* rowptr,
                   const int32_t* csr_senders, const AttnBufs& w, float* hbuf, cudaStream_t stream,
                   int math = GNF_MATH_FP32);
int fwd_layer_norm(const Flow& f, int mlp, float* x, int64_t n, cudaStream_t stream);
int fwd_agg_input(const Flow& f, const float* xa, int64_t n, const int32_t* rowptr, const int32_t* csr_senders,
                  float* hbuf, cudaStream_t stream);

// pack.cu
int pack_build_jobs(Flow& f);
int pack_all(const Flow& f, const float* params, cudaStream_t stream);
int pack_tc_image(const float* W, int k, int n, int kpad, int npad, int nhc, int kcc, uint8_t* img_f16, uint8_t* img_bf16,
                  void* job_scratch, cudaStream_t stream);

// backward.cu (shared with backward_tc.cu)
int bwd_agg_transpose(const Flow& f, const float* gh, int gh_stride, const int32_t* rowptr_s,
                      const int32_t* csr_receivers, const int32_t* rowptr_r, int64_t n, float* gxa,
                      cudaStream_t stream);
int bwd_split_scale(const float* z, int64_t n, int d, int h, int hp, float scale, float* x0, float* x1, float* g0,
                    float* g1, cudaStream_t stream);
int bwd_merge(const float* x0, const float* x1, int64_t n, int d, int h, int hp, float* x, cudaStream_t stream);

// linear_tc.cu
// gemm_tc.cu
void tc_gemm_geometry(int k, int n, int& kpad, int& npad, int& nb);
size_t tc_gemm_image_bytes(int k, int n);
int tc_gemm(const Flow& f, int math, const float* A, int lda, int kvalid, const uint8_t* img_f16, const uint8_t* img_bf16,
            int k, int n, const float* bias, int act, float* C, int ldc, int nvalid, int64_t M, cudaStream_t stream);
bool tc_linear_shape_ok(int k, int n);
size_t tc_linear_image_bytes(int k, int n);
int tc_linear(const Flow& f, int math, const float* A, int lda, int kvalid, const uint8_t* img_f16, const uint8_t* img_bf16,
              int k, int n, const float* bias, int act, float* C, int ldc, int nvalid, int64_t M, cudaStream_t stream,
              int accumulate = 0);

// backward_tc.cu
bool tc_bwd_supported(const Flow& f);
size_t tc_bwd_workspace(const Flow& f, int64_t n);
void tc_bwd_layout(const Flow& f, int64_t n, int64_t* out);
int tc_grevnet_backward(const Flow& f, const float* z, int64_t n, const int32_t* rowptr, const int32_t* csr_senders,
                        const int32_t* rowptr_s, const int32_t* csr_receivers, double loss_scale, float* grads,
                        float* x_out, void* ws, size_t ws_bytes, int dw_parts, int fwd_f16, void* stream);
int tc_half_backward(const Flow& f, int half, int step, const float* xa, float* xb, float* ga, float* gb, int64_t n,
                     const int32_t* rowptr, const int32_t* csr_senders, const int32_t* rowptr_s,
                     const int32_t* csr_receivers, double loss_scale, float* grads, void* ws, int dw_parts, int fwd_f16,
                     void* stream);
bool tc_bwd_inject_supported(const Flow& f);
size_t tc_bwd_inject_workspace(const Flow& f, int64_t n);
void tc_bwd_inject_buffers(const Flow& f, int64_t n, void* ws, float* pre0[2], float* d0[2]);
int tc_half_backward_inject(const Flow& f, int ms, int mt, const float* const hin[2], int in_ld, float* xb, float* gb,
                            int64_t n, double loss_scale, float* grads, void* ws, int dw_parts, int fwd_f16, void* stream);
int tc_dw_gemm_test(const float* A, const float* B, int64_t n, int fa, int fb, int parts, int n_splits, float* out,
                    void* ws, size_t ws_bytes, void* stream);

// coupling_tc.cu
size_t tc_bytes_per_mlp(int L, int K);
bool tc_shape_supported(const Flow& f);
void tc_set_trace(void* buf);
int tc_kernel_timing(int enable);
int tc_kernel_time(double* total_ms, int64_t* launches);
bool tc_inject_supported(const Flow& f);
int tc_coupling_inject(const Flow& f, int mlp_s, int mlp_t, int math, int inverse, const float* pre0_s,
                       const float* pre0_t, float* xb, int64_t n_nodes, double* ldj_partials, double* ldj_accum,
                       unsigned int* counter, void* stream);
int tc_build_half_tables(Flow& f);
bool tc_persistent_wanted(int64_t n_nodes);
int tc_flow_persistent(const Flow& f, int math, int inverse, float* x0, float* x1, int64_t n_nodes,
                       const int32_t* rowptr, const int32_t* csr_senders, double* ldj_partials, double* ldj_accum,
                       unsigned int* counter, void* stream);
int tc_coupling_half(const Flow& f, int mlp_s, int mlp_t, int math, int inverse,
                     const float* xa, float* xb, int64_t n_nodes,
                     const int32_t* rowptr, const int32_t* csr_senders,
                     double* ldj_partials, double* ldj_accum, unsigned int* counter, void* stream);

}  // namespace gnf

struct gnf_flow {
  gnf::Flow f;
};
