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
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

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
  // tensor-core packed weights (two images: fp16 and bf16 element type), per MLP a stream of
  // shared-memory chunk images in consumption order (see coupling_tc.cu)
  uint8_t* wtc[2] = {nullptr, nullptr};   // [0]=fp16 hi/lo, [1]=bf16 hi/lo
  int64_t wtc_per_mlp = 0;
  float* btc = nullptr;        // biases for the TC kernel: per MLP K*256 floats
  bool tc_ok = false;

  int mlp_index(int which, int half, int step) const {
    int T = d.weight_sharing ? 1 : d.num_timesteps;
    int s = d.weight_sharing ? 0 : step;
    return (which * 2 + half) * T + s;
  }
};

// coupling_tc.cu
size_t tc_bytes_per_mlp(int L, int K);
int tc_pack_mlp(const Flow& f, int mlp, const float* params, void* stream);
bool tc_shape_supported(const Flow& f);
void tc_set_trace(void* buf);
int tc_coupling_half(const Flow& f, int mlp_s, int mlp_t, int math, int inverse,
                     const float* xa, float* xb, int64_t n_nodes,
                     const int32_t* rowptr, const int32_t* csr_senders,
                     double* ldj_partials, int* n_partials, void* stream);

}  // namespace gnf
