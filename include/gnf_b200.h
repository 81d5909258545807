/*
 * gnf_b200.h -- C ABI of the B200-native GRevNet hot path (libgnf_b200.so).
 *
 * Drop-in boundary for the one path this repository accelerates: GRevNet forward
 * (+log-det), inverse and log-prob over packed batches of graphs, with the
 * message-passing GNN (sender gather -> segment reduce by receiver -> node MLP)
 * inside every affine coupling.  Reference: jliu/graph-normalizing-flows @ d8b9256.
 * Each entry point cites the reference interface (file:line under /root/reference)
 * it replaces.  The arithmetic the reference delegates to graph_nets / Sonnet /
 * TF / TFP (not vendored) is named where it applies.
 *
 * Conventions
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer unless
 *     the name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - all calls are asynchronous on `stream`; none synchronises the device, except
 *     gnf_flow_create / gnf_flow_destroy (cudaMalloc / cudaFree).
 *   - return value: 0 on success, a negative GNF_E* code otherwise;
 *     gnf_last_error() returns a thread-local message for the last failure.
 *   - inputs are never written; outputs never alias inputs unless stated.
 *   - node features are row-major float32 [N, D] exactly as graph_nets' GraphsTuple.nodes
 *     (train_grevnet_with_data.py:265-271); indices are int32 as GraphsTuple.senders/receivers.
 */
#ifndef GNF_B200_H_
#define GNF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNF_ABI_VERSION 5

/* status codes */
#define GNF_OK            0
#define GNF_EINVAL       -1   /* bad argument (shape, null pointer, unsupported size)        */
#define GNF_ECUDA        -2   /* CUDA runtime error (message in gnf_last_error)              */
#define GNF_EUNSUPPORTED -3   /* shape outside what the requested kernel family supports     */
#define GNF_EWORKSPACE   -4   /* workspace too small                                         */

/* aggregation reducers: tf.unsorted_segment_sum / tf.unsorted_segment_mean (gnn.py:239-257) */
#define GNF_AGG_SUM  0
#define GNF_AGG_MEAN 1
/* node blocks: ConcatThenMLPBlock (gnn.py:100-111) / AggThenMLPBlock (gnn.py:114-126) */
#define GNF_BLOCK_CONCAT   0
#define GNF_BLOCK_AGG_THEN 1
/* f1  DMSelfAttentionMLP (gnn.py:480-573): edge-softmax multi-head attention over the in-edges,
 * head-concat projection, then concat-with-input + MLP.  fp32 arithmetic only. */
#define GNF_BLOCK_DM_ATTN  2
#define GNF_ATTN_CONCAT    1   /* concat=True   gnn.py:547-548 */
#define GNF_ATTN_RESIDUAL  2   /* residual=True gnn.py:551-552 */
#define GNF_ATTN_KQ_DIV    4   /* kq_dim_division=True gnn.py:462-464 */
#define GNF_ATTN_LAYER_NORM 8  /* layer_norm=True: snt.LayerNorm on the GNN output (gnn.py:554-556), eps 1e-5, per-GNN
                                  gamma[H], beta[H] appended to the GNN's parameters after the MLP */
/* MLP activations: tf.nn.leaky_relu alpha=0.2 (run_grevnet.py:158) / tf.nn.relu (gnn.py:162) */
#define GNF_ACT_LEAKY_RELU 0
#define GNF_ACT_RELU       1
/* arithmetic of the MLP contraction */
#define GNF_MATH_FP32   0  /* FFMA fp32, layer-by-layer kernels, any supported shape           */
#define GNF_MATH_TC3X   1  /* tcgen05 kind::f16, fp16 hi/lo split, 3 MMAs per product, fp32 acc */
#define GNF_MATH_BF16   2  /* tcgen05 kind::f16, bf16 single pass, fp32 accumulate             */
#define GNF_MATH_TC3X_BF16 3 /* as TC3X with a bf16 hi/lo split (no fp16 range limit)          */
#define GNF_MATH_TC2X   4  /* fp16 activations (one rounding, unbiased) x fp16 hi/lo weights: 2 MMAs per
                              product; log-prob parity holds, per-element z error ~1e-4            */

int         gnf_abi_version(void);
const char* gnf_last_error(void);
/* number of kernels this library launched on the calling thread since the last reset
 * (bench.py's "gpu_launches") */
int64_t     gnf_launch_count(int reset);
/* Developer aid: when device_buf (>= 10*2048 uint64, zeroed) is non-NULL, CTA 0 of every fused
 * coupling kernel records a clock64 timeline of its warp roles into it; NULL switches it off. */
int         gnf_debug_set_trace(void* device_buf);
/* Device-side duration of every fused coupling launch (k_coupling_tc): enable = 1 resets the slots and starts recording
 * (up to 8192 launches), 0 stops.  Each launch writes {first CTA past its dependency wait, last CTA done} from
 * %globaltimer into its own slot -- no host events between launches, so a benchmark can keep this on inside its timed
 * steps.  gnf_debug_kernel_time synchronises the device and returns the summed duration and the launch count.
 * Host-side state, not thread safe. */
int         gnf_debug_kernel_timing(int32_t enable);
int         gnf_debug_kernel_time(double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------------------------------
 * a1  batch structure.  Replaces the per-call index handling of graph_nets' aggregator:
 * CSR by receiver, STABLE (in-segment order = ascending edge index), so the in-order serial
 * accumulation reproduces TF-CPU UnsortedSegmentSum bit for bit.
 *   rowptr[N+1], perm[E] (edge ids in CSR order), csr_senders[E] = senders[perm[.]]
 * workspace: gnf_build_csr_workspace(N, E) bytes.
 * ------------------------------------------------------------------------------------------ */
size_t gnf_build_csr_workspace(int64_t n_nodes, int64_t n_edges);
int gnf_build_csr(const int32_t* receivers, const int32_t* senders, int64_t n_nodes, int64_t n_edges,
                  int32_t* rowptr, int32_t* perm, int32_t* csr_senders,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Error behaviour of the reference (TF InvalidArgumentError on out-of-range gather/segment ids,
 * gnn.py:151 / gnn.py:103): counts indices outside [0, n_nodes) into *bad_count (device int32). */
int gnf_validate_indices(const int32_t* senders, const int32_t* receivers, int64_t n_nodes,
                         int64_t n_edges, int32_t* bad_count, void* stream);

/* a3  gn.blocks.EdgeBlock(IdentityModule, use_sender_nodes) (gnn.py:135-140,151-152):
 * edges[e,:] = x[senders[e],:].  x [N,H] f32, edges [E,H] f32. */
int gnf_gather_rows(const float* x, int32_t h, const int32_t* senders, int64_t n_edges,
                    float* edges, void* stream);

/* a4  gn.blocks.ReceivedEdgesToNodesAggregator(reducer) (gnn.py:103-104,117-118) on a
 * materialised edge tensor: out[r,:] = sum_{e: receivers[e]=r} edges[e,:] in ascending e
 * (mean: / max(count,1)).  Needs rowptr/perm from gnf_build_csr. */
int gnf_segment_sum(const float* edges, int32_t h, const int32_t* rowptr, const int32_t* perm,
                    int64_t n_nodes, int32_t agg, float* out, void* stream);

/* a3+a4 fused (what NodeBlockGNN._build, gnn.py:155-156, amounts to before the MLP):
 * out[r,:] = reduce_{e: receivers[e]=r} x[senders[e],:].  Same order, same bits as
 * gnf_gather_rows followed by gnf_segment_sum. */
int gnf_gather_segment_sum(const float* x, int32_t h, const int32_t* rowptr,
                           const int32_t* csr_senders, int64_t n_nodes, int32_t agg,
                           float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a2  GRevNet.__init__ (gnn.py:274-302): the 4*T message-passing GNNs s[2][T], t[2][T]
 * (4 when weight_sharing), each a NodeBlockGNN (gnn.py:143-156) around one
 * make_mlp_model MLP (gnn.py:159-180).
 *
 * params: flat float32 DEVICE buffer, MLPs in the order
 *     which (0 = s, 1 = t) -> half (0, 1) -> step (0..T-1; one entry when weight_sharing)
 * each MLP as  W0[in,L] b0[L]  W1[L,L] b1[L] ... W_{K-1}[L,H] b_{K-1}[H],  row-major,
 * Sonnet Linear convention y = x @ W + b;  in = D (concat) or D/2 (agg_then), H = D/2.
 * GNF_BLOCK_DM_ATTN: every GNN is  Wq[H, heads*kq]  Wk[H, heads*kq]  Wv[H, v]  Wo[heads*v, out]
 * (all without bias, gnn.py:509-545) followed by its MLP with in = H + out (concat) or out, and, with
 * GNF_ATTN_LAYER_NORM, the LayerNorm gamma[H] beta[H].
 * ------------------------------------------------------------------------------------------ */
typedef struct gnf_flow_desc {
  int32_t num_timesteps;       /* T   GRevNet(num_timesteps)       gnn.py:276 */
  int32_t node_embedding_dim;  /* D   GRevNet(node_embedding_dim)  gnn.py:277 (must be even) */
  int32_t latent_dim;          /* L   make_mlp_model(latent_dim)   gnn.py:159 */
  int32_t num_layers;          /* K   make_mlp_model(num_layers)   gnn.py:161 (>= 2) */
  int32_t agg;                 /* GNF_AGG_*   */
  int32_t block;               /* GNF_BLOCK_* */
  int32_t act;                 /* GNF_ACT_*   */
  int32_t weight_sharing;      /* gnn.py:279,283-286 */
  float   eps;                 /* AggThenMLPBlock epsilon, gnn.py:121 */
  /* GNF_BLOCK_DM_ATTN only (dm_self_attn_gnn arguments, gnn.py:555-573); zero otherwise */
  int32_t attn_num_heads;      /* num_heads               */
  int32_t attn_kq_dim;         /* kq_dim   (<= 64)        */
  int32_t attn_v_dim;          /* v_dim    (<= 64)        */
  int32_t attn_out_dim;        /* concat_heads_output_dim */
  int32_t attn_flags;          /* GNF_ATTN_* */
} gnf_flow_desc;

typedef struct gnf_flow gnf_flow;   /* opaque: packed device-side weights */

int64_t gnf_flow_param_count(const gnf_flow_desc* desc);
int gnf_flow_create(gnf_flow** out, const gnf_flow_desc* desc);
/* (re)pack weights from the flat buffer into every image the kernels consume (padded fp32, transposed fp32, fp16 and
 * bf16 hi/lo tensor-core chunk images, their transposes, biases): ONE kernel launch; call after every optimiser step */
int gnf_flow_set_params(gnf_flow* flow, const float* params, void* stream);
/* fp16 range guard of GNF_MATH_TC3X / GNF_MATH_TC2X.  The reference computes in fp32 and never clamps exp(s)
 * (gnn.py:323), so MLP inputs / hidden activations beyond the fp16 range (|v| > 65504) are legal; the fp16 hi/lo split
 * turns them into inf/NaN.  The fused kernels therefore set a STICKY device flag whenever an operand they split exceeds
 * that range.  gnf_flow_range_flag enqueues on `stream` an asynchronous copy of the flag into *host_flag (pinned host
 * memory; may be NULL) and, if reset != 0, clears the device flag after the copy.  The host mirror raises
 * FloatingPointError and re-runs in GNF_MATH_TC3X_BF16 (same 3-MMA scheme, bf16 range = fp32 range). */
int gnf_flow_range_flag(const gnf_flow* flow, int32_t* host_flag, int32_t reset, void* stream);
int gnf_flow_destroy(gnf_flow* flow);
/* 1 if (flow shape, math) is served by the tcgen05 kernels in the density / sampling direction, 0 if only
 * GNF_MATH_FP32 is.  Two shapes qualify: MLP input <= 16 (message-passing blocks up to D = 16: everything fused), and
 * wider MLP inputs (dm_self_attn without residual / layer_norm; message passing with 16 < D <= 32), where the input
 * assembly and layer 0 run in the fp32 kernels and layers 1..K-1 + the coupling update in the fused kernel. */
int gnf_flow_supports(const gnf_flow* flow, int32_t math);
/* same question for gnf_grevnet_backward / gnf_coupling_half_backward (the tensor-core backward needs the fully fused
 * shape; everything else trains through GNF_MATH_FP32) */
int gnf_flow_supports_backward(const gnf_flow* flow, int32_t math);

/* ------------------------------------------------------------------------------------------
 * a7  GRevNet.f (gnn.py:304-341) / GRevNet.g (gnn.py:343-373) / _build(inverse) (gnn.py:379-381),
 * use_batch_norm = False.
 *   x, z       [N, D] f32 row-major
 *   ldj        device double[1]: sum over steps/halves/nodes/features of s  (gnn.py:322,337)
 * workspace: gnf_grevnet_workspace(flow, N, math) bytes, 256-byte aligned.
 * One fused kernel per half coupling step (aggregate + s-MLP + t-MLP + affine update +
 * log-det partial) when math != GNF_MATH_FP32.
 * ------------------------------------------------------------------------------------------ */
size_t gnf_grevnet_workspace(const gnf_flow* flow, int64_t n_nodes, int32_t math);
int gnf_grevnet_forward(const gnf_flow* flow, const float* x, int64_t n_nodes, int64_t n_edges,
                        const int32_t* rowptr, const int32_t* csr_senders,
                        float* z, double* ldj, int32_t math,
                        void* workspace, size_t workspace_bytes, void* stream);
int gnf_grevnet_inverse(const gnf_flow* flow, const float* z, int64_t n_nodes, int64_t n_edges,
                        const int32_t* rowptr, const int32_t* csr_senders,
                        float* x, int32_t math,
                        void* workspace, size_t workspace_bytes, void* stream);

/* One full coupling step i (both halves) of f / g on planar halves: x0, x1 [N, HP] with
 * HP = gnf_padded_half(D/2) floats per row (zero padded), updated in place.
 * ldj_accum (device double[1]) is ADDED to (forward only; may be NULL for inverse). */
int32_t gnf_padded_half(int32_t h);
int gnf_coupling_step(const gnf_flow* flow, int32_t step, int32_t inverse,
                      float* x0, float* x1, int64_t n_nodes, int64_t n_edges,
                      const int32_t* rowptr, const int32_t* csr_senders,
                      double* ldj_accum, int32_t math,
                      void* workspace, size_t workspace_bytes, void* stream);

/* One HALF coupling step: the s/t GNNs of (half, step) read xa and update xb in place
 * (forward: xb*exp(s)+t and *ldj_accum += sum(s); inverse: (xb-t)*exp(-s)).  gnn.py:320-323 is
 * (half=0: xa=x0, xb=x1), gnn.py:335-338 is (half=1: xa=x1, xb=x0).  Used by the batch-norm
 * variant of the flow, which interleaves gnf_bn_moments / gnf_affine_rows between half steps. */
int gnf_coupling_half(const gnf_flow* flow, int32_t half, int32_t step, int32_t inverse,
                      const float* xa, float* xb, int64_t n_nodes, int64_t n_edges,
                      const int32_t* rowptr, const int32_t* csr_senders,
                      double* ldj_accum, int32_t math,
                      void* workspace, size_t workspace_bytes, void* stream);

/* tf.split(x.nodes, 2, axis=1) (gnn.py:306) into planar zero-padded halves [N, HP], and
 * tf.concat([x0, x1], 1) (gnn.py:340) back. */
int gnf_split_halves(const float* x, int64_t n_nodes, int32_t d, float* x0, float* x1, void* stream);
int gnf_merge_halves(const float* x0, const float* x1, int64_t n_nodes, int32_t d, float* x, void* stream);

/* a9  pieces of tfb.BatchNormalization (gnn.py:260-263,310-313,325-328,356-358,369-371) on a
 * planar half x [N, HP]:
 *   gnf_bn_moments : sums[0:H] = sum_n x[n,f], sums[H:2H] = sum_n x[n,f]^2, sums[2H] = n_nodes  (device
 *                    double[2H+1]; sums, not moments, so that ranks can all-reduce them before dividing)
 *   gnf_bn_finalize: the [H]-sized bookkeeping in one launch: scale_shift[0:H] = gamma/sqrt(var+eps),
 *                    scale_shift[H:2H] = beta - mean*scale (inputs of gnf_affine_rows); *ldj_accum +=
 *                    n_local * sum_f(log gamma - 1/2 log(var+eps)) (the N-tiled ildj, this rank's share);
 *                    stats = {mean[H], var[H], N} (double, kept for the backward); moving statistics updated
 *                    with `momentum` (< 0: left alone)
 *   gnf_affine_rows: x[n,f] <- x[n,f] * scale[f] + shift[f]  (normalise / de-normalise) */
size_t gnf_bn_moments_workspace(int32_t h);
int gnf_bn_moments(const float* x, int64_t n_nodes, int32_t h, double* sums,
                   void* workspace, size_t workspace_bytes, void* stream);
int gnf_bn_finalize(const double* sums, int32_t h, const float* gamma, const float* beta, double eps,
                    double n_local, double* ldj_accum, float* scale_shift, double* stats,
                    float* moving_mean, float* moving_var, float momentum, void* stream);
int gnf_affine_rows(float* x, int64_t n_nodes, int32_t h, const float* scale, const float* shift, void* stream);
/* The whole batch-norm flow in ONE call (single-rank batches: GRevNet(use_batch_norm=True), the default of both scripts,
 * run_grevnet.py:83 / train_grevnet_with_data.py:109): the chain the host mirror would otherwise drive half step by half
 * step -- density direction gnn.py:309-338 (per half step: gnf_bn_moments, gnf_bn_finalize, gnf_affine_rows on the
 * conditioning half, then gnf_coupling_half), sampling direction gnn.py:347-372 (gnf_coupling_half inverse, then the
 * de-normalisation with the moving statistics).  gamma, beta, moving_mean, moving_var: device float [2, T, H] (half,
 * step, feature); stats: device double [2, T, 2H+1] = {mean, var, N} of every half step (kept for the backward);
 * momentum < 0 leaves the moving statistics alone.  bn_workspace >= gnf_grevnet_bn_workspace(flow) bytes, workspace as
 * gnf_grevnet_forward.  Sharded runs all-reduce the moment sums between gnf_bn_moments and gnf_bn_finalize and keep
 * using the per-half-step entries. */
size_t gnf_grevnet_bn_workspace(const gnf_flow* flow);
int gnf_grevnet_forward_bn(const gnf_flow* flow, const float* x, int64_t n_nodes, int64_t n_edges,
                           const int32_t* rowptr, const int32_t* csr_senders, const float* gamma, const float* beta,
                           float* moving_mean, float* moving_var, double eps, float momentum, float* z, double* ldj,
                           double* stats, int32_t math, void* workspace, size_t workspace_bytes, void* bn_workspace,
                           size_t bn_workspace_bytes, void* stream);
int gnf_grevnet_inverse_bn(const gnf_flow* flow, const float* z, int64_t n_nodes, int64_t n_edges,
                           const int32_t* rowptr, const int32_t* csr_senders, const float* gamma, const float* beta,
                           const float* moving_mean, const float* moving_var, double eps, float* x, int32_t math,
                           void* workspace, size_t workspace_bytes, void* bn_workspace, size_t bn_workspace_bytes,
                           void* stream);
/* Backward of the bijector in training mode (batch statistics depend on x), for loss = -loss_scale * log_prob_xs:
 *   gnf_bn_backward_sums : sums[0:H] = sum_n G_y[n,f], sums[H:2H] = sum_n G_y[n,f] * xhat[n,f],
 *                          xhat = (y - beta) * inv_gamma   (device double[2H]; all-reduce across ranks before use;
 *                          also d/d beta and the data term of d/d gamma).  workspace: gnf_bn_moments_workspace(H).
 *   gnf_bn_backward_apply: in place, coef = 7 rows of H floats {beta, 1/gamma, c1, c2, c3, s, mu}:
 *                          g <- c1*g + c2*xhat + c3   (c1 = gamma/s, c2 = (loss_scale - gamma*S2/N)/s, c3 = -gamma*S1/(N*s),
 *                          s = sqrt(var + eps): the batch-statistics terms and the -N/2 log(var+eps) log-det term)
 *                          y <- xhat*s + mu           (the bijector undone with the statistics saved by the forward) */
/* gnf_bn_backward_coef: with sums == NULL writes inv_gamma[H] = 1/gamma (input of gnf_bn_backward_sums); with the
 * (all-reduced) sums it writes the 7 coefficient rows of gnf_bn_backward_apply from stats = {mean, var, N} and
 * accumulates dL/dgamma = S2 - loss_scale*N/gamma, dL/dbeta = S1 into g_gamma / g_beta (device double[H]). */
int gnf_bn_backward_coef(const double* sums, const double* stats, int32_t h, const float* gamma, const float* beta,
                         double eps, double loss_scale, float* coef, float* inv_gamma, double* g_gamma,
                         double* g_beta, void* stream);
int gnf_bn_backward_sums(const float* y, const float* g, int64_t n_nodes, int32_t h, const float* beta,
                         const float* inv_gamma, double* sums, void* workspace, size_t workspace_bytes, void* stream);
int gnf_bn_backward_apply(float* y, float* g, int64_t n_nodes, int32_t h, const float* coef, void* stream);

/* One message-passing GNN of the flow on its own: NodeBlockGNN._build (gnn.py:155-156) =
 * node_block(edge_block(graph)); x, out [N, D/2] f32.  which: 0 = s, 1 = t.  fp32 arithmetic.
 * workspace: gnf_grevnet_workspace(flow, N, GNF_MATH_FP32). */
int gnf_gnn_forward(const gnf_flow* flow, int32_t which, int32_t half, int32_t step,
                    const float* x, int64_t n_nodes, int64_t n_edges,
                    const int32_t* rowptr, const int32_t* csr_senders, float* out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * f2  Training-step gradients, reversible (what GNFBlock(use_efficient_backprop=True),
 * run_grevnet.py:46,288, did by name; the optimiser side is run_grevnet.py:345-377).
 *   loss = -loss_scale * (sum_n log N(z_n; 0, I) + log_det_jacobian)      run_grevnet.py:292-296
 *          (loss_scale = 1 for total_loss, 1/N for loss_per_node, train_grevnet_with_data.py:353-355)
 * Input is z = f(x) from gnf_grevnet_forward; nothing else is kept from the forward pass: each half
 * step is undone with the inverse update and its MLP activations are recomputed.
 *   grads     float[gnf_flow_param_count], same layout as params; ACCUMULATED into (zero it first)
 *   x_out     optional [N, D]: the reconstructed input (equals x up to rounding)
 *   rowptr / csr_senders            CSR by receiver (gnf_build_csr(receivers, senders))
 *   rowptr_by_sender / csr_receivers CSR by sender   (gnf_build_csr(senders, receivers)): the
 *                                   transpose the aggregation's backward walks
 * math (use_batch_norm = False):
 *   GNF_MATH_FP32                     layered FFMA kernels, any supported shape, including GNF_BLOCK_DM_ATTN
 *                                     (attention backward: softmax statistics per receiver, then a gather
 *                                     over the CSR by sender; projections Wq Wk Wv Wo get gradients too);
 *   any other mode                    tcgen05 path (flows gnf_flow_supports() accepts): per half step one fused
 *                                     kernel (recompute s,t; undo the update; dX chain with the transposed
 *                                     weights; bias gradients) + one weight-gradient GEMM kernel, fp32
 *                                     accumulate, hi/lo split 16-bit operands, 3 MMAs per product.  Gradient
 *                                     operands are always bf16 hi/lo (fp16 would underflow); the recomputed
 *                                     forward chains are fp16 hi/lo for GNF_MATH_TC3X / GNF_MATH_TC2X (fp32-class
 *                                     pre-activations, so act' masks match fp32 arithmetic) and bf16 hi/lo for
 *                                     GNF_MATH_TC3X_BF16 / GNF_MATH_BF16.  GNF_MATH_TC2X / GNF_MATH_BF16
 *                                     additionally run the weight-gradient GEMM on the hi parts only (one MMA
 *                                     per product; its rounding errors average over the nodes).
 * workspace: gnf_grevnet_backward_workspace(flow, N, math), 256-byte aligned.
 * ------------------------------------------------------------------------------------------ */
size_t gnf_grevnet_backward_workspace(const gnf_flow* flow, int64_t n_nodes, int32_t math);
int gnf_grevnet_backward(const gnf_flow* flow, const float* z, int64_t n_nodes, int64_t n_edges,
                         const int32_t* rowptr, const int32_t* csr_senders,
                         const int32_t* rowptr_by_sender, const int32_t* csr_receivers,
                         double loss_scale, float* grads, float* x_out, int32_t math,
                         void* workspace, size_t workspace_bytes, void* stream);

/* One reversed HALF step of gnf_grevnet_backward on planar halves [N, HP] (the batch-norm variant interleaves
 * gnf_bn_backward_* between half steps): xa is read; xb holds the post-update half and is restored in place;
 * gb holds dLoss/d(post-update xb) and becomes dLoss/d(xb); ga is ACCUMULATED into; grads as above.
 * workspace: gnf_grevnet_backward_workspace(flow, N, math). */
int gnf_coupling_half_backward(const gnf_flow* flow, int32_t half, int32_t step, const float* xa, float* xb,
                               float* ga, float* gb, int64_t n_nodes, int64_t n_edges,
                               const int32_t* rowptr, const int32_t* csr_senders,
                               const int32_t* rowptr_by_sender, const int32_t* csr_receivers,
                               double loss_scale, float* grads, int32_t math,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Weight-gradient GEMM of the tensor-core backward on its own (unit tests, profiling):
 * out[fa, fb] = a^T b for row-major fp32 a [n, fa], b [n, fb]; fa in {128,256}, fb in {16,128,256};
 * parts 2 = bf16 hi/lo split (3 MMAs per product), 1 = single bf16 pass; n_splits node ranges (one CTA
 * each) reduced in fixed order.  workspace >= 4*ceil(n/128)*128*(fa+fb) + 4*n_splits*fa*fb + 2048 bytes. */
/* Byte offsets inside the tensor-core backward workspace (tests decode the tile images left by the last
 * half step): out8 = {act_img, dlt_img, h_img, g_img, g_h, x0, g0, total bytes}. */
int gnf_debug_bwd_layout(const gnf_flow* flow, int64_t n_nodes, int64_t* out8);
int gnf_debug_dw_gemm(const float* a, const float* b, int64_t n, int32_t fa, int32_t fb, int32_t parts,
                      int32_t n_splits, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* One Sonnet Linear of the layered tensor-core path on its own (unit tests, profiling; gnn.py:143-180 make_mlp_model
 * layers of any width, e.g. the 2048-wide MLPs of train_grevnet_with_data.py:41-47):
 * c[m, n] = act(a[m, k] @ w[k, n] + bias[n]), row-major fp32, k % 4 == 0, n % 4 == 0, act = GNF_ACT_* or 2 (none),
 * math = GNF_MATH_TC3X / TC3X_BF16 / BF16.  The entry packs w into the kernel's image format inside the workspace
 * (>= gnf_debug_linear_tc_workspace(k, n) bytes) and runs k_gemm_tc (csrc/gemm_tc.cu). */
size_t gnf_debug_linear_tc_workspace(int32_t k, int32_t n);
int gnf_debug_linear_tc(const float* a, const float* w, const float* bias, int64_t m, int32_t k, int32_t n, int32_t act,
                        int32_t math, float* c, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * f3  Decode tail of the sampling pass (train_grevnet_with_data.py:414-416):
 * pred_adj(graph, scaled_hacky_sigmoid_l2) = sigmoid(temp * (shift - D / sqrt(d))) with the pairwise
 * squared L2 distance D, masked to the block diagonal and with a zero diagonal
 * (loss.py:154-159,45-53,131-151,83-85; temp = 10, shift = 1 for scaled_hacky_sigmoid_l2,
 * loss.py:56-62 sigmoid_l2 for other values).  Written per graph: out[adj_off[g] + i*n_g + j];
 * node_off[G+1], adj_off[G+1] are device int64 prefix sums of n_node and n_node^2.
 * ------------------------------------------------------------------------------------------ */
int gnf_pred_adj(const float* nodes, int32_t d, const int64_t* node_off, const int64_t* adj_off,
                 int64_t n_graphs, float temp, float shift, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a8  log-prob assembly (run_grevnet.py:292-296, train_grevnet_with_data.py:348-350):
 *   out[0] = log_prob_zs = sum_n ( -1/2 |z_n|^2 - D/2 ln 2pi )   (tfd.MultivariateNormalDiag)
 *   out[1] = log_det_jacobian (copied from ldj)
 *   out[2] = log_prob_xs = out[0] + out[1]
 *   out[3] = N   (sum n_node, the per-node normaliser of run_grevnet.py:298)
 * out: device double[4].  workspace: gnf_log_prob_workspace(N, D) bytes.
 * ------------------------------------------------------------------------------------------ */
size_t gnf_log_prob_workspace(int64_t n_nodes, int32_t d);
int gnf_log_prob(const float* z, int64_t n_nodes, int32_t d, const double* ldj, double* out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * e  The one collective of the path over NVLink peer memory (the reference has no distributed code):
 * all-reduce(SUM) of the fp64 4-vector of gnf_log_prob across the ranks of ONE node (world <= 8), one process per
 * GPU.  Every rank owns a small slot buffer, exported with cudaIpc; peers write into it with plain stores through
 * the NVLink mapping, a release/acquire flag word orders them, the sum is taken in rank order (bit-identical on all
 * ranks).  The kernels are one warp wide and co-reside with the fused coupling kernel, which an NCCL kernel cannot.
 *   gnf_peer_create   allocates this rank's slots, returns the 64-byte cudaIpcMemHandle_t in handle_out (HOST)
 *   gnf_peer_connect  handles_host = the `world` handles in rank order (exchange them with any host-side all-gather)
 *   gnf_peer_allreduce4      vec (device double[4]) all-reduced in place, one tiny kernel on `stream`
 *   gnf_log_prob_allreduce   gnf_log_prob AND the all-reduce of its result in ONE kernel: out holds the global
 *                            (log_prob_zs, log_det_jacobian, log_prob_xs, num_nodes) on every rank
 * Every rank must issue the same sequence of peer calls.  A peer that never arrives turns the result into NaN after
 * ~2 s instead of hanging the device.
 * ------------------------------------------------------------------------------------------ */
#define GNF_PEER_HANDLE_BYTES 64
typedef struct gnf_peer gnf_peer;
int gnf_peer_create(gnf_peer** out, int32_t rank, int32_t world, uint8_t* handle_out);
int gnf_peer_connect(gnf_peer* peer, const uint8_t* handles_host);
int gnf_peer_destroy(gnf_peer* peer);
int gnf_peer_allreduce4(gnf_peer* peer, double* vec, void* stream);
int gnf_log_prob_allreduce(const float* z, int64_t n_nodes, int32_t d, const double* ldj, double* out,
                           void* workspace, size_t workspace_bytes, gnf_peer* peer, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNF_B200_H_ */
